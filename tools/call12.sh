#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_zz_large_window.py tests/test_gpu_pipeline.py -x -q -m gpu -p no:cacheprovider > gpurun_out/c12_tests.txt 2>&1
tail -3 gpurun_out/c12_tests.txt
XRVIO_KPROF=1 python xrslam_b200/build.py --force > /dev/null 2>&1
timeout 600 python tools/phase_probe.py 64 2>/dev/null > gpurun_out/c12_phase_probe.txt; grep -E "^steps|lane-step" gpurun_out/c12_phase_probe.txt; tail -24 gpurun_out/c12_phase_probe.txt
python xrslam_b200/build.py --force > /dev/null 2>&1
for mode in "" "--sync-steps"; do
timeout 400 python bench.py --steps 100 --warmup 3 --no-cpu-baseline $mode 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('$mode', {k:d[k] for k in ('value','ms_per_step','lane_busy_ms','stepping')}, d['e2e']['value'], d['roofline']['family_ms'])"
done > gpurun_out/c12_bench.txt 2>&1
cat gpurun_out/c12_bench.txt
