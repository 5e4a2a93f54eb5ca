#!/bin/bash
mkdir -p gpurun_out
nproc > gpurun_out/c7_host.txt; lscpu | grep -E "Model name|Socket|NUMA node" >> gpurun_out/c7_host.txt; cat gpurun_out/c7_host.txt
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_zz_large_window.py tests/test_gpu_pipeline.py -x -q -m gpu -p no:cacheprovider -s > gpurun_out/c7_tests.txt 2>&1
grep -E "config 2|prior form|passed|failed|Error|assert" gpurun_out/c7_tests.txt | head -20
XRVIO_KPROF=1 python xrslam_b200/build.py --force > /dev/null 2>&1
timeout 600 python tools/phase_probe.py 64 > gpurun_out/c7_phase_probe.txt 2>&1
python xrslam_b200/build.py --force > /dev/null 2>&1
cat gpurun_out/c7_phase_probe.txt
for mode in "" "--sync-steps"; do
timeout 400 python bench.py --steps 60 --warmup 3 --no-cpu-baseline $mode 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('$mode', {k:d[k] for k in ('value','ms_per_step','lane_busy_ms','host_cpus','stepping')}, d['e2e']['value'], d['roofline']['family_ms'])"
done > gpurun_out/c7_bench.txt 2>&1
cat gpurun_out/c7_bench.txt
