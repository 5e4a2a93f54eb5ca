#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_zz_large_window.py tests/test_gpu_pipeline.py -x -q -m gpu -p no:cacheprovider -k "solver or marg or large or golden or long_stream or hit_the" > gpurun_out/c11_tests.txt 2>&1
tail -3 gpurun_out/c11_tests.txt
XRVIO_SLOW_STEP_MS=9 timeout 600 python tools/phase_probe.py 64 > gpurun_out/c11_phase_probe.txt 2> gpurun_out/c11_slow.txt
grep -E "^steps|lane-step|slowest" gpurun_out/c11_phase_probe.txt; grep -v "p10=" gpurun_out/c11_slow.txt | head -30 | cut -c1-900
XRVIO_KPROF=1 python xrslam_b200/build.py --force > /dev/null 2>&1
timeout 600 python tools/phase_probe.py 64 2>/dev/null | tail -24
python xrslam_b200/build.py --force > /dev/null 2>&1
