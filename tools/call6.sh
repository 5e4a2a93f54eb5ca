#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_pipeline.py tests/test_gpu_zinit.py tests/test_gpu_zplayer.py tests/test_gpu_zundistort.py -x -q -m gpu -p no:cacheprovider -s > gpurun_out/c6_tests.txt 2>&1
grep -E "config 2|prior form|passed|failed|Error|assert" gpurun_out/c6_tests.txt | head -20
XRVIO_KPROF=1 python xrslam_b200/build.py --force > /dev/null 2>&1
timeout 600 python tools/phase_probe.py 64 > gpurun_out/c6_phase_probe.txt 2>&1
python xrslam_b200/build.py --force > /dev/null 2>&1
cat gpurun_out/c6_phase_probe.txt
timeout 400 python bench.py --steps 60 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-1700 > gpurun_out/c6_bench.txt; cat gpurun_out/c6_bench.txt
timeout 400 python bench.py --steps 60 --warmup 3 --no-cpu-baseline --sync-steps 2>&1 | tail -1 | cut -c1-1700 > gpurun_out/c6_bench_sync.txt; cat gpurun_out/c6_bench_sync.txt
XRVIO_KLT_TMA=0 timeout 400 python bench.py --steps 60 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-1700 > gpurun_out/c6_bench_notma.txt; cat gpurun_out/c6_bench_notma.txt
