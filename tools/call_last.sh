#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider -s -rs > gpurun_out/r2_gpu_tests.txt 2>&1
grep -E "config 2|prior form|passed|failed|skipped" gpurun_out/r2_gpu_tests.txt | head
XRVIO_KPROF=1 python xrslam_b200/build.py --force > /dev/null 2>&1
timeout 400 python tools/phase_probe.py 64 2>/dev/null > gpurun_out/r2_phase_probe.txt; grep -E "^steps|lane-step" gpurun_out/r2_phase_probe.txt; tail -22 gpurun_out/r2_phase_probe.txt
python xrslam_b200/build.py --force > /dev/null 2>&1
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2_bench_k20.json; cut -c1-900 gpurun_out/r2_bench_k20.json
