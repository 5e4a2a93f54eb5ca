#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider -s -rs > gpurun_out/r2_gpu_tests.txt 2>&1
tail -2 gpurun_out/r2_gpu_tests.txt
timeout 300 python bench.py --workload ba_stress --problems 64 --steps 3 --warmup 1 > gpurun_out/r2_ba_stress.json 2>/dev/null; cut -c1-420 gpurun_out/r2_ba_stress.json
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/r2_bench.json 2>/dev/null; cut -c1-1200 gpurun_out/r2_bench.json
bash tools/profile_ncu.sh r2
