#!/bin/bash
# last GPU call of round 2: headline bench lines on the final library (the GPU test suite ran on the same kernels in the call before)
mkdir -p gpurun_out
timeout 200 python bench.py 2>gpurun_out/bench_err.txt | tail -1 > gpurun_out/r2_bench.json; cut -c1-1200 gpurun_out/r2_bench.json; tail -3 gpurun_out/bench_err.txt
timeout 100 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2_bench_k20.json; cut -c1-400 gpurun_out/r2_bench_k20.json
timeout 100 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sync-steps 2>/dev/null | tail -1 > gpurun_out/r2_bench_sync.json; cut -c1-400 gpurun_out/r2_bench_sync.json
timeout 120 python bench.py --workload ba_stress --problems 64 --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r2_ba_stress.json; cut -c1-600 gpurun_out/r2_ba_stress.json
