#!/bin/bash
# One gpurun call that collects everything round 1 left unmeasured:
#   gpurun --timeout 2400 -- 'bash tools/round2_first_call.sh'
# Outputs under gpurun_out/: gpu_tests.log (every GPU test under its own timeout), bench.json, bench_ba_stress.json,
# fp64_probe.txt (tools/fp64_probe + the library probe inside bench_ba_stress.json), launches.csv (ncu launch list of the bench command), ncu_solve_large.ncu-rep (+ raw csv).
mkdir -p gpurun_out
PER_TEST_TIMEOUT=${PER_TEST_TIMEOUT:-240} bash tools/gpu_tests.sh > gpurun_out/gpu_tests_summary.txt 2>&1
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 900 python bench.py --workload ba_stress --problems 64 --steps 3 --warmup 1 > gpurun_out/bench_ba_stress.json 2> gpurun_out/bench_ba_stress.err
# A/B at a smaller batch: is one problem per SM latency bound?  (time per problem should not depend on the batch below 148)
timeout 600 python bench.py --workload ba_stress --problems 8 --steps 3 --warmup 1 > gpurun_out/bench_ba_stress_8.json 2>> gpurun_out/bench_ba_stress.err
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_probe tools/fp64_probe.cu && timeout 120 /tmp/fp64_probe > gpurun_out/fp64_probe.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:solve_large_kernel -c 1 -o gpurun_out/ncu_solve_large \
    python bench.py --workload ba_stress --problems 64 --steps 1 --warmup 0 > gpurun_out/ncu_solve_large.log 2>&1
ncu -i gpurun_out/ncu_solve_large.ncu-rep --page raw --csv > gpurun_out/ncu_solve_large_raw.csv 2>/dev/null
tail -3 gpurun_out/gpu_tests_summary.txt; cat gpurun_out/bench.json gpurun_out/bench_ba_stress.json gpurun_out/bench_ba_stress_8.json
