#!/bin/bash
# One gpurun call that refreshes every committed measurement (run on the GPU box):
#   gpurun --timeout 2400 -- 'bash tools/profile_final.sh r2'
# Everything lands in gpurun_out/ as small text / gzip files (the .ncu-rep files stay in /tmp on the box, except the one
# source-level capture of the solve kernel): tools/summarize_ncu.py turns them into profiles/<tag>_*.
TAG=${1:-r2}
KREGEX='regex:(clahe|pyrdown|scharr|harris|corner_select|klt_|poisson|preint|solve_|marg_|tri_kernel|undistort|gram_)'
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > gpurun_out/${TAG}_clocks.csv &
SMI=$!
# 0. the GPU suite on this state (zero skips expected)
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -s -rs > gpurun_out/${TAG}_gpu_tests.txt 2>&1
tail -2 gpurun_out/${TAG}_gpu_tests.txt
# 1. the bench lines themselves (not under a profiler)
timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --steps 100 --warmup 5 --sync-steps --no-cpu-baseline > gpurun_out/${TAG}_bench_sync.json 2>> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --workload ba_stress --problems 64 --steps 3 --warmup 1 > gpurun_out/${TAG}_ba_stress.json 2>> gpurun_out/${TAG}_bench.err
kill $SMI
# 2. launch list of the repo's kernels in the tracking region (shares of the step, cold-cache and serialised)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name "$KREGEX" --launch-skip 62000 -c 2500 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 60 --warmup 3 --no-cpu-baseline --sync-steps > /tmp/under_ncu1.log 2>&1
gzip -f gpurun_out/${TAG}_launches.csv
# 3. ncu --set full of ~90 consecutive steady-state launches: every kernel of the streaming path (raw page as csv, the report stays here)
timeout 1500 ncu --set full --clock-control none --kernel-name "$KREGEX" --launch-skip 66000 -c 90 -o /tmp/${TAG}_full \
    python bench.py --steps 60 --warmup 3 --no-cpu-baseline --sync-steps > /tmp/under_ncu2.log 2>&1
ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/${TAG}_full_raw.csv.gz
# 4. source-level capture of one steady-state window-BA launch of the solve kernel (kept: the source page names the stalls per line)
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:solve_kernel --launch-skip 7000 -c 2 -o gpurun_out/${TAG}_solve_src \
    python bench.py --steps 60 --warmup 3 --no-cpu-baseline --sync-steps > /tmp/under_ncu3.log 2>&1
ncu -i gpurun_out/${TAG}_solve_src.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/${TAG}_solve_src_raw.csv.gz
# 5. the large-window solve (config 5)
timeout 900 ncu --set full --clock-control none --kernel-name regex:solve_large_kernel -c 1 -o /tmp/${TAG}_large \
    python bench.py --workload ba_stress --problems 64 --steps 1 --warmup 0 > /tmp/under_ncu4.log 2>&1
ncu -i /tmp/${TAG}_large.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/${TAG}_large_raw.csv.gz
ls -la gpurun_out | head -30; du -sh gpurun_out
cat gpurun_out/${TAG}_bench.json | cut -c1-2500
