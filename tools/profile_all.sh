#!/bin/bash
# One gpurun call that refreshes every measurement the roofline lines are built from (run on the GPU box):
#   gpurun --timeout 1500 -- 'bash tools/profile_all.sh [tag]'
# Outputs under gpurun_out/<tag>_*: phase_probe.txt (host wall time per step phase + clock64 sections of the window BA),
# launches.csv (ncu launch list of the repo's kernels in the tracking region of the bench command), full.ncu-rep + full_raw.csv
# (ncu --set full of ~160 consecutive steady-state launches: every kernel of the streaming path).
TAG=${1:-cur}
KREGEX='regex:(clahe|pyrdown|scharr|harris|corner_select|klt_|poisson|preint|solve_|marg_|tri_kernel|undistort|gram_|essential|rot_ransac)'
mkdir -p gpurun_out
XRVIO_KPROF=1 python xrslam_b200/build.py --force > /dev/null 2>&1          # diagnostic build with the clock64 sections
timeout 600 python tools/phase_probe.py 64 > gpurun_out/${TAG}_phase_probe.txt 2>&1
python xrslam_b200/build.py --force > /dev/null 2>&1                         # back to the product build
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name "$KREGEX" --launch-skip 4000 -c 3000 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name "$KREGEX" --launch-skip 5500 -c 160 -o gpurun_out/${TAG}_full \
    python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_full.log 2>&1
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
tail -40 gpurun_out/${TAG}_phase_probe.txt
