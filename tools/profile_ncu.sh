#!/bin/bash
# ncu captures of every kernel of the streaming path and of the large-window solve (run on the GPU box, after tools/profile_final.sh):
#   gpurun --timeout 1500 -- 'bash tools/profile_ncu.sh r2'
# To stay inside the GPU budget the per-kernel capture collects the metrics the summaries use (listed below: ~8 replay passes per
# launch instead of the ~45 of --set full) over 200 consecutive launches of a shortened run (100 priming frames instead of 190:
# same kernels, slightly younger maps); the dominant kernel gets --set full with sources, and so does solve_large_kernel.
TAG=${1:-r2}
KREGEX='regex:(clahe|pyrdown|scharr|harris|corner_select|klt_|poisson|preint|solve_|marg_|tri_kernel|undistort|gram_)'
ST='smsp__average_warps_issue_stalled'
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct"
M="$M,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active"
M="$M,sm__warps_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active"
M="$M,launch__registers_per_thread,launch__grid_size,launch__block_size,launch__shared_mem_per_block_dynamic,smsp__sass_inst_executed_op_local_ld.sum,smsp__sass_inst_executed_op_local_st.sum"
for s in barrier long_scoreboard short_scoreboard wait math_pipe_throttle mio_throttle lg_throttle not_selected dispatch_stall branch_resolving no_instruction membar sleeping drain misc tex_throttle selected; do
  M="$M,${ST}_${s}_per_issue_active.ratio"
done
mkdir -p gpurun_out
export XRVIO_BENCH_PRIME=100
timeout 600 ncu --metrics "$M" --clock-control none --kernel-name "$KREGEX" --launch-skip 9000 -c 400 -o /tmp/${TAG}_full \
    python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sync-steps --dev-only > /tmp/under_ncu2.log 2>&1
ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/${TAG}_full_raw.csv.gz
timeout 420 ncu --set full --clock-control none --import-source on --kernel-name regex:solve_kernel --launch-skip 1400 -c 2 -o gpurun_out/${TAG}_solve_src \
    python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sync-steps --dev-only > /tmp/under_ncu3.log 2>&1
ncu -i gpurun_out/${TAG}_solve_src.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/${TAG}_solve_src_raw.csv.gz
timeout 300 ncu --set full --clock-control none --kernel-name regex:solve_large_kernel -c 1 -o /tmp/${TAG}_large \
    python bench.py --workload ba_stress --problems 64 --steps 1 --warmup 0 > /tmp/under_ncu4.log 2>&1
ncu -i /tmp/${TAG}_large.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/${TAG}_large_raw.csv.gz
ls -la gpurun_out | grep ${TAG}_; du -sh gpurun_out; tail -2 /tmp/under_ncu2.log | cut -c1-300
