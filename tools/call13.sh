#!/bin/bash
mkdir -p gpurun_out
for mode in "" "--sync-steps"; do
timeout 400 python bench.py --steps 60 --warmup 3 --no-cpu-baseline $mode 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('$mode', {k:d[k] for k in ('value','ms_per_step','lane_busy_ms','phase_ms')})"
done > gpurun_out/c13_bench.txt 2>&1
cat gpurun_out/c13_bench.txt
