#!/bin/bash
mkdir -p gpurun_out
for L in 16 10 6 4; do
XRVIO_LANES=$L timeout 400 python bench.py --steps 60 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('lanes $L', {k:d[k] for k in ('value','ms_per_step','lane_busy_ms')}, d['e2e']['value'])"
done > gpurun_out/c14_bench.txt 2>&1
cat gpurun_out/c14_bench.txt
