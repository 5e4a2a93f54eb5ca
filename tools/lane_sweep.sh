#!/bin/bash
# Throughput of the streaming bench against the lane split (sessions per lane) and the host threads per lane.
#   gpurun --timeout 1200 -- 'bash tools/lane_sweep.sh'
mkdir -p gpurun_out
: > gpurun_out/lane_sweep.txt
for cfg in "8 64" "8 8" "16 16" "16 64" "32 32" "64 64"; do
  set -- $cfg
  echo "== XRVIO_LANES=$1 XRVIO_HOST_THREADS=$2" >> gpurun_out/lane_sweep.txt
  XRVIO_LANES=$1 XRVIO_HOST_THREADS=$2 timeout 400 python bench.py --steps 60 --warmup 3 --no-cpu-baseline 2>&1 | \
    python -c "import sys,json; [print({k:(round(v,1) if isinstance(v,float) else v) for k,v in json.loads(l).items() if k in ('value','ms_per_step','gpu_launches')}, json.loads(l)['e2e']['value'], json.loads(l)['roofline']['family_ms']) for l in sys.stdin if l.startswith('{')]" >> gpurun_out/lane_sweep.txt 2>&1
done
cat gpurun_out/lane_sweep.txt
