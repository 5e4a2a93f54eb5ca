#!/bin/bash
mkdir -p gpurun_out
for cfg in "8 1" "16 2" "32 4" "64 8"; do
  set -- $cfg
  echo "== sessions $1 lanes $2"
  XRVIO_LANES=$2 timeout 300 python tools/phase_probe.py $1 2>&1 | grep -E "^steps|31  " 
done > gpurun_out/c8_scaling.txt 2>&1
cat gpurun_out/c8_scaling.txt
