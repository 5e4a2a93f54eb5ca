#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/tma_probe tools/tma_probe.cu -ldl 2> gpurun_out/c5_tma_probe.txt
for v in 0 1 2 3 4 5 6 7 8 9; do timeout 30 /tmp/tma_probe $v >> gpurun_out/c5_tma_probe.txt 2>&1; done
cat gpurun_out/c5_tma_probe.txt
export XRVIO_KLT_TMA=0
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_pipeline.py -x -q -m gpu -p no:cacheprovider -s > gpurun_out/c5_tests.txt 2>&1
grep -E "config 2|prior form|passed|failed|Error|assert" gpurun_out/c5_tests.txt | head -20
XRVIO_KPROF=1 python xrslam_b200/build.py --force > /dev/null 2>&1
timeout 600 python tools/phase_probe.py 64 > gpurun_out/c5_phase_probe.txt 2>&1
python xrslam_b200/build.py --force > /dev/null 2>&1
cat gpurun_out/c5_phase_probe.txt
for cfg in "8 64" "4 64" "2 64"; do
  set -- $cfg
  echo "== XRVIO_LANES=$1 XRVIO_HOST_THREADS=$2"
  XRVIO_LANES=$1 XRVIO_HOST_THREADS=$2 timeout 400 python bench.py --steps 60 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-1500
done > gpurun_out/c5_lanes.txt 2>&1
cat gpurun_out/c5_lanes.txt
