#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/tma_probe tools/tma_probe.cu && timeout 60 /tmp/tma_probe > gpurun_out/c4_tma_probe.txt 2>&1
cat gpurun_out/c4_tma_probe.txt
export XRVIO_KLT_TMA=0
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_zz_large_window.py tests/test_gpu_pipeline.py -x -q -m gpu -p no:cacheprovider > gpurun_out/c4_tests.txt 2>&1
tail -15 gpurun_out/c4_tests.txt
XRVIO_KPROF=1 python xrslam_b200/build.py --force > /dev/null 2>&1
timeout 600 python tools/phase_probe.py 64 > gpurun_out/c4_phase_probe.txt 2>&1
python xrslam_b200/build.py --force > /dev/null 2>&1
cat gpurun_out/c4_phase_probe.txt
bash tools/lane_sweep.sh
