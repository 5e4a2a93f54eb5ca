#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_zz_large_window.py tests/test_gpu_pipeline.py -x -q -m gpu -p no:cacheprovider > gpurun_out/c3_tests.txt 2>&1
tail -15 gpurun_out/c3_tests.txt
XRVIO_KPROF=1 python xrslam_b200/build.py --force > /dev/null 2>&1
timeout 600 python tools/phase_probe.py 64 > gpurun_out/c3_phase_probe.txt 2>&1
python xrslam_b200/build.py --force > /dev/null 2>&1
cat gpurun_out/c3_phase_probe.txt
timeout 300 python bench.py --workload ba_stress --problems 64 --steps 3 --warmup 1 > gpurun_out/c3_ba_stress.json 2> gpurun_out/c3_ba_stress.err; cat gpurun_out/c3_ba_stress.json; tail -3 gpurun_out/c3_ba_stress.err
bash tools/lane_sweep.sh
