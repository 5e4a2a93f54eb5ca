#!/bin/bash
mkdir -p gpurun_out
XRVIO_KPROF=1 python xrslam_b200/build.py --force > /dev/null 2>&1
timeout 600 python tools/phase_probe.py 64 > gpurun_out/c9_phase_probe.txt 2>&1
python xrslam_b200/build.py --force > /dev/null 2>&1
cat gpurun_out/c9_phase_probe.txt
XRVIO_LANES=4 timeout 300 python tools/phase_probe.py 32 2>&1 | grep -E "^steps|lane-step|slowest"
timeout 400 python bench.py --steps 60 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','lane_busy_ms','stepping')}, d['e2e']['value'], d['roofline']['family_ms'])"
