#!/bin/bash
mkdir -p gpurun_out
XRVIO_SLOW_STEP_MS=9 timeout 600 python tools/phase_probe.py 64 > gpurun_out/c10_phase_probe.txt 2> gpurun_out/c10_slow.txt
grep -E "^steps|lane-step|slowest" gpurun_out/c10_phase_probe.txt; head -40 gpurun_out/c10_slow.txt | cut -c1-400
