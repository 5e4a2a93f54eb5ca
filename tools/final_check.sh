#!/bin/bash
# The round-end checks in one gpurun call: the whole GPU suite (zero skips expected), the smoke entry point, the default bench line.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -s -rs > gpurun_out/r2_gpu_tests.txt 2>&1
grep -E "config 2|prior form|passed|failed|skipped|SKIP|Error" gpurun_out/r2_gpu_tests.txt | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | cut -c1-3000
