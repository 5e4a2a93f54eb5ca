#!/bin/bash
# Runs every GPU test id separately under a hard timeout so a hung kernel cannot eat the budget.
# Usage (on the GPU box): bash tools/gpu_tests.sh [pytest -k expression]
mkdir -p gpurun_out
LOG=gpurun_out/gpu_tests.log
: > $LOG
nvidia-smi --query-gpu=name,memory.total --format=csv >> $LOG 2>&1
ids=$(python -m pytest tests -m gpu --collect-only -q ${1:+-k "$1"} 2>/dev/null | grep '::')
for t in $ids; do
  echo "=== $t" >> $LOG
  timeout -k 5 ${PER_TEST_TIMEOUT:-120} python -m pytest "$t" -x -q -m gpu 2>&1 | tail -25 >> $LOG
  echo "rc=$?" >> $LOG
done
grep -E "^=== |passed|failed|error|rc=" $LOG | tail -80
