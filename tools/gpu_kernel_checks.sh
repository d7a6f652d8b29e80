#!/bin/bash
# Run every kernel-level GPU test in its own process (a device trap in one kernel must not mask the others).
# Usage (on the GPU box): bash tools/gpu_kernel_checks.sh [pytest -k filter]
mkdir -p gpurun_out
LOG=gpurun_out/kernels.log
: > $LOG
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm --format=csv >> $LOG 2>&1
TESTS=$(python -m pytest tests/test_kernels_gpu.py -m gpu --collect-only -q ${1:+-k "$1"} 2>/dev/null | grep "::")
pass=0; fail=0
for t in $TESTS; do
  echo "=== $t" >> $LOG
  if timeout 180 python -m pytest "$t" -m gpu -x -q -s >> $LOG 2>&1; then pass=$((pass+1)); else fail=$((fail+1)); echo "FAILED: $t" >> $LOG; fi
done
echo "SUMMARY pass=$pass fail=$fail" | tee -a $LOG
grep -E "^\[|FAILED|SUMMARY|watchdog|Error|error" $LOG | tail -120
