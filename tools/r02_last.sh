#!/bin/bash
# Round-2 last GPU call: GPU tests on HEAD, then soak with what is left of the budget.
mkdir -p gpurun_out/last
O=gpurun_out/last
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
for i in 1 2 3; do
  timeout 320 python tools/soak.py --seconds 150 --tag "last.$i" >> $O/soak.jsonl 2>> $O/soak.err
  echo "soak last.$i rc=$? $(nvidia-smi --query-gpu=temperature.gpu,power.draw,clocks.sm --format=csv,noheader)" | tee -a $O/summary.txt
done
tail -2 $O/pytest.log | cut -c1-200; cut -c1-200 $O/soak.jsonl
