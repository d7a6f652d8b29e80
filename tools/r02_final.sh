#!/bin/bash
# Round-2 final validation: what the driver runs at round end (GPU tests, smoke, bench N = 1, reference arm), then soak with what is left.
mkdir -p gpurun_out/final
O=gpurun_out/final
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/summary.txt
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference rc=$?" | tee -a $O/summary.txt
N=${1:-6}
for i in $(seq 1 $N); do
  timeout 320 python tools/soak.py --seconds 150 --tag "final2.$i" >> $O/soak.jsonl 2>> $O/soak.err
  echo "soak final2.$i rc=$? $(nvidia-smi --query-gpu=temperature.gpu,power.draw,clocks.sm --format=csv,noheader)" | tee -a $O/summary.txt
done
tail -3 $O/pytest.log | cut -c1-200; tail -4 $O/smoke.log; cut -c1-400 $O/bench.json; cut -c1-300 $O/soak.jsonl
