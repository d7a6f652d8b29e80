#!/bin/bash
# Round-2 GPU call 1: tests on the fixed protocol, what the watchdog record costs, and a same-box A/B soak of the
# round-1 wait protocol (A: every lane of the producer / MMA warps polls) against the fixed one (B: one poller).
mkdir -p gpurun_out/c1
O=gpurun_out/c1
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
for s in "" _nd; do
  F5_LIB_SUFFIX=$s timeout 200 python tools/attn_bench.py >> $O/attn_bench.txt 2>&1
done
F5_LIB_SUFFIX= timeout 200 python tools/gemm_bench.py > $O/gemm_bench.txt 2>&1
F5_LIB_SUFFIX=_nd timeout 200 python tools/gemm_bench.py > $O/gemm_bench_nd.txt 2>&1
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/summary.txt
SOAK_S=${SOAK_S:-140}
ROUNDS=${ROUNDS:-4}
for i in $(seq 1 $ROUNDS); do
  for s in _A _B; do
    F5_LIB_SUFFIX=$s timeout $((SOAK_S + 120)) python tools/soak.py --seconds $SOAK_S --tag "$s.$i" >> $O/soak.jsonl 2>> $O/soak.err
    rc=$?
    echo "soak $s.$i rc=$rc $(nvidia-smi --query-gpu=temperature.gpu,power.draw,clocks.sm --format=csv,noheader)" | tee -a $O/summary.txt
    if [ $rc -ne 0 ]; then
      nvidia-smi -q -d PAGE_RETIREMENT,ECC > $O/fail_${s}_${i}_smi.txt 2>&1
      dmesg 2>/dev/null | grep -i -E "xid|nvrm" | tail -20 > $O/fail_${s}_${i}_xid.txt
      sleep 5
    fi
  done
done
cat $O/soak.jsonl | cut -c1-400
tail -5 $O/pytest.log; cat $O/attn_bench.txt; head -12 $O/gemm_bench.txt; head -12 $O/gemm_bench_nd.txt; cut -c1-600 $O/bench.json
