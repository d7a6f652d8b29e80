#!/bin/bash
# Long soak of the shipped library (tools/soak.py: the bench's own call sequence at C2 size, one process per ~150 s).
N=${1:-18}
mkdir -p gpurun_out/soak_long
O=gpurun_out/soak_long
for i in $(seq 1 $N); do
  timeout 320 python tools/soak.py --seconds 150 --tag "final.$i" >> $O/soak.jsonl 2>> $O/soak.err
  rc=$?
  echo "soak final.$i rc=$rc $(nvidia-smi --query-gpu=temperature.gpu,power.draw,clocks.sm --format=csv,noheader)" | tee -a $O/summary.txt
  if [ $rc -ne 0 ]; then
    nvidia-smi -q -d PAGE_RETIREMENT,ECC > $O/fail_${i}_smi.txt 2>&1
    dmesg 2>/dev/null | grep -i -E "xid|nvrm" | tail -20 > $O/fail_${i}_xid.txt
    sleep 5
  fi
done
python - <<'PY'
import json
rows = [json.loads(l) for l in open("gpurun_out/soak_long/soak.jsonl")]
print("processes", len(rows), "cycles", sum(r["cycles"] for r in rows), "launches", sum(r["launches"] for r in rows), "failures", sum(r["failed"] for r in rows))
for r in rows:
    if r["failed"]:
        print(r)
PY
