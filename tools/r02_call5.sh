#!/bin/bash
# Round-2 GPU call 5: out-of-line warp waits + small-M tile rule: tests, attention bench, C1 latency A/B, bench, soak.
mkdir -p gpurun_out/c5
O=gpurun_out/c5
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
timeout 200 python tools/attn_bench.py > $O/attn_bench.txt 2>&1
timeout 200 python tools/gemm_bench.py > $O/gemm_bench.txt 2>&1
F5_SMALL_M_RULE=0 timeout 300 python tools/latency_c1.py >> $O/latency.txt 2>&1
timeout 300 python tools/latency_c1.py >> $O/latency.txt 2>&1
F5_PDL=0 timeout 300 python tools/latency_c1.py >> $O/latency.txt 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/summary.txt
for i in 1 2 3 4; do
  timeout 300 python tools/soak.py --seconds 150 --tag "ship2.$i" >> $O/soak.jsonl 2>> $O/soak.err
  echo "soak ship2.$i rc=$? $(nvidia-smi --query-gpu=temperature.gpu,power.draw,clocks.sm --format=csv,noheader)" | tee -a $O/summary.txt
done
tail -4 $O/pytest.log | cut -c1-300; cat $O/attn_bench.txt; head -14 $O/gemm_bench.txt; cat $O/latency.txt
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c5/bench.json").read())
print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "c1 ms", round(d["latency_c1"]["ms_median"], 2), "c3", round(d["c3"]["value"], 1),
      "gemm", round(d["roofline"]["achieved"]), "attn", round(d["roofline"]["secondary"]["achieved"]), d["clocks"])
PY
cut -c1-200 $O/soak.jsonl
