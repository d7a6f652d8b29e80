#!/bin/bash
# Round-2 GPU call 3: tests with PDL, B = 1 latency and the C2 step with PDL off / on, soak of the shipped library.
mkdir -p gpurun_out/c3
O=gpurun_out/c3
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
F5_PDL=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_pdl0.json 2> $O/bench_pdl0.err; echo "bench pdl0 rc=$?" | tee -a $O/summary.txt
F5_PDL=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_pdl1.json 2> $O/bench_pdl1.err; echo "bench pdl1 rc=$?" | tee -a $O/summary.txt
timeout 200 python tools/attn_bench.py >> $O/attn_bench.txt 2>&1
for i in 1 2 3; do
  timeout 300 python tools/soak.py --seconds 150 --tag "ship.$i" >> $O/soak.jsonl 2>> $O/soak.err
  echo "soak ship.$i rc=$? $(nvidia-smi --query-gpu=temperature.gpu,power.draw,clocks.sm --format=csv,noheader)" | tee -a $O/summary.txt
done
tail -12 $O/pytest.log | cut -c1-300
python - <<'PY'
import json
for n in ("pdl0", "pdl1"):
    try:
        d = json.loads(open(f"gpurun_out/c3/bench_{n}.json").read())
        print(n, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "c1 ms", round(d["latency_c1"]["ms_median"], 2), "c3", round(d["c3"]["value"], 1),
              "gemm", round(d["roofline"]["achieved"]), "attn", round(d["roofline"]["secondary"]["achieved"]), d["clocks"])
    except Exception as e:
        print(n, "failed", e)
PY
cat $O/attn_bench.txt; cut -c1-300 $O/soak.jsonl; tail -3 $O/bench_pdl1.err
