#!/bin/bash
# Round-2 GPU call 6: split-row attention (two threads per query row): tests, attention bench, whole-step bench.
mkdir -p gpurun_out/c6
O=gpurun_out/c6
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k attention > $O/pytest_attn.log 2>&1; echo "pytest attn rc=$?" | tee -a $O/summary.txt
timeout 200 python tools/attn_bench.py > $O/attn_bench.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/summary.txt
tail -6 $O/pytest_attn.log | cut -c1-300; cat $O/attn_bench.txt; tail -4 $O/pytest.log | cut -c1-300
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/c6/bench.json").read())
    print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "c1 ms", round(d["latency_c1"]["ms_median"], 2), "c3", round(d["c3"]["value"], 1),
          "gemm", round(d["roofline"]["achieved"]), "attn", round(d["roofline"]["secondary"]["achieved"]), d["roofline"]["secondary"]["avg_launch_ms"], d["clocks"])
except Exception as e:
    print("bench parse failed", e)
PY
tail -5 $O/bench.err
