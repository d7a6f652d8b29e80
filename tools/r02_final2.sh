#!/bin/bash
# Round-2 GPU call 16 (the last 3.8 GPU-minutes): final library (dwconv7_ln default = form 2) — full GPU suite, the bench line
# with its secondary blocks (incl. memory_kernels), then one ncu --set full capture of the Vocos-side memory kernels.
mkdir -p gpurun_out/c16
O=gpurun_out/c16
timeout 120 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
timeout 150 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/summary.txt
timeout 40 python tools/ncu_vocos_mem.py > $O/ncu_plain.log 2>&1; echo "plain rc=$?" | tee -a $O/summary.txt
timeout 100 ncu --set full --clock-control none --import-source on -f -o $O/prof_r02_vocos_mem -k regex:"dwconv7_ln|istft" -c 3 \
    python tools/ncu_vocos_mem.py > $O/ncu.log 2>&1; echo "ncu rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest.log | cut -c1-300; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/c16/bench.json").read())
    print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d.get("latency_c1", {}).get("ms_median"), d.get("c3", {}).get("value"), d.get("c5"), json.dumps(d.get("memory_kernels"))[:900])
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 $O/bench.err | cut -c1-300; tail -2 $O/ncu_plain.log; tail -4 $O/ncu.log | cut -c1-200; ls -la $O
