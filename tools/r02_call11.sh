#!/bin/bash
# Round-2 GPU call 11: new tests (trajectory, fp32 at C2 length) and an in-step scan of attention build variants (whole-step A/B on one box).
mkdir -p gpurun_out/c11
O=gpurun_out/c11
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
for s in "" _p0 _p2 _g1 _k4 ""; do
  F5_LIB_SUFFIX=$s timeout 400 python bench.py --steps 3 --warmup 2 --no-extras --no-cpu-baseline > $O/bench$s.json 2> $O/bench$s.err
  python - "$s" <<'PY'
import json, sys
s = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/c11/bench{s}.json").read())
    print(f"variant '{s}': value {d['value']:.1f} e2e {d['e2e']['value']:.1f} gemm {d['roofline']['achieved']:.0f} attn {d['roofline']['secondary']['achieved']:.0f} "
          f"attn_ms {d['roofline']['secondary']['avg_launch_ms']:.3f} clk {d['clocks']['sm_mhz']}")
except Exception as e:
    print(f"variant '{s}' failed: {e}")
PY
done | tee -a $O/variants.txt
tail -3 $O/pytest.log | cut -c1-200
