#!/bin/bash
# Round-2 GPU call 13: LayerNorm folded into the residual GEMMs: tests, then whole-step A/B on one box (F5_FUSE_LN=0/1).
mkdir -p gpurun_out/c13
O=gpurun_out/c13
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "folded or gemm" > $O/pytest_k.log 2>&1; echo "pytest kernels rc=$?" | tee -a $O/summary.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
for f in 0 1 0 1; do
  F5_FUSE_LN=$f timeout 400 python bench.py --steps 3 --warmup 2 --no-extras --no-cpu-baseline > $O/bench_f$f.json 2> $O/bench_f$f.err
  python - "$f" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/c13/bench_f{f}.json").read())
    print(f"fuse_ln={f}: value {d['value']:.1f} e2e {d['e2e']['value']:.1f} ms/step {d['ms_per_step']:.1f} gemm {d['roofline']['achieved']:.0f} ({d['roofline']['avg_launch_ms']:.3f} ms) attn {d['roofline']['secondary']['achieved']:.0f} clk {d['clocks']['sm_mhz']}")
except Exception as e:
    print(f"fuse_ln={f} failed: {e}")
PY
done | tee -a $O/ab.txt
tail -3 $O/pytest_k.log | cut -c1-250; tail -3 $O/pytest.log | cut -c1-250; tail -3 $O/bench_f1.err | cut -c1-250
