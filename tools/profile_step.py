"""Per-op device-time breakdown of the hot path on one workload (CUDA events around every C-ABI launch, eager mode).
  python tools/profile_step.py [workload] [euler_steps]"""
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tts_indic_server_f5_b200 import _lib, api, ops, synthetic as S, weights as W  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
model = api.load_model(state_dict=W.make_dit_state_dict(W.INDICF5, seed=0))
voc = api.load_vocoder(state_dict=W.make_vocos_state_dict(W.VOCOS_24K, seed=0))
syn = api.Synthesizer(model, voc)
specs = S.workload(wl)
noise = [S.initial_noise(4096, s.noise_index) for s in specs]
model.engine.use_graphs = False
st = syn.stage(specs, nfe_step=nsteps, y0=noise)
syn.run(st)                                  # warm-up
torch.cuda.synchronize()

recs = []
orig_call, orig_gemm = _lib.call, ops.gemm
tag = {"t": ""}


def timed_call(name, *args):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    orig_call(name, *args)
    b.record()
    recs.append((name + tag["t"], a, b))


def tagged_gemm(A, B, **kw):
    N = kw.get("N") or (B.shape[0] if kw.get("num_taps", 1) == 1 else kw.get("b_tap_rows"))
    tag["t"] = f" M={A.shape[0]} N={N} K={B.shape[1]}x{kw.get('num_taps', 1)} mode={kw['mode']} act={kw.get('act', 0)}"
    try:
        orig_gemm(A, B, **kw)
    finally:
        tag["t"] = ""


ops.call = timed_call
ops.gemm = tagged_gemm
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
syn.run(st)
e1.record()
torch.cuda.synchronize()
tot = collections.OrderedDict()
for name, a, b in recs:
    t, c = tot.get(name, (0.0, 0))
    tot[name] = (t + a.elapsed_time(b), c + 1)
total = e0.elapsed_time(e1)
print(f"workload {wl}: rows={2 * st.ws.R} real_tokens={st.layout.real_tokens} euler_steps={nsteps}  total {total:.1f} ms (eager)")
for name, (t, c) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print(f"{t:10.2f} ms  {100 * t / total:5.1f}%  x{c:5d}  avg {t / c * 1e3:9.1f} us  {name}")
