"""Per-op device-time breakdown of one Vocos decode (CUDA events around every C-ABI launch).
  python tools/vocos_profile.py [T] [B]"""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tts_indic_server_f5_b200 import _lib, api, ops, weights as W  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
voc = api.load_vocoder(state_dict=W.make_vocos_state_dict(W.VOCOS_24K, seed=0))
mel = torch.randn(B, 100, T) * 2 - 4
voc.decode(mel)
torch.cuda.synchronize()
recs, orig_call, orig_gemm, tag = [], _lib.call, ops.gemm, {"t": ""}


def timed_call(name, *args):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    orig_call(name, *args)
    b.record()
    recs.append((name + tag["t"], a, b))


def tagged_gemm(A, Bm, **kw):
    N = kw.get("N") or (Bm.shape[0] if kw.get("num_taps", 1) == 1 else kw.get("b_tap_rows"))
    tag["t"] = f" N={N} K={Bm.shape[1]}x{kw.get('num_taps', 1)} mode={kw['mode']} act={kw.get('act', 0)}"
    try:
        orig_gemm(A, Bm, **kw)
    finally:
        tag["t"] = ""


ops.call, ops.gemm = timed_call, tagged_gemm
voc.decode(mel)
torch.cuda.synchronize()
tot = collections.OrderedDict()
for name, a, b in recs:
    t, c = tot.get(name, (0.0, 0))
    tot[name] = (t + a.elapsed_time(b), c + 1)
s = sum(t for t, _ in tot.values())
print(f"vocos decode T={T} B={B} ({T * B} frames): {s:.2f} ms in kernels")
for name, (t, c) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print(f"  {t:8.2f} ms {100 * t / s:5.1f}%  x{c:3d}  {name}")
