"""Same-box A/B of the kernel variants of the two Vocos-side memory kernels (f5_set_dwconv7_variant / f5_set_istft_variant),
one process, CUDA events, median of 9 after 3 warm-ups:
  * f5_dwconv7_ln alone, C = 512, 262 k rows (64 utterances of 4096 frames): algorithmic bytes = 2 KB fp32 in + 1 KB bf16 out per row
  * f5_istft_frames alone, 262 k frames: 4104 B spectrum in + 4096 B windowed frame out per frame
  * the whole Vocos decode at 64 x 2048 and 64 x 4096 frames for (dwconv, istft) = (1, 1) [round-1 kernels], (2, 2), (3, 2)
  python tools/vocos_kernels_ab.py [out.json]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tts_indic_server_f5_b200 import _lib, api, ops, weights as W  # noqa: E402

dev = torch.device("cuda")
HBM = 6551.0
try:
    HBM = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timed(fn, reps=9, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


out = {"hbm_peak_gbs": HBM, "dwconv7_ln": {}, "istft_frames": {}, "decode": {}}
voc = api.load_vocoder(state_dict=W.make_vocos_state_dict(W.VOCOS_24K, seed=0), device=dev)
eng = voc.engine
T, B = 4096, 64
starts, Rv, pos, offs, tot = eng.plan([T] * B)
pos = pos.to(dev)
g = torch.Generator("cpu").manual_seed(0)
x = torch.randn(Rv, 512, generator=g).to(dev)
y = torch.zeros(Rv, 512, device=dev, dtype=torch.bfloat16)
blk = eng.blocks[0]
ref = None
for v in (1, 2, 3):
    old = _lib.lib.f5_set_dwconv7_variant(v)
    ms = timed(lambda: ops.dwconv7_ln(x, y, pos, blk["dw_w"], blk["dw_b"], blk["ln_w"], blk["ln_b"]))
    _lib.lib.f5_set_dwconv7_variant(old)
    gbs = Rv * 3072 / ms / 1e6
    same = None
    if ref is None:
        ref = y.clone()
    else:
        same = float((y.float() - ref.float()).abs().max())
    out["dwconv7_ln"][v] = {"rows": Rv, "ms": ms, "gbs": gbs, "frac_hbm": gbs / HBM, "max_abs_vs_v1": same}
    print(f"dwconv7_ln v{v}: {Rv} rows x 512  {ms * 1e3:8.1f} us  {gbs:7.0f} GB/s = {100 * gbs / HBM:4.1f} % of HBM peak  (max |y - y_v1| = {same})", flush=True)
del x, y, ref

spec = torch.randn(Rv, 1152, generator=g).to(dev)
frames = torch.zeros(Rv, 1024, device=dev)
ref = None
for v in (1, 2):
    old = _lib.lib.f5_set_istft_variant(v)
    ms = timed(lambda: ops.call("f5_istft_frames", ops.ptr(spec), spec.stride(0), Rv, ops.ptr(eng.window), ops.ptr(frames), ops.stream_ptr()))
    _lib.lib.f5_set_istft_variant(old)
    gbs = Rv * (4104 + 4096) / ms / 1e6
    same = None
    if ref is None:
        ref = frames.clone()
    else:
        same = float((frames - ref).norm() / ref.norm())
    out["istft_frames"][v] = {"rows": Rv, "ms": ms, "gbs": gbs, "frac_hbm": gbs / HBM, "rel_l2_vs_v1": same}
    print(f"istft_frames v{v}: {Rv} frames  {ms * 1e3:8.1f} us  {gbs:7.0f} GB/s = {100 * gbs / HBM:4.1f} % of HBM peak  (rel-L2 vs v1 = {same})", flush=True)
del spec, frames, ref
torch.cuda.empty_cache()

for (T, B) in ((2048, 64), (4096, 64)):
    mel = (torch.randn(B, 100, T, generator=g) * 2 - 4).to(dev)
    for dv, iv in ((1, 1), (2, 2), (3, 2)):
        o1, o2 = _lib.lib.f5_set_dwconv7_variant(dv), _lib.lib.f5_set_istft_variant(iv)
        ms = timed(lambda: voc.decode(mel), reps=5, warm=2)
        _lib.lib.f5_set_dwconv7_variant(o1)
        _lib.lib.f5_set_istft_variant(o2)
        out["decode"][f"{B}x{T} dwconv{dv} istft{iv}"] = {"ms": ms, "mframes_per_s": B * T / ms / 1e3}
        print(f"decode {B} x {T} frames, dwconv v{dv}, istft v{iv}: {ms:7.3f} ms = {B * T / ms / 1e3:6.2f} Mframe/s", flush=True)
    del mel
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
