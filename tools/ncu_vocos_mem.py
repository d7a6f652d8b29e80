"""Launch the Vocos-side memory kernels once each at C5 scale (32 x 4096 frames) for `ncu -k regex:... -c N`:
f5_dwconv7_ln (default variant), f5_istft_frames (default variant), f5_istft_ola, and the round-1 forms after them."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tts_indic_server_f5_b200 import _lib, ops  # noqa: E402
from tts_indic_server_f5_b200.vocos import VocosEngine  # noqa: E402

dev = torch.device("cuda")
T, B, C = 4096, 32, 512
starts, Rv, pos, offs, tot = VocosEngine.plan([T] * B)
pos = pos.to(dev)
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(Rv, C, generator=g, device=dev)
y = torch.empty(Rv, C, device=dev, dtype=torch.bfloat16)
w, b = torch.randn(C, 7, generator=g, device=dev) * 0.4, torch.randn(C, generator=g, device=dev)
lw, lb = torch.randn(C, generator=g, device=dev) + 1, torch.randn(C, generator=g, device=dev)
spec = torch.randn(Rv, 1152, generator=g, device=dev)
frames = torch.empty(Rv, 1024, device=dev)
window = torch.hann_window(1024, device=dev)
seg = torch.tensor([[s, T, o, 0] for s, o in zip(starts, offs)], dtype=torch.int32).to(dev)
wav = torch.empty(tot, device=dev)
torch.cuda.synchronize()
for dv, iv in ((0, 0), (1, 1)):                      # (0, 0): out of range = keep the defaults; then the round-1 forms (ncu -c 3 stops before them)
    o1, o2 = _lib.lib.f5_set_dwconv7_variant(dv), _lib.lib.f5_set_istft_variant(iv)
    ops.dwconv7_ln(x, y, pos, w, b, lw, lb)
    ops.istft(spec, window, frames, seg, 256 * (T - 1), wav)
    torch.cuda.synchronize()
    _lib.lib.f5_set_dwconv7_variant(o1)
    _lib.lib.f5_set_istft_variant(o2)
print("ok rows", Rv, "finite", bool(torch.isfinite(wav).all()))
