"""Launch one instance of each hot kernel at C2 scale (for `ncu -k regex:... -c N`): QKV GEMM, FF1+GELU GEMM,
out-proj residual GEMM, attention."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tts_indic_server_f5_b200 import ops  # noqa: E402
from tts_indic_server_f5_b200.layout import build_layout  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
g = torch.Generator().manual_seed(0)
lens = [469 + int(torch.randint(560, 941, (1,), generator=g)) for _ in range(64)]
L = build_layout(lens)
M, D = L.rows, 1024
A = torch.randn(M, D, device=dev).to(torch.bfloat16)
bias3, bias2, bias1 = torch.randn(3 * D, device=dev), torch.randn(2 * D, device=dev), torch.randn(D, device=dev)
Wqkv = (torch.randn(3 * D, D, device=dev) / 32).to(torch.bfloat16)
W1 = (torch.randn(2 * D, D, device=dev) / 32).to(torch.bfloat16)
Wo = (torch.randn(D, D, device=dev) / 32).to(torch.bfloat16)
qkv = torch.zeros(M, 3 * D, device=dev, dtype=torch.bfloat16)
fb = torch.zeros(M, 2 * D, device=dev, dtype=torch.bfloat16)
ab = torch.zeros(M, D, device=dev, dtype=torch.bfloat16)
xres = torch.randn(M, D, device=dev)
gate = torch.randn(D, device=dev)
tiles = L.attn_tiles.to(dev)
W2 = (torch.randn(D, 2 * D, device=dev) / 45).to(torch.bfloat16)
scale, shift = torch.randn(D, device=dev) * 0.1, torch.randn(D, device=dev) * 0.1
hb = torch.zeros(M, D, device=dev, dtype=torch.bfloat16)
for rep in range(1):           # launch order per repetition: QKV, attention, out-proj, FF1, FF2, LayerNorm (one DiT layer)
    ops.gemm(A, Wqkv, mode=ops.F5_EPI_STORE_BF16, bias=bias3, out=qkv)
    ops.attention(qkv, tiles, ab, 16, 0, D, 2 * D, 0.125)
    ops.gemm(ab, Wo, mode=ops.F5_EPI_RESID_F32, bias=bias1, gate=gate, resid=xres)
    ops.gemm(A, W1, mode=ops.F5_EPI_STORE_BF16, act=ops.F5_ACT_GELU_TANH, bias=bias2, out=fb)
    ops.gemm(fb, W2, mode=ops.F5_EPI_RESID_F32, bias=bias1, gate=gate, resid=xres)
    ops.layernorm_mod(xres, hb, scale, shift, 1.0)
torch.cuda.synchronize()
print("ok rows", M, "tiles", tiles.shape[0])
