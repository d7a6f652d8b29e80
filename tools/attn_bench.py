"""Attention-only benchmark at C2 / C3 scale: median and min of 30 launches (CUDA events)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tts_indic_server_f5_b200 import ops
from tts_indic_server_f5_b200.layout import build_layout
dev = "cuda"; torch.manual_seed(0)
g = torch.Generator().manual_seed(0)
cases = {"c2": [469 + int(torch.randint(560, 941, (1,), generator=g)) for _ in range(64)], "c3": [3069] * 16}
D = 1024
for name, lens in cases.items():
    L = build_layout(lens)
    qkv = torch.randn(L.rows, 3 * D, device=dev).to(torch.bfloat16)
    ab = torch.zeros(L.rows, D, device=dev, dtype=torch.bfloat16)
    tiles = L.attn_tiles.to(dev)
    for _ in range(5):
        ops.attention(qkv, tiles, ab, 16, 0, D, 2 * D, 0.125)
    n = 30
    e = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    e[0].record()
    for i in range(n):
        ops.attention(qkv, tiles, ab, 16, 0, D, 2 * D, 0.125)
        e[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(e[i].elapsed_time(e[i + 1]) for i in range(n))
    fl = 2 * sum(4.0 * D * x * x for x in lens)
    print(f"{os.environ.get('F5_LIB_SUFFIX','')} attention {name}: median {ts[n//2]*1e3:7.1f} us  min {ts[0]*1e3:7.1f} us  -> {fl/ts[n//2]/1e9:6.1f} TFLOP/s (algorithmic)", flush=True)
