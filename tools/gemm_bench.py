"""Microbenchmark of the tcgen05 GEMM launcher on the DiT layer shapes (CUDA events, L2-exceeding operands)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tts_indic_server_f5_b200 import ops  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 158976
dev = "cuda"
torch.manual_seed(0)


def run(tag, N, K, mode, act, bn=256, iters=5, gate=False):
    A = torch.randn(M, K, device=dev).to(torch.bfloat16)
    B = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    g = torch.randn(N, device=dev) if gate else None
    kw = dict(mode=mode, act=act, bias=bias, block_n=bn)
    if mode == ops.F5_EPI_STORE_BF16:
        kw["out"] = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    elif mode == ops.F5_EPI_STORE_F32:
        kw["out"] = torch.zeros(M, N, device=dev)
    else:
        kw["resid"] = torch.zeros(M, N, device=dev)
        kw["gate"] = g
    for _ in range(2):
        ops.gemm(A, B, **kw)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    e[0].record()
    for i in range(iters):
        ops.gemm(A, B, **kw)
        e[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(e[i].elapsed_time(e[i + 1]) for i in range(iters))
    t = ts[len(ts) // 2]
    print(f"{tag:28s} N={N:5d} K={K:5d} mode={mode} act={act} bn={bn}: {t * 1e3:8.1f} us  {2.0 * M * N * K / t / 1e9:7.1f} TFLOP/s", flush=True)


run("qkv  store bf16", 3072, 1024, 0, 0)
run("ff1  store bf16 +gelu_tanh", 2048, 1024, 0, 1)
run("ff1  store bf16 no act", 2048, 1024, 0, 0)
run("ff1  store bf16 +gelu_erf", 2048, 1024, 0, 2)
run("ff1  store bf16 +mish", 2048, 1024, 0, 3)
run("N=3072 +gelu_tanh", 3072, 1024, 0, 1)
run("out  resid gate", 1024, 1024, 2, 0, gate=True)
run("ff2  resid gate", 1024, 2048, 2, 0, gate=True)
run("ff2  resid nogate", 1024, 2048, 2, 0)
run("N=1024 store bf16", 1024, 1024, 0, 0)
run("N=1024 store bf16 K2048", 1024, 2048, 0, 0)
run("N=1024 store f32", 1024, 1024, 1, 0)
run("qkv bn128", 3072, 1024, 0, 0, bn=128)
run("ff1 bn128 gelu", 2048, 1024, 0, 1, bn=128)

# attention at C2 scale
from tts_indic_server_f5_b200.layout import build_layout  # noqa: E402
g = torch.Generator().manual_seed(0)
lens = [469 + int(torch.randint(560, 941, (1,), generator=g)) for _ in range(64)]
L = build_layout(lens)
D = 1024
qkv = torch.randn(L.rows, 3 * D, device=dev).to(torch.bfloat16)
ab = torch.zeros(L.rows, D, device=dev, dtype=torch.bfloat16)
tiles = L.attn_tiles.to(dev)
for _ in range(2):
    ops.attention(qkv, tiles, ab, 16, 0, D, 2 * D, 0.125)
e = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
e[0].record()
for i in range(5):
    ops.attention(qkv, tiles, ab, 16, 0, D, 2 * D, 0.125)
    e[i + 1].record()
torch.cuda.synchronize()
t = sorted(e[i].elapsed_time(e[i + 1]) for i in range(5))[2]
fl = 2 * sum(4.0 * D * n * n for n in lens)
print(f"attention C2 (64 utts x2, 16 heads): {t * 1e3:8.1f} us  {fl / t / 1e9:7.1f} TFLOP/s", flush=True)
