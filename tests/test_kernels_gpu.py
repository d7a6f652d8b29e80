"""Kernel-level parity on a B200: every C-ABI launcher against a plain PyTorch fp32 restatement of the same op
(bf16-rounded operands, fp32 math, TF32 off).  Tolerances are written next to each check."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from tts_indic_server_f5_b200 import ops
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False

DEV = "cuda"


def rnd(*shape, seed=0, scale=1.0, dtype=torch.float32):
    g = torch.Generator("cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV).to(dtype)


def rel_err(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()


def check(name, got, want, rel, amax=None):
    r = rel_err(got, want)
    m = (got.float() - want.float()).abs().max().item()
    print(f"[{name}] rel-L2 {r:.3e} max-abs {m:.3e} (ref max {want.float().abs().max().item():.3e})")
    assert math.isfinite(r) and r <= rel, f"{name}: rel-L2 {r:.3e} > {rel:.1e} (max-abs {m:.3e})"
    if amax is not None:
        assert m <= amax, f"{name}: max-abs {m:.3e} > {amax:.1e}"


@pytest.mark.parametrize("M,N,K,bn", [(128, 256, 64, 256), (300, 256, 256, 256), (1000, 3072, 1024, 256),
                                      (515, 1024, 2048, 128), (257, 128, 128, 128), (130, 64, 192, 64),
                                      (200, 104, 128, 128), (4096, 2048, 1024, 256)])
def test_gemm_store_bf16(M, N, K, bn):
    A = rnd(M, K, seed=1, dtype=torch.bfloat16)
    B = rnd(N, K, seed=2, scale=1 / math.sqrt(K), dtype=torch.bfloat16)
    bias = rnd(N, seed=3)
    out = torch.full((M, N), 7.0, device=DEV, dtype=torch.bfloat16)
    ops.gemm(A, B, mode=ops.F5_EPI_STORE_BF16, bias=bias, out=out, block_n=bn)
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t() + bias
    check(f"gemm_bf16 {M}x{N}x{K}/{bn}", out, ref, rel=4e-3)   # bf16 output rounding: 2^-9 relative


@pytest.mark.parametrize("act", [0, 1, 2, 3])
def test_gemm_store_f32_addend_mask_act(act):
    M, N, K = 777, 512, 320
    A = rnd(M, K, seed=4, dtype=torch.bfloat16)
    B = rnd(N, K, seed=5, scale=1 / math.sqrt(K), dtype=torch.bfloat16)
    bias, add = rnd(N, seed=6), rnd(M, N, seed=7)
    row_pos = torch.arange(M, device=DEV, dtype=torch.int32)
    row_pos[100:116] = -1
    out = torch.zeros(M, N, device=DEV)
    out2 = torch.full((M, N), 3.0, device=DEV, dtype=torch.bfloat16)
    ops.gemm(A, B, mode=ops.F5_EPI_STORE_F32, act=act, bias=bias, out=out, out2=out2, addend=add, row_pos=row_pos,
             mask_rows=True)
    torch.cuda.synchronize()
    z = A.float() @ B.float().t() + bias
    z = [z, F.gelu(z, approximate="tanh"), F.gelu(z), F.mish(z)][act] + add
    check(f"gemm_f32 act{act}", out, z, rel=2e-5, amax=2e-4)   # fp32 accumulate, different summation order
    zm = z.clone()
    zm[100:116] = 0
    check(f"gemm_f32 act{act} bf16 copy", out2, zm, rel=4e-3)


def test_gemm_resid_gate():
    M, N, K = 640, 1024, 1024
    A = rnd(M, K, seed=8, dtype=torch.bfloat16)
    B = rnd(N, K, seed=9, scale=1 / math.sqrt(K), dtype=torch.bfloat16)
    bias, gate, x0 = rnd(N, seed=10), rnd(N, seed=11), rnd(M, N, seed=12)
    x = x0.clone()
    ops.gemm(A, B, mode=ops.F5_EPI_RESID_F32, bias=bias, gate=gate, resid=x)
    torch.cuda.synchronize()
    check("gemm_resid", x, x0 + gate * (A.float() @ B.float().t() + bias), rel=2e-5, amax=3e-4)
    x = x0.clone()
    ops.gemm(A, B, mode=ops.F5_EPI_RESID_F32, act=ops.F5_ACT_MISH, bias=bias, resid=x)
    torch.cuda.synchronize()
    check("gemm_resid_mish_nogate", x, x0 + F.mish(A.float() @ B.float().t() + bias), rel=2e-5, amax=3e-4)


@pytest.mark.parametrize("M,N,K,act,use_gate", [(130, 512, 1536, 0, True), (777, 1024, 2048, 0, True), (1000, 96, 128, 2, False),
                                               (4096, 1024, 1024, 0, True), (257, 512, 320, 1, False)])
def test_gemm_resid_shapes(M, N, K, act, use_gate):
    """Residual epilogue through the TMA reduce-add (N % 32 == 0): ragged M (rows clipped by the tensor map, whole 32-row boxes
    out of range), narrow N (units beyond N skipped), activations, missing gate."""
    A = rnd(M, K, seed=13, dtype=torch.bfloat16)
    B = rnd(N, K, seed=14, scale=1 / math.sqrt(K), dtype=torch.bfloat16)
    bias, gate, x0 = rnd(N, seed=15), rnd(N, seed=16), rnd(M + 64, N, seed=17)
    x = x0.clone()
    ops.gemm(A, B, mode=ops.F5_EPI_RESID_F32, act=act, bias=bias, gate=gate if use_gate else None, resid=x[:M])
    torch.cuda.synchronize()
    z = A.float() @ B.float().t() + bias
    z = [z, F.gelu(z, approximate="tanh"), F.gelu(z), F.mish(z)][act]
    want = x0[:M] + (gate if use_gate else 1.0) * z
    check(f"gemm_resid {M}x{N}x{K} act{act}", x[:M], want, rel=2e-5, amax=4e-4)
    assert torch.equal(x[M:], x0[M:]), "rows beyond M were touched"


@pytest.mark.parametrize("M,N,K", [(37900, 512, 256), (38912, 1024, 1024), (45000, 256, 192)])
def test_gemm_cluster_multicast(M, N, K):
    """>= 296 M-blocks: the 256-wide GEMM runs as clusters of two CTAs that multicast halves of the weight tile to each other.
    Odd and even numbers of M-blocks (a pair's second block beyond M), ragged last block, all three epilogue modes; the result
    must also equal the single-CTA form bit for bit (same MMAs, same order)."""
    A = rnd(M, K, seed=21, dtype=torch.bfloat16)
    B = rnd(N, K, seed=22, scale=1 / math.sqrt(K), dtype=torch.bfloat16)
    bias, gate, x0 = rnd(N, seed=23), rnd(N, seed=24), rnd(M, N, seed=25)
    ref = A.float() @ B.float().t() + bias
    out = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(A, B, mode=ops.F5_EPI_STORE_BF16, bias=bias, out=out)
    x = x0.clone()
    ops.gemm(A, B, mode=ops.F5_EPI_RESID_F32, bias=bias, gate=gate, resid=x)
    o32 = torch.zeros(M, N, device=DEV)
    ops.gemm(A, B, mode=ops.F5_EPI_STORE_F32, act=1, bias=bias, out=o32)
    torch.cuda.synchronize()
    check(f"cluster gemm bf16 {M}x{N}x{K}", out, ref, rel=4e-3)
    check(f"cluster gemm resid {M}x{N}x{K}", x, x0 + gate * ref, rel=2e-5, amax=4e-4)
    check(f"cluster gemm f32+gelu {M}x{N}x{K}", o32, F.gelu(ref, approximate="tanh"), rel=2e-5, amax=4e-4)
    half = M // 2 // 128 * 128                       # < 296 M-blocks per call: the single-CTA kernel
    out1 = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(A[:half], B, mode=ops.F5_EPI_STORE_BF16, bias=bias, out=out1[:half])
    ops.gemm(A[half:], B, mode=ops.F5_EPI_STORE_BF16, bias=bias, out=out1[half:])
    torch.cuda.synchronize()
    assert torch.equal(out, out1)


@pytest.mark.parametrize("M", [515, 38017])
@pytest.mark.parametrize("act", [1, 2, 3])
def test_gemm_store_bf16_activation_eight_epilogue_warps(M, act):
    """bf16-store GEMMs with an activation and 256-wide tiles run eight epilogue warps (two per TMEM lane quarter, each half of
    the tile's columns), alone (M = 515) and inside two-CTA clusters (M = 38017: 298 M-blocks, ragged last block)."""
    N, K = 768, 320
    A = rnd(M, K, seed=31, dtype=torch.bfloat16)
    B = rnd(N, K, seed=32, scale=1 / math.sqrt(K), dtype=torch.bfloat16)
    bias = rnd(N, seed=33)
    row_pos = torch.arange(M, device=DEV, dtype=torch.int32)
    row_pos[40:56] = -1
    out = torch.full((M, N), 5.0, device=DEV, dtype=torch.bfloat16)
    ops.gemm(A, B, mode=ops.F5_EPI_STORE_BF16, act=act, bias=bias, out=out, row_pos=row_pos, mask_rows=True)
    torch.cuda.synchronize()
    z = A.float() @ B.float().t() + bias
    z = [z, F.gelu(z, approximate="tanh"), F.gelu(z), F.mish(z)][act]
    z[40:56] = 0
    check(f"gemm bf16 act{act} M={M}", out, z, rel=4e-3)


def test_gemm_qkv_rope():
    D, M = 256, 400
    A = rnd(M, D, seed=13, dtype=torch.bfloat16)
    B = rnd(3 * D, D, seed=14, scale=1 / math.sqrt(D), dtype=torch.bfloat16)
    bias = rnd(3 * D, seed=15)
    pos = torch.cat([torch.arange(150), torch.full((16,), -1), torch.arange(234)]).to(DEV).to(torch.int32)
    inv = 1.0 / (10000.0 ** (torch.arange(0, 64, 2).float() / 64))
    ang = torch.arange(4096).float()[:, None] * inv[None]
    rope = torch.stack((ang.cos(), ang.sin()), dim=-1).reshape(4096, 64).contiguous().to(DEV)
    out = torch.zeros(M, 3 * D, device=DEV, dtype=torch.bfloat16)
    ops.gemm(A, B, mode=ops.F5_EPI_STORE_BF16, bias=bias, out=out, row_pos=pos, rope=rope, rope_period=D, rope_tiles=2)
    torch.cuda.synchronize()
    z = A.float() @ B.float().t() + bias
    live = pos >= 0
    ref = z.clone()
    for base in (0, D):
        blk = z[:, base:base + 64]
        c, s = ang.to(DEV)[pos.clamp_min(0).long()].cos(), ang.to(DEV)[pos.clamp_min(0).long()].sin()
        x0, x1 = blk[:, 0::2], blk[:, 1::2]
        rot = torch.stack((x0 * c - x1 * s, x1 * c + x0 * s), dim=-1).reshape(M, 64)
        ref[:, base:base + 64] = torch.where(live[:, None], rot, blk)
    check("gemm_qkv_rope", out[live], ref[live], rel=4e-3)


def test_gemm_grouped_conv31():
    C, G, Kw, n1, n2, gap = 256, 4, 31, 150, 230, 16
    M = gap + n1 + gap + n2 + gap
    x = torch.zeros(M, C, device=DEV)
    x[gap:gap + n1] = rnd(n1, C, seed=16)
    x[2 * gap + n1:2 * gap + n1 + n2] = rnd(n2, C, seed=17)
    xb = x.to(torch.bfloat16)
    w = rnd(C, C // G, Kw, seed=18, scale=1 / math.sqrt(C // G * Kw))
    bias = rnd(C, seed=19)
    wt = w.permute(2, 0, 1).contiguous().reshape(Kw * C, C // G).to(torch.bfloat16)   # [tap][out][in]
    pos = torch.full((M,), -1, dtype=torch.int32)
    pos[gap:gap + n1] = torch.arange(n1)
    pos[2 * gap + n1:2 * gap + n1 + n2] = torch.arange(n2)
    pos = pos.to(DEV)
    out = torch.full((M, C), 5.0, device=DEV, dtype=torch.bfloat16)
    ops.gemm(xb, wt, M=M, N=C, mode=ops.F5_EPI_STORE_BF16, act=ops.F5_ACT_MISH, bias=bias, out=out, row_pos=pos,
             mask_rows=True, block_n=64, num_taps=Kw, kc_per_tap=1, tap_pad=Kw // 2, a_grouped=True, b_tap_rows=C)
    torch.cuda.synchronize()
    wq = wt.float().reshape(Kw, C, C // G).permute(1, 2, 0).contiguous()
    for s0, n in ((gap, n1), (2 * gap + n1, n2)):
        ref = F.mish(F.conv1d(xb[s0:s0 + n].float().t()[None], wq, bias, padding=Kw // 2, groups=G))[0].t()
        check(f"conv31 seg@{s0}", out[s0:s0 + n], ref, rel=4e-3)
    assert out[:gap].abs().max().item() == 0 and out[gap + n1:2 * gap + n1].abs().max().item() == 0


def test_gemm_dense_conv7():
    Cin, Cin_pad, Cout, Kw, n, gap = 100, 128, 512, 7, 333, 3
    M = gap + n + gap
    x = torch.zeros(M, Cin_pad, device=DEV)
    x[gap:gap + n, :Cin] = rnd(n, Cin, seed=20)
    xb = x.to(torch.bfloat16)
    w = rnd(Cout, Cin, Kw, seed=21, scale=1 / math.sqrt(Cin * Kw))
    bias = rnd(Cout, seed=22)
    wt = torch.zeros(Kw, Cout, Cin_pad, device=DEV)
    wt[:, :, :Cin] = w.permute(2, 0, 1)
    wt = wt.reshape(Kw * Cout, Cin_pad).to(torch.bfloat16)
    out = torch.zeros(M, Cout, device=DEV)
    ops.gemm(xb, wt, M=M, N=Cout, mode=ops.F5_EPI_STORE_F32, bias=bias, out=out, num_taps=Kw, kc_per_tap=2,
             tap_pad=3, b_tap_rows=Cout)
    torch.cuda.synchronize()
    wq = wt.float().reshape(Kw, Cout, Cin_pad)[:, :, :Cin].permute(1, 2, 0).contiguous()
    ref = F.conv1d(xb[gap:gap + n, :Cin].float().t()[None], wq, bias, padding=3)[0].t()
    check("conv7 dense", out[gap:gap + n], ref, rel=2e-5, amax=2e-4)


def _attn_case(lens, H, scale_q=1.0, seed=30):
    D = H * 64
    gap = 16
    starts, rows = [], gap
    for n in lens:
        starts.append(rows)
        rows += n + gap
    rows = (rows + 127) // 128 * 128
    qkv = rnd(rows, 3 * D, seed=seed, dtype=torch.bfloat16)
    qkv[:, :D] *= scale_q
    tiles = []
    for s0, n in zip(starts, lens):
        for q0 in range(0, n, 256):
            tiles.append([s0 + q0, s0, n, min(256, n - q0)])
    tiles = torch.tensor(tiles, dtype=torch.int32, device=DEV)
    out = torch.zeros(rows, D, device=DEV, dtype=torch.bfloat16)
    ops.attention(qkv, tiles, out, H, 0, D, 2 * D, 0.125)
    torch.cuda.synchronize()
    for s0, n in zip(starts, lens):
        q, k, v = (qkv[s0:s0 + n, i * D:(i + 1) * D].float().view(n, H, 64).transpose(0, 1) for i in range(3))
        ref = F.scaled_dot_product_attention(q[None], k[None], v[None])[0].transpose(0, 1).reshape(n, D)
        check(f"attn n={n} H={H} sq={scale_q}", out[s0:s0 + n], ref, rel=1e-2)   # P and O rounded to bf16


def test_attention_small():
    _attn_case([128], 1)
    _attn_case([1], 1)
    _attn_case([129], 2)


def test_attention_ragged():
    _attn_case([300, 77, 513, 128], 4)
    _attn_case([785], 16, scale_q=6.0, seed=31)   # peaky softmax exercises the running-max rescale
    _attn_case([3069], 2, scale_q=3.0, seed=32)   # long-form (C3) length


def test_attention_many_items_per_cta():
    """Persistent schedule: every CTA walks several work items, with odd and even numbers of query tiles mixed (single-tile
    items idle softmax group B), tail key tiles and one-tile utterances in between."""
    g = torch.Generator("cpu").manual_seed(5)
    lens = [int(x) for x in torch.randint(60, 900, (48,), generator=g)] + [128, 129, 1, 257, 384]
    _attn_case(lens, 16, seed=33)
    _attn_case([1219] * 6 + [1100] * 5, 16, scale_q=4.0, seed=34)


def test_attention_properties_at_c2_size():
    """Size-independent properties at the benchmark's own size (C2: 64 utterances x 2 CFG halves, 16 heads, 158 k rows) — no
    reference needed: (1) softmax rows sum to one: with V == 1 every valid output is 1 up to the bf16 rounding of P;
    (2) keys and values permuted together inside each utterance leave the output unchanged (up to bf16 rounding of P, whose
    per-tile running max changes with the order); (3) gap rows are never written."""
    from tts_indic_server_f5_b200.layout import build_layout
    g = torch.Generator("cpu").manual_seed(0)
    lens = [469 + int(torch.randint(560, 941, (1,), generator=g)) for _ in range(64)]
    L = build_layout(lens)
    H, D = 16, 1024
    qkv = rnd(L.rows, 3 * D, seed=50, dtype=torch.bfloat16)
    tiles = L.attn_tiles.to(DEV)
    valid = (L.row_pos >= 0).to(DEV)
    qkv1 = qkv.clone()
    qkv1[:, 2 * D:] = 1.0
    out = torch.full((L.rows, D), 7.0, device=DEV, dtype=torch.bfloat16)
    ops.attention(qkv1, tiles, out, H, 0, D, 2 * D, 0.125)
    torch.cuda.synchronize()
    assert (out[valid].float() - 1.0).abs().max().item() <= 2 ** -6          # sum of ~1200 bf16-rounded probabilities / fp32 row sum
    assert torch.all(out[~valid] == 7.0)                                      # gap rows untouched
    base = torch.zeros(L.rows, D, device=DEV, dtype=torch.bfloat16)
    ops.attention(qkv, tiles, base, H, 0, D, 2 * D, 0.125)
    perm = torch.arange(L.rows)
    for half in (0, L.half_rows):
        for s0, n in zip(L.starts, lens):
            perm[half + s0: half + s0 + n] = half + s0 + torch.randperm(n, generator=g)
    qkvp = qkv.clone()
    qkvp[:, D:] = qkv[perm.to(DEV), D:]                                       # K and V rows permuted, Q in place
    outp = torch.zeros(L.rows, D, device=DEV, dtype=torch.bfloat16)
    ops.attention(qkvp, tiles, outp, H, 0, D, 2 * D, 0.125)
    torch.cuda.synchronize()
    check("attn key-permutation invariance (C2 size)", outp[valid], base[valid], rel=6e-3)


def test_gemm_linearity_at_c2_size():
    """D(A1 + A2) == D(A1) + D(A2) at the layer-GEMM size of the benchmark (fp32 outputs, exact bf16 operands chosen so that
    A1 + A2 is representable): exercises every tile, the cluster path and the TMA reduce epilogue (resid accumulates twice)."""
    M, N, K = 158976, 1024, 1024
    A1 = (torch.randint(-8, 9, (M, K), generator=torch.Generator().manual_seed(1)).float() / 8).to(DEV).to(torch.bfloat16)
    A2 = (torch.randint(-8, 9, (M, K), generator=torch.Generator().manual_seed(2)).float() / 8).to(DEV).to(torch.bfloat16)
    B = rnd(N, K, seed=3, scale=1 / 32, dtype=torch.bfloat16)
    x12 = torch.zeros(M, N, device=DEV)
    ops.gemm(A1, B, mode=ops.F5_EPI_RESID_F32, resid=x12)
    ops.gemm(A2, B, mode=ops.F5_EPI_RESID_F32, resid=x12)
    xs = torch.zeros(M, N, device=DEV)
    ops.gemm((A1.float() + A2.float()).to(torch.bfloat16), B, mode=ops.F5_EPI_RESID_F32, resid=xs)
    torch.cuda.synchronize()
    check("gemm linearity (C2 size)", x12, xs, rel=1e-6, amax=1e-4)


def test_layernorm_mod():
    for D in (128, 256, 512, 1024):
        M = 1001
        x = rnd(M, D, seed=40, scale=3.0) + 1.5
        a, b = rnd(D, seed=41), rnd(D, seed=42)
        y = torch.zeros(M, D, device=DEV, dtype=torch.bfloat16)
        ops.layernorm_mod(x, y, a, b, 1.0)
        torch.cuda.synchronize()
        check(f"ln_mod D={D}", y, F.layer_norm(x, (D,), eps=1e-6) * (1 + a) + b, rel=4e-3)
        ops.layernorm_mod(x, y, a, b, 0.0)
        torch.cuda.synchronize()
        check(f"ln_affine D={D}", y, F.layer_norm(x, (D,), a, b, eps=1e-6), rel=4e-3)


class _variant:
    """Select a kernel variant through the C-ABI setter for the duration of a block (A/B forms of the same op)."""

    def __init__(self, setter: str, v: int):
        self.setter, self.v = setter, v

    def __enter__(self):
        from tts_indic_server_f5_b200 import _lib
        self.old = getattr(_lib.lib, self.setter)(self.v)
        assert getattr(_lib.lib, self.setter)(0) == self.v          # out-of-range argument: query only

    def __exit__(self, *a):
        from tts_indic_server_f5_b200 import _lib
        getattr(_lib.lib, self.setter)(self.old)


@pytest.mark.parametrize("variant", [3, 2, 1])
def test_dwconv7_ln(variant):
    with _variant("f5_set_dwconv7_variant", variant):
        _dwconv7_ln_case()


def _ragged_pos(lens, gaps, M=None):
    """row_pos of utterances of `lens` rows separated by `gaps[i]` dead rows (gaps has one more entry than lens)."""
    pos, starts = [], []
    for g, n in zip(gaps, lens):
        pos += [-1] * g
        starts.append(len(pos))
        pos += list(range(n))
    pos += [-1] * gaps[-1]
    if M is not None:
        pos += [-1] * (M - len(pos))
    return torch.tensor(pos, dtype=torch.int32), starts


@pytest.mark.parametrize("C", [512, 256])
def test_dwconv7_ln_variants_agree_on_a_ragged_pack(C):
    """The three forms of f5_dwconv7_ln on a pack that exercises every window case: gaps of 1 and 2 rows (a +-3 window then
    spans two utterances: only the key / position test keeps them apart), an utterance of a single row, utterances that start
    at row 0 / end at the last row, run boundaries inside utterances.  Form 2 repeats form 1's operation order (bit-identical);
    form 3 combines per-warp LayerNorm statistics (Chan) and may differ by one bf16 rounding.  All against torch."""
    lens = [1, 37, 2, 701, 3, 1500, 64, 5, 333, 4097]
    gaps = [0, 1, 2, 1, 8, 3, 1, 16, 2, 1, 0]
    pos, starts = _ragged_pos(lens, gaps)
    M = pos.numel()
    pos = pos.to(DEV)
    x = rnd(M, C, seed=143) * 3 + 0.5
    x[pos < 0] = float("nan")                      # dead rows must never be read into a live output
    w, bias, lw, lb = rnd(C, 7, seed=144, scale=0.4), rnd(C, seed=145), rnd(C, seed=146) + 1, rnd(C, seed=147)
    outs = {}
    for v in (1, 2, 3):
        y = torch.full((M, C), 9.0, device=DEV, dtype=torch.bfloat16)
        with _variant("f5_set_dwconv7_variant", v):
            ops.dwconv7_ln(x, y, pos, w, bias, lw, lb)
        torch.cuda.synchronize()
        assert torch.isfinite(y.float()).all(), f"variant {v}: a dead row leaked"
        assert y[pos < 0].abs().max().item() == 0
        outs[v] = y
    for s0, n in zip(starts, lens):
        h = F.conv1d(x[s0:s0 + n].t()[None], w[:, None, :], bias, padding=3, groups=C)[0].t()
        ref = F.layer_norm(h, (C,), lw, lb, eps=1e-6)
        for v in (1, 2, 3):
            check(f"dwconv7_ln v{v} C={C} n={n}", outs[v][s0:s0 + n], ref, rel=4e-3)
    assert torch.equal(outs[2], outs[1])
    d = (outs[3].float() - outs[1].float()).abs()
    assert (d <= outs[1].float().abs() * 2 ** -7 + 4e-6).all()      # at most one bf16 ulp (+ fp32 round-off where the value is ~0)
    assert (d > 0).float().mean().item() < 0.02
    # split-operand (fp32 mode) planes: hi + lo carries 16 mantissa bits
    planes = {}
    for v in (1, 2, 3):
        y = torch.full((M, 2 * C), 9.0, device=DEV, dtype=torch.bfloat16)
        with _variant("f5_set_dwconv7_variant", v):
            ops.dwconv7_ln(x, y, pos, w, bias, lw, lb, lo_off=C)
        torch.cuda.synchronize()
        planes[v] = y[:, :C].float() + y[:, C:].float()
        assert y[pos < 0].abs().max().item() == 0
    assert torch.equal(planes[2], planes[1])
    live = pos >= 0
    check(f"dwconv7_ln planes v3 vs v1 C={C}", planes[3][live], planes[1][live], rel=2e-5)


def _dwconv7_ln_case():
    for C in (128, 512):
        n1, n2, gap = 100, 57, 5
        M = gap + n1 + gap + n2 + gap
        pos = torch.full((M,), -1, dtype=torch.int32)
        pos[gap:gap + n1] = torch.arange(n1)
        pos[2 * gap + n1:2 * gap + n1 + n2] = torch.arange(n2)
        pos = pos.to(DEV)
        x = rnd(M, C, seed=43)      # garbage in gap rows must not leak
        w, bias, lw, lb = rnd(C, 7, seed=44, scale=0.4), rnd(C, seed=45), rnd(C, seed=46) + 1, rnd(C, seed=47)
        y = torch.full((M, C), 9.0, device=DEV, dtype=torch.bfloat16)
        ops.dwconv7_ln(x, y, pos, w, bias, lw, lb)
        torch.cuda.synchronize()
        for s0, n in ((gap, n1), (2 * gap + n1, n2)):
            h = F.conv1d(x[s0:s0 + n].t()[None], w[:, None, :], bias, padding=3, groups=C)[0].t()
            check(f"dwconv7_ln C={C}", y[s0:s0 + n], F.layer_norm(h, (C,), lw, lb, eps=1e-6), rel=4e-3)
        assert y[:gap].abs().max().item() == 0


def test_grn():
    C, segs = 1024, [(3, 200), (220, 77)]
    x = rnd(300, C, seed=48, dtype=torch.bfloat16)
    x0 = x.clone()
    seg = torch.tensor(segs, dtype=torch.int32, device=DEV)
    sumsq = torch.zeros(len(segs), C, device=DEV)
    gamma, beta = rnd(C, seed=49), rnd(C, seed=50)
    ops.grn(x, seg, sumsq, gamma, beta)
    torch.cuda.synchronize()
    for s0, n in segs:
        xs = x0[s0:s0 + n].float()
        gx = xs.norm(p=2, dim=0, keepdim=True)
        nx = gx / (gx.mean(dim=-1, keepdim=True) + 1e-6)
        check("grn", x[s0:s0 + n], gamma * (xs * nx) + beta + xs, rel=4e-3)
    assert torch.equal(x[:3], x0[:3])


def test_gather_pack_where_silu_time():
    M, C, V = 200, 512, 50
    ids = torch.randint(0, V, (M,), dtype=torch.int32, device=DEV)
    pos = torch.arange(M, dtype=torch.int32, device=DEV) - 10
    emb, table = rnd(V, C, seed=51), rnd(64, C, seed=52)
    out = torch.full((M, C), 1.0, device=DEV)
    ops.text_gather_pos(ids, pos, emb, table, out)
    torch.cuda.synchronize()
    ref = emb[ids.long()] + table[pos.clamp(0, 63).long()]
    ref[:10] = 0
    assert torch.equal(out, ref)
    src = rnd(M, 100, seed=53)
    dst = torch.full((M, 256), 2.0, device=DEV, dtype=torch.bfloat16)
    rows = torch.arange(M, dtype=torch.int32, device=DEV).flip(0).contiguous()
    rows[5] = -1
    ops.pack_bf16(src, dst, 128, 100, 128, src_rows=rows)
    torch.cuda.synchronize()
    want = torch.zeros(M, 128, device=DEV)
    want[:, :100] = src[rows.clamp_min(0).long()]
    want[5] = 0
    assert torch.equal(dst[:, 128:], want.to(torch.bfloat16)) and (dst[:, :128] == 2).all()
    x, c = rnd(M, 128, seed=54), rnd(M, 128, seed=55)
    flag = (torch.arange(M, device=DEV) % 3 == 0).to(torch.int32)
    x1 = x.clone()
    ops.where_rows(x1, c, flag, 100)
    torch.cuda.synchronize()
    want = x.clone()
    want[flag.bool(), :100] = c[flag.bool(), :100]
    assert torch.equal(x1, want)
    v = rnd(1000, 6144, seed=56)
    o = torch.zeros(1000, 6144, device=DEV, dtype=torch.bfloat16)
    ops.silu_bf16(v, o)
    torch.cuda.synchronize()
    check("silu", o, F.silu(v), rel=4e-3)
    t = torch.linspace(0, 1, 33, device=DEV)
    half = 128
    freqs = torch.exp(torch.arange(half).float() * -(math.log(10000) / (half - 1))).to(DEV)
    te = torch.zeros(33, 256, device=DEV, dtype=torch.bfloat16)
    ops.time_sinus(t, freqs, te)
    torch.cuda.synchronize()
    arg = (1000 * t)[:, None] * freqs[None]
    check("time_sinus", te, torch.cat((arg.sin(), arg.cos()), -1), rel=5e-3)


def test_cfg_euler():
    half, C, Cp = 300, 100, 128
    x0 = rnd(half, Cp, seed=57)
    pred = rnd(2 * half, Cp, seed=58)
    pos = torch.arange(half, dtype=torch.int32, device=DEV)
    pos[40:56] = -1
    dts = torch.tensor([0.1, 0.03, 0.2], device=DEV)
    xb = torch.full((2 * half, Cp), 4.0, device=DEV, dtype=torch.bfloat16)
    x = x0.clone()
    ops.cfg_euler(x, pred, half, C, pos, dts, 1, 2.0, xb, Cp)
    torch.cuda.synchronize()
    pc, pu = pred[:half, :C], pred[half:, :C]
    want = x0.clone()
    live = pos >= 0
    want[live, :C] = (x0[:, :C] + dts[1] * (pc + (pc - pu) * 2.0))[live]
    check("cfg_euler x", x, want, rel=1e-6)
    wb = torch.zeros(half, Cp, device=DEV)
    wb[live, :C] = want[live, :C]
    assert torch.equal(xb[:half], wb.to(torch.bfloat16)) and torch.equal(xb[half:], wb.to(torch.bfloat16))


@pytest.mark.parametrize("variant", [2, 1])
def test_istft_vs_torch(variant):
    with _variant("f5_set_istft_variant", variant):
        _istft_vs_torch_case()


@pytest.mark.parametrize("rows", [1, 2, 7, 1001])
def test_istft_frames_variants_agree(rows):
    """Windowed frames of the real-input form (512-point FFT, two frames per warp, MUFU sin / cos / ex2) against the first
    kernel (1024-point complex FFT, sincosf / expf) and against torch.fft.irfft: odd and even frame counts (the last pair has
    one frame), phases well outside [-pi, pi], clipped and tiny magnitudes."""
    spec = rnd(rows, 1152, seed=259)
    spec[:, :513] = spec[:, :513] * 2.5 - 0.5          # log-magnitudes: exp() from ~1e-4 to the 1e2 clip
    spec[:, 513:1026] *= 25.0                          # phases up to +-100 rad
    window = torch.hann_window(1024, device=DEV)
    fr = {}
    for v in (1, 2):
        fr[v] = torch.full((rows, 1024), 7.0, device=DEV)
        with _variant("f5_set_istft_variant", v):
            ops.call("f5_istft_frames", ops.ptr(spec), spec.stride(0), rows, ops.ptr(window), ops.ptr(fr[v]), ops.stream_ptr())
    torch.cuda.synchronize()
    mag = spec[:, :513].double().exp().clip(max=1e2)
    ph = spec[:, 513:1026].double()
    S = mag * (ph.cos() + 1j * ph.sin())
    ref = (torch.fft.irfft(S, n=1024, dim=1) * window.double()).float()
    check(f"istft_frames v1 rows={rows}", fr[1], ref, rel=2e-5)
    check(f"istft_frames v2 rows={rows}", fr[2], ref, rel=2e-5)
    check(f"istft_frames v2 vs v1 rows={rows}", fr[2], fr[1], rel=2e-5)


def _istft_vs_torch_case():
    segs = [(2, 50), (60, 1), (70, 129)]
    rows = 210
    spec = rnd(rows, 1152, seed=59)
    spec[:, :513] = spec[:, :513] * 1.5 + 0.5      # some magnitudes hit the 1e2 clip
    window = torch.hann_window(1024, device=DEV)
    frames = torch.zeros(rows, 1024, device=DEV)
    offs, tot = [], 0
    for _, T in segs:
        offs.append(tot)
        tot += 256 * (T - 1)
    seg = torch.tensor([[r0, T, o, 0] for (r0, T), o in zip(segs, offs)], dtype=torch.int32, device=DEV)
    wav = torch.zeros(max(tot, 1), device=DEV)
    gains = torch.tensor([1.0, 1.0, 0.5], device=DEV)
    ops.istft(spec, window, frames, seg, 256 * 128, wav, gains)
    torch.cuda.synchronize()
    for (r0, T), o, g in zip(segs, offs, gains.tolist()):
        if T < 2:
            continue
        mag = spec[r0:r0 + T, :513].exp().clip(max=1e2)
        ph = spec[r0:r0 + T, 513:1026]
        S = (mag * (ph.cos() + 1j * ph.sin())).t()[None]
        ref = torch.istft(S, 1024, 256, 1024, window, center=True)[0] * g
        check(f"istft T={T}", wav[o:o + 256 * (T - 1)], ref, rel=2e-5)


# ------------------------------------------------------------------------------------------------ round 2: device noise
def test_randn_rows_vs_philox_oracle():
    """f5_randn_rows against oracle/philox.py (itself pinned to the Random123 known answers): identical Philox words, so the
    normals agree to float round-off of log / sincospi (1e-5 abs on values up to ~5); gap rows are zero; the draw of an
    utterance does not depend on where it sits in the pack."""
    from oracle import philox as P
    from tts_indic_server_f5_b200.layout import build_layout
    lens, seeds = [37, 130, 5], [7, 0xDEADBEEFCAFEF00D, 2 ** 63 + 11]
    L = build_layout(lens)
    R = L.half_rows
    x = torch.full((R, 128), 9.0, device=DEV)
    sd = torch.tensor([s - (1 << 64) if s >= (1 << 63) else s for s in seeds], dtype=torch.int64, device=DEV)
    ops.randn_rows(x, 100, L.row_pos[:R].to(DEV), L.row_utt.to(DEV), sd)
    torch.cuda.synchronize()
    got = x.cpu().numpy()
    assert (got[:, 100:] == 9.0).all()                                   # columns past C untouched
    live = L.row_pos[:R].numpy() >= 0
    assert (got[~live, :100] == 0.0).all()
    for s0, n, seed in zip(L.starts, lens, seeds):
        want = P.randn_rows(seed, n)
        err = float(abs(got[s0:s0 + n, :100] - want).max())
        print(f"[randn_rows] n={n} max-abs {err:.2e}")
        assert err < 2e-5
    L2 = build_layout([130])                                              # the same utterance alone: bit-identical rows
    y = torch.zeros(L2.half_rows, 128, device=DEV)
    ops.randn_rows(y, 100, L2.row_pos[:L2.half_rows].to(DEV), L2.row_utt.to(DEV), sd[1:2].contiguous())
    assert torch.equal(y[L2.starts[0]:L2.starts[0] + 130, :100], x[L.starts[1]:L.starts[1] + 130, :100])
    big = torch.zeros(4096 + 128, 128, device=DEV)                         # moments at the longest utterance the path allows
    Lb = build_layout([4096])
    ops.randn_rows(big[:Lb.half_rows], 100, Lb.row_pos[:Lb.half_rows].to(DEV), Lb.row_utt.to(DEV), sd[:1].contiguous())
    z = big[Lb.starts[0]:Lb.starts[0] + 4096, :100]
    assert abs(float(z.mean())) < 5e-3 and abs(float(z.std()) - 1) < 5e-3 and float(z.abs().max()) < 6.0


# ------------------------------------------------------------------------------------------------ round 2: fp32 precision mode
def _planes(x, lo_first_col):
    """fp32 [M, K] -> bf16 [M, 2K] = hi | lo through the product's own splitter (f5_pack_bf16 with lo_off)."""
    M, K = x.shape
    out = torch.zeros(M, 2 * K, device=DEV, dtype=torch.bfloat16)
    ops.pack_bf16(x.contiguous(), out, 0, K, K, lo_off=lo_first_col)
    return out


def test_split_planes_reconstruct_the_value():
    """hi + lo carries 16 mantissa bits: |v - (hi + lo)| <= 2^-17 |v| (pack, LayerNorm and silu writers)."""
    from tts_indic_server_f5_b200.engine import split_planes
    x = rnd(257, 384, seed=70, scale=3.0)
    p = _planes(x, 384).float()
    rec = p[:, :384] + p[:, 384:]
    assert ((rec - x).abs() <= x.abs() * 2.0 ** -16 + 1e-30).all()
    w3 = split_planes(x.cpu())
    assert torch.equal(w3[:257], w3[514:]) and ((w3[:257].float() + w3[257:514].float() - x.cpu()).abs() <= x.cpu().abs() * 2.0 ** -16).all()
    a, b = rnd(384, seed=71), rnd(384, seed=72)
    y = torch.zeros(257, 768, device=DEV, dtype=torch.bfloat16)
    y32 = torch.zeros(257, 384, device=DEV)
    ops.layernorm_mod(x, y, a, b, 1.0, y32=y32, lo_off=384)
    rec = y[:, :384].float() + y[:, 384:].float()
    assert ((rec - y32).abs() <= y32.abs() * 2.0 ** -16 + 1e-30).all()
    s_out = torch.zeros(257, 768, device=DEV, dtype=torch.bfloat16)
    ops.silu_bf16(x, s_out, split=True)
    rec = s_out[:, :384].float() + s_out[:, 384:].float()
    ref = F.silu(x)
    assert ((rec - ref).abs() <= ref.abs() * 2.0 ** -15 + 1e-6).all()


@pytest.mark.parametrize("M,N,K,mode", [(300, 256, 128, "f32"), (1000, 3072, 1024, "f32"), (515, 1024, 2048, "resid"), (257, 104, 640, "f32")])
def test_gemm_split_operand_is_fp32_class(M, N, K, mode):
    """D = A_hi B_hi^T + A_hi B_lo^T + A_lo B_hi^T against a float64 product of the SAME fp32 operands: rel-L2 <= 2e-5
    (the plain bf16-operand GEMM sits at ~3e-3 on these inputs)."""
    from tts_indic_server_f5_b200.engine import split_planes
    A, W = rnd(M, K, seed=80), rnd(N, K, seed=81, scale=1 / math.sqrt(K))
    bias = rnd(N, seed=82)
    Ap, W3 = _planes(A, K), split_planes(W.cpu()).to(DEV)
    ref = (A.double() @ W.double().t() + bias.double())
    if mode == "f32":
        out = torch.zeros(M, N, device=DEV)
        ops.gemm(Ap, W3, mode=ops.F5_EPI_STORE_F32, bias=bias, out=out, split=True)
    else:
        gate, x0 = rnd(N, seed=83), rnd(M, N, seed=84)
        out = x0.clone()
        ops.gemm(Ap, W3, mode=ops.F5_EPI_RESID_F32, bias=bias, gate=gate, resid=out, split=True)
        ref = x0.double() + gate.double() * ref
    torch.cuda.synchronize()
    check(f"split gemm {M}x{N}x{K} {mode}", out, ref.float(), rel=2e-5)
    plain = torch.zeros(M, N, device=DEV)
    ops.gemm(A.to(torch.bfloat16), W.to(torch.bfloat16), mode=ops.F5_EPI_STORE_F32, bias=bias, out=plain)
    if mode == "f32":
        assert rel_err(plain, ref.float()) > 20 * rel_err(out, ref.float())


def test_gemm_split_operand_grouped_conv31():
    from tts_indic_server_f5_b200.engine import split_planes
    C, G, Kw, n1, n2, gap = 256, 4, 31, 150, 230, 16
    M = gap + n1 + gap + n2 + gap
    x = torch.zeros(M, C, device=DEV)
    x[gap:gap + n1] = rnd(n1, C, seed=16)
    x[2 * gap + n1:2 * gap + n1 + n2] = rnd(n2, C, seed=17)
    w = rnd(C, C // G, Kw, seed=18, scale=1 / math.sqrt(C // G * Kw))
    bias = rnd(C, seed=19)
    wt = w.permute(2, 0, 1).contiguous().reshape(Kw * C, C // G)                 # [tap][out][in], fp32
    out = torch.zeros(M, C, device=DEV)
    ops.gemm(_planes(x, C), split_planes(wt.cpu()).to(DEV), M=M, N=C, mode=ops.F5_EPI_STORE_F32, act=ops.F5_ACT_MISH, bias=bias, out=out,
             block_n=64, num_taps=Kw, kc_per_tap=1, tap_pad=Kw // 2, a_grouped=True, b_tap_rows=C, split=True)
    torch.cuda.synchronize()
    for s0, n in ((gap, n1), (2 * gap + n1, n2)):
        ref = F.mish(F.conv1d(x[s0:s0 + n].double().t()[None], w.double(), bias.double(), padding=Kw // 2, groups=G))[0].t()
        check(f"split conv31 seg@{s0}", out[s0:s0 + n], ref.float(), rel=3e-5)


def test_attention_f32_vs_sdpa():
    """f5_attention_f32 against fp32 SDPA per utterance and head, with the x-transformers rotary embedding on head 0
    (interleaved pairs, modules.py:414-426); output planes hi + lo reconstruct the fp32 result to 2^-16."""
    from tts_indic_server_f5_b200.layout import build_layout
    lens, H, D = [300, 77, 513, 128, 1], 4, 256
    L = build_layout(lens)
    qkv = rnd(L.rows, 3 * D, seed=90)
    qkv[:, :D] *= 3.0                                                            # peaky rows exercise the running-max rescale
    inv = 1.0 / (10000.0 ** (torch.arange(0, 64, 2).float() / 64))
    ra = torch.arange(1024).float()[:, None] * inv[None]
    rope = torch.stack((ra.cos(), ra.sin()), dim=-1).reshape(1024, 64).to(DEV)
    out = torch.zeros(L.rows, 2 * D, device=DEV, dtype=torch.bfloat16)
    ops.attention_f32(qkv, L.attn_tiles.to(DEV), out, H, 0, D, 2 * D, 0.125, rope=rope, lo_off=D)
    torch.cuda.synchronize()
    got = out[:, :D].float() + out[:, D:].float()

    def rot(x, n):                                                               # x [n, 64]
        c, s = ra[:n].cos().to(DEV), ra[:n].sin().to(DEV)
        x0, x1 = x[:, 0::2], x[:, 1::2]
        return torch.stack((x0 * c - x1 * s, x1 * c + x0 * s), dim=-1).reshape(n, 64)

    worst = 0.0
    for half in (0, L.half_rows):
        for s0, n in zip(L.starts, lens):
            blk = qkv[half + s0: half + s0 + n].double()
            for h in range(H):
                q, k, v = blk[:, h * 64:(h + 1) * 64], blk[:, D + h * 64:D + (h + 1) * 64], blk[:, 2 * D + h * 64:2 * D + (h + 1) * 64]
                if h == 0:
                    q, k = rot(q.float(), n).double(), rot(k.float(), n).double()
                ref = torch.softmax(q @ k.t() * 0.125, dim=-1) @ v
                worst = max(worst, float((got[half + s0: half + s0 + n, h * 64:(h + 1) * 64].double() - ref).abs().max()))
    print(f"[attention_f32] max-abs {worst:.2e}")
    assert worst < 2e-5
    gaps = (L.row_pos < 0).to(DEV)
    assert out[gaps].abs().max().item() == 0
