"""Thin tensor-level wrappers over the C ABI (`include/f5_b200.h`).  Each wrapper only validates shapes/dtypes,
extracts raw pointers and forwards to the launcher on torch's current stream.  No arithmetic happens here and there
is no fallback: if the CUDA library cannot run the op, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib
from ._lib import (F5_ACT_GELU_ERF, F5_ACT_GELU_TANH, F5_ACT_MISH, F5_ACT_NONE, F5_EPI_RESID_F32,  # noqa: F401
                   F5_EPI_STORE_BF16, F5_EPI_STORE_F32, GemmArgs, call, ptr, stream_ptr)

BF16, F32, I32 = torch.bfloat16, torch.float32, torch.int32


def _ld(t: torch.Tensor) -> int:
    assert t.dim() == 2 and t.stride(1) == 1, "row-major 2-D tensor expected"
    return t.stride(0)


SMALL_M_RULE = os.environ.get("F5_SMALL_M_RULE", "1") != "0"   # 0 switches the small-batch rule off (A/B runs)
NUM_SMS = 148
COST_128 = 0.58    # time of a 128-wide tile relative to a 256-wide one of the same K (same A tile, half the B tile and MMA work)


def pick_block_n(N: int, M: int | None = None) -> int:
    """Tile width of the persistent GEMM.  256 wherever N allows (one A tile feeds 256 columns) EXCEPT where the tile count is
    so small that whole waves of SMs stand idle — a single request has M = 1792 rows: 56 tiles for the N = 1024 GEMMs (92 SMs
    idle, and each busy one streams a 1.5 MB weight slab through its own ~120 GB/s L2 port: 12 us for a 4 us MMA chain, ncu
    launch list profiles/r02_launches_c1.csv), 168 tiles = two waves of 148 for QKV.  The rule compares waves x tile cost of
    the two widths; for a full batch (thousands of tiles) it always answers 256."""
    if N % 256 == 0:
        if M is not None and SMALL_M_RULE and N % 128 == 0:
            mt = (M + 127) // 128
            waves_256 = -(-(mt * (N // 256)) // NUM_SMS)
            waves_128 = -(-(mt * (N // 128)) // NUM_SMS)
            if waves_128 * COST_128 < waves_256:
                return 128
        return 256
    if N >= 128:
        return 128
    return 64


def gemm(A: torch.Tensor, B: torch.Tensor, *, M: int | None = None, N: int | None = None, mode: int, act: int = F5_ACT_NONE,
         bias=None, gate=None, out=None, out2=None, addend=None, resid=None, row_pos=None, mask_rows=False,
         rope=None, rope_period=0, rope_tiles=0, block_n: int | None = None,
         num_taps=1, kc_per_tap: int | None = None, tap_pad=0, a_grouped=False, b_tap_rows=0, num_sms=0,
         split: bool = False) -> None:
    """D = A @ B^T with fused epilogue (see f5_gemm_bf16).  A [a_rows, K] bf16, B [b_rows, Kb] bf16.
    split: split-operand mode — A is [rows, 2K] = hi | lo planes, B holds the three row-stacked planes [hi | lo | hi] of each
    tap (3 x the rows of the ordinary weight): D = A_hi B_hi^T + A_hi B_lo^T + A_lo B_hi^T."""
    assert A.dtype == BF16 and B.dtype == BF16 and A.is_cuda and B.is_cuda
    a = GemmArgs()
    if split:
        assert A.shape[1] % 128 == 0 and B.shape[0] % 3 == 0
        a.taps_per_seg, a.a_lo_off = num_taps, A.shape[1] // 2
        if num_taps == 1:
            b_tap_rows = B.shape[0] // 3 if N is None else N
            N = b_tap_rows
        num_taps = 3 * num_taps
    a.A, a.B = ptr(A), ptr(B)
    a.lda, a.ldb = _ld(A), _ld(B)
    a.a_rows, a.a_cols = A.shape
    a.b_rows, a.b_cols = B.shape
    a.M = A.shape[0] if M is None else M
    a.N = (B.shape[0] if num_taps == 1 else b_tap_rows) if N is None else N
    a.block_n = block_n or pick_block_n(a.N, a.M)
    a.num_taps = num_taps
    a.kc_per_tap = kc_per_tap if kc_per_tap is not None else (B.shape[1] + 63) // 64      # split: B has K columns, A has 2K
    a.tap_pad, a.a_grouped, a.b_tap_rows = tap_pad, int(a_grouped), b_tap_rows
    a.mode, a.act = mode, act
    for name, t, dt in (("bias", bias, F32), ("gate", gate, F32), ("row_pos", row_pos, I32), ("rope", rope, F32)):
        if t is not None:
            assert t.dtype == dt and t.is_contiguous()
        setattr(a, name, ptr(t))
    if out is not None:
        assert out.dtype == (BF16 if mode == F5_EPI_STORE_BF16 else F32)
        a.out, a.ldo = ptr(out), _ld(out)
    if out2 is not None:
        assert out2.dtype == BF16
        a.out2, a.ldo2 = ptr(out2), _ld(out2)
    if addend is not None:
        assert addend.dtype == F32
        a.addend, a.ld_add = ptr(addend), _ld(addend)
    if resid is not None:
        assert resid.dtype == F32
        a.resid, a.ldr = ptr(resid), _ld(resid)
    a.mask_rows = int(mask_rows)
    a.rope_period, a.rope_tiles = rope_period, rope_tiles
    a.num_sms = num_sms
    call("f5_gemm_bf16", C.byref(a), stream_ptr())


def attention_f32(qkv: torch.Tensor, tiles: torch.Tensor, out: torch.Tensor, heads: int, q_col: int, k_col: int, v_col: int,
                  softmax_scale: float = 0.125, rope: torch.Tensor | None = None, lo_off: int = 0) -> None:
    """fp32 attention (fp32 precision mode): qkv fp32, out bf16 (lo_off > 0: hi | lo planes), RoPE on head 0 from `rope`."""
    assert qkv.dtype == F32 and out.dtype == BF16 and tiles.dtype == I32 and tiles.is_contiguous() and tiles.shape[1] == 4
    assert rope is None or (rope.dtype == F32 and rope.is_contiguous() and rope.shape[1] == 64)
    call("f5_attention_f32", ptr(qkv), _ld(qkv), q_col, k_col, v_col, heads, ptr(tiles), tiles.shape[0], ptr(rope), ptr(out),
         _ld(out), lo_off, float(softmax_scale), stream_ptr())


def attention(qkv: torch.Tensor, tiles: torch.Tensor, out: torch.Tensor, heads: int, q_col: int, k_col: int, v_col: int,
              softmax_scale: float = 0.125) -> None:
    assert qkv.dtype == BF16 and out.dtype == BF16 and tiles.dtype == I32 and tiles.is_contiguous() and tiles.shape[1] == 4
    call("f5_attention_d64", ptr(qkv), _ld(qkv), qkv.shape[0], q_col, k_col, v_col, heads, ptr(tiles), tiles.shape[0],
         ptr(out), _ld(out), float(softmax_scale), stream_ptr())


def layernorm_mod(x: torch.Tensor, y: torch.Tensor | None, a: torch.Tensor, b: torch.Tensor, a_off: float, eps: float = 1e-6,
                  M: int | None = None, y32: torch.Tensor | None = None, lo_off: int = 0) -> None:
    assert x.dtype == F32 and a.dtype == F32 and b.dtype == F32
    assert (y is None or y.dtype == BF16) and (y32 is None or y32.dtype == F32)
    call("f5_layernorm_mod", ptr(x), _ld(x), ptr(y), _ld(y) if y is not None else 0, ptr(y32),
         _ld(y32) if y32 is not None else 0, x.shape[0] if M is None else M, x.shape[1], ptr(a), ptr(b),
         float(a_off), float(eps), lo_off, stream_ptr())


def dwconv7_ln(x, y, row_pos, w, bias, ln_w, ln_b, eps: float = 1e-6, lo_off: int = 0) -> None:
    assert x.dtype == F32 and y.dtype == BF16 and row_pos.dtype == I32 and w.dtype == F32 and w.is_contiguous()
    call("f5_dwconv7_ln", ptr(x), _ld(x), ptr(y), _ld(y), x.shape[0], x.shape[1], ptr(row_pos), ptr(w), ptr(bias),
         ptr(ln_w), ptr(ln_b), float(eps), lo_off, stream_ptr())


def grn_f32(x: torch.Tensor, seg_rows: torch.Tensor, sumsq: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, C_: int) -> None:
    """In-place GRN on fp32 activations (fp32 precision mode); x [rows, >= C_]."""
    assert x.dtype == F32 and seg_rows.dtype == I32 and sumsq.dtype == F32
    S = seg_rows.shape[0]
    call("f5_grn_sumsq_f32", ptr(x), _ld(x), C_, ptr(seg_rows), S, ptr(sumsq), stream_ptr())
    call("f5_grn_apply_f32", ptr(x), _ld(x), C_, ptr(seg_rows), S, ptr(sumsq), ptr(gamma), ptr(beta), stream_ptr())


def grn(x: torch.Tensor, seg_rows: torch.Tensor, sumsq: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor) -> None:
    """In-place GRN on bf16 activations; seg_rows int32 [S,2] = {row0, rows}; sumsq fp32 [S, C] scratch."""
    assert x.dtype == BF16 and seg_rows.dtype == I32 and sumsq.dtype == F32
    S, Cc = seg_rows.shape[0], x.shape[1]
    call("f5_grn_sumsq", ptr(x), _ld(x), Cc, ptr(seg_rows), S, ptr(sumsq), stream_ptr())
    call("f5_grn_apply", ptr(x), _ld(x), Cc, ptr(seg_rows), S, ptr(sumsq), ptr(gamma), ptr(beta), stream_ptr())


def text_gather_pos(ids, row_pos, emb, pos_table, out) -> None:
    assert ids.dtype == I32 and row_pos.dtype == I32 and emb.dtype == F32 and pos_table.dtype == F32 and out.dtype == F32
    call("f5_text_gather_pos", ptr(ids), ptr(row_pos), ptr(emb), ptr(pos_table), pos_table.shape[0], ptr(out), _ld(out),
         out.shape[0], out.shape[1], stream_ptr())


def pack_bf16(src, dst, dst_col: int, C_: int, C_pad: int, src_rows=None, row_pos=None, M: int | None = None,
              lo_off: int = 0) -> None:
    assert src.dtype == F32 and dst.dtype == BF16
    call("f5_pack_bf16", ptr(src), _ld(src), ptr(dst), _ld(dst), dst_col, dst.shape[0] if M is None else M, C_, C_pad,
         ptr(src_rows), ptr(row_pos), lo_off, stream_ptr())


def where_rows(x, c, flag, C_: int) -> None:
    assert x.dtype == F32 and c.dtype == F32 and flag.dtype == I32
    call("f5_where_rows", ptr(x), _ld(x), ptr(c), _ld(c), ptr(flag), x.shape[0], C_, stream_ptr())


def cfg_euler(x, pred, half_rows: int, C_: int, row_pos, dts, step: int, cfg_strength: float, xb, C_pad: int) -> None:
    assert x.dtype == F32 and pred.dtype == F32 and xb.dtype == BF16 and dts.dtype == F32
    call("f5_cfg_euler", ptr(x), _ld(x), ptr(pred), _ld(pred), half_rows, C_, ptr(row_pos), ptr(dts), step,
         float(cfg_strength), ptr(xb), _ld(xb), C_pad, stream_ptr())


def randn_rows(x, C_: int, row_pos, row_utt, utt_seed, M: int | None = None) -> None:
    """x[row, :C] = N(0, 1) noise of (utt_seed[row_utt[row]], row_pos[row], channel); gap rows zero (f5_randn_rows)."""
    assert x.dtype == F32 and row_pos.dtype == I32 and row_utt.dtype == I32 and utt_seed.dtype == torch.int64
    call("f5_randn_rows", ptr(x), _ld(x), x.shape[0] if M is None else M, C_, ptr(row_pos), ptr(row_utt), ptr(utt_seed),
         stream_ptr())


def time_sinus(t, freqs, out, lo_off: int = 0) -> None:
    assert t.dtype == F32 and freqs.dtype == F32 and out.dtype == BF16
    call("f5_time_sinus", ptr(t), t.shape[0], ptr(freqs), 2 * freqs.shape[0], ptr(out), _ld(out), lo_off, stream_ptr())


def silu_bf16(x, out, split: bool = False) -> None:
    """out = silu(x); split: out is [rows, 2 * cols] = hi | lo planes of x [rows, cols]."""
    assert x.dtype == F32 and out.dtype == BF16 and x.is_contiguous() and out.is_contiguous()
    assert not split or out.shape == (x.shape[0], 2 * x.shape[1])
    call("f5_silu_bf16", ptr(x), ptr(out), x.numel(), x.shape[1] if split else 0, stream_ptr())


def istft(spec, window, frames, seg, max_wav_len: int, wav, gains=None) -> None:
    """spec fp32 [rows, >=1026]; frames fp32 [rows,1024] scratch; seg int32 [S,4]; wav fp32 flat."""
    assert spec.dtype == F32 and frames.dtype == F32 and seg.dtype == I32 and wav.dtype == F32 and window.dtype == F32
    call("f5_istft_frames", ptr(spec), _ld(spec), spec.shape[0], ptr(window), ptr(frames), stream_ptr())
    call("f5_istft_ola", ptr(frames), ptr(window), ptr(seg), seg.shape[0], max_wav_len, ptr(wav), ptr(gains), stream_ptr())


def mel_frames(wave: torch.Tensor, seg: torch.Tensor, max_frames: int, window: torch.Tensor, fbank: torch.Tensor,
               band: torch.Tensor, mel: torch.Tensor) -> None:
    """Prompt log-mel rows (f5_mel_frames): wave fp32 [samples], seg int32 [S,4], mel fp32 [rows, >= n_mels]."""
    assert wave.dtype == F32 and mel.dtype == F32 and seg.dtype == I32 and band.dtype == I32 and fbank.dtype == F32
    assert wave.is_contiguous() and seg.is_contiguous() and fbank.is_contiguous() and band.is_contiguous()
    call("f5_mel_frames", ptr(wave), ptr(seg), seg.shape[0], max_frames, ptr(window), ptr(fbank), ptr(band), fbank.shape[1],
         ptr(mel), _ld(mel), stream_ptr())
