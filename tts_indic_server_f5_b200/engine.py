"""B200-native CFM sampler + DiT backbone (the hot path of `CFM.sample`, reference `f5_tts/model/cfm.py:81-210`
and `f5_tts/model/backbones/dit.py:130-163`).

Design (B200-first, not a port):
  * all utterances of a batch are packed into one row matrix (layout.py); the conditional and unconditional CFG
    branches are the two halves of the SAME matrix, so every GEMM / attention launch covers the whole batch once;
  * everything that does not depend on the ODE state is hoisted out of the 32-step loop: the time embedding and ALL
    AdaLN modulation vectors for all steps and layers are one GEMM (t is a scalar shared by the batch,
    dit.py:141-142); the text embedding (ConvNeXtV2 stack, both CFG variants) and the step-invariant part of the
    input projection `W_c cond + W_t text + b` are computed once per batch; per step only `W_x x` (K=100) is new;
  * the residual stream, LayerNorm statistics, softmax, ODE state and time grid stay fp32; bf16 appears only as
    tensor-core operands;
  * the step loop (32 x (3 + 7*depth + 3) launches) is captured once per layout shape in a CUDA graph.
Every arithmetic op is a launcher of `libf5b200.so` (ops.py); torch provides device memory and streams only.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass

import torch

from . import _lib, ops
from .layout import PackedLayout, build_layout
from .weights import DiTConfig

BF16, F32, I32 = torch.bfloat16, torch.float32, torch.int32
MELP = 128  # mel channels padded to one 128-B bf16 swizzle row pair (K = 128)


def sway_time_grid(steps: int, sway_sampling_coef: float | None) -> torch.Tensor:
    """cfm.py:196-198 in fp32 (the fp32 grid is exact to 1 ulp; a bf16 grid would duplicate a point)."""
    t = torch.linspace(0, 1, steps + 1, dtype=torch.float32)
    if sway_sampling_coef is not None:
        t = t + sway_sampling_coef * (torch.cos(torch.pi / 2 * t) - 1 + t)
    return t


def _conv_pos_weight(w: torch.Tensor) -> torch.Tensor:
    """[D, cpg, K] grouped-conv weight -> [K*D, 64] per-tap K-major tiles of 64 input channels.  Groups narrower than
    64 channels are merged block-diagonally into 64-wide super-groups (exact; only tiny test configs need it)."""
    D, cpg, K = w.shape
    if cpg == 64:
        return w.permute(2, 0, 1).reshape(K * D, 64).contiguous()
    assert 64 % cpg == 0 and D % 64 == 0, "conv_pos group width must divide 64"
    wt = torch.zeros(K, D, 64, dtype=w.dtype)
    o = torch.arange(D)
    off = ((o // cpg) % (64 // cpg)) * cpg
    for j in range(cpg):
        wt[:, o, off + j] = w[:, j, :].t()
    return wt.reshape(K * D, 64).contiguous()


def split_planes(w: torch.Tensor) -> torch.Tensor:
    """fp32 weight [rows, K] -> the B operand of the split-operand GEMM mode: bf16 planes stacked by rows [hi | lo | hi] with
    hi = bf16(w), lo = bf16(w - hi) (3 x rows; for a conv weight the rows are (tap, out) so each plane holds all taps)."""
    w = w.float()
    hi = w.to(BF16)
    lo = (w - hi.float()).to(BF16)
    return torch.cat([hi, lo, hi])


class DiTWeights:
    """Device-resident weights in the layout the kernels consume (bf16 K-major GEMM operands, fp32 vectors)."""

    def __init__(self, sd: dict, cfg: DiTConfig, device, x3: bool = False):
        self.cfg, self.x3 = cfg, x3
        p = "transformer."
        D, TD, mel, L = cfg.dim, cfg.text_dim, cfg.mel_dim, cfg.depth
        assert D % 256 == 0 and cfg.dim_head == 64 and cfg.heads * 64 == D and TD % 128 == 0 and mel <= MELP

        def bf(t):
            return split_planes(t).to(device).contiguous() if x3 else t.to(device=device, dtype=BF16).contiguous()

        def f32(t):
            return t.to(device=device, dtype=F32).contiguous()

        g = lambda k: sd[p + k].float()  # noqa: E731
        self.t0_w, self.t0_b = bf(g("time_embed.time_mlp.0.weight")), f32(g("time_embed.time_mlp.0.bias"))
        self.t2_w, self.t2_b = bf(g("time_embed.time_mlp.2.weight")), f32(g("time_embed.time_mlp.2.bias"))
        half = cfg.freq_embed_dim // 2
        self.t_freqs = f32(torch.exp(torch.arange(half).float() * -(math.log(10000) / (half - 1))))  # modules.py:157-158
        mod_w = [g(f"transformer_blocks.{l}.attn_norm.linear.weight") for l in range(L)] + [g("norm_out.linear.weight")]
        mod_b = [g(f"transformer_blocks.{l}.attn_norm.linear.bias") for l in range(L)] + [g("norm_out.linear.bias")]
        self.mod_w, self.mod_b = bf(torch.cat(mod_w)), f32(torch.cat(mod_b))
        # text embedding
        self.emb = f32(g("text_embed.text_embed.weight"))
        freqs = 1.0 / (10000.0 ** (torch.arange(0, TD, 2)[: TD // 2].float() / TD))          # modules.py:196-207
        ang = torch.outer(torch.arange(cfg.max_pos), freqs).float()
        self.pos_table = f32(torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1))
        self.text_blocks = []
        for i in range(cfg.conv_layers):
            b = f"text_embed.text_blocks.{i}."
            self.text_blocks.append(dict(
                dw_w=f32(g(b + "dwconv.weight").reshape(TD, 7)), dw_b=f32(g(b + "dwconv.bias")),
                ln_w=f32(g(b + "norm.weight")), ln_b=f32(g(b + "norm.bias")),
                pw1_w=bf(g(b + "pwconv1.weight")), pw1_b=f32(g(b + "pwconv1.bias")),
                grn_g=f32(g(b + "grn.gamma").reshape(-1)), grn_b=f32(g(b + "grn.beta").reshape(-1)),
                pw2_w=bf(g(b + "pwconv2.weight")), pw2_b=f32(g(b + "pwconv2.bias"))))
        # input projection split by source (dit.py:85): [x | cond | text]
        W = g("input_embed.proj.weight")
        wx = torch.zeros(D, MELP)
        wx[:, :mel] = W[:, :mel]
        wc = torch.zeros(D, MELP)
        wc[:, :mel] = W[:, mel:2 * mel]
        self.wx = bf(wx)
        self.wct = bf(torch.cat([wc, W[:, 2 * mel:]], dim=1))
        self.proj_b = f32(g("input_embed.proj.bias"))
        self.conv_k = cfg.conv_pos_kernel
        self.c1_w = bf(_conv_pos_weight(g("input_embed.conv_pos_embed.conv1d.0.weight")))
        self.c1_b = f32(g("input_embed.conv_pos_embed.conv1d.0.bias"))
        self.c2_w = bf(_conv_pos_weight(g("input_embed.conv_pos_embed.conv1d.2.weight")))
        self.c2_b = f32(g("input_embed.conv_pos_embed.conv1d.2.bias"))
        self.blocks = []
        for l in range(L):
            b = f"transformer_blocks.{l}."
            self.blocks.append(dict(
                qkv_w=bf(torch.cat([g(b + "attn.to_q.weight"), g(b + "attn.to_k.weight"), g(b + "attn.to_v.weight")])),
                qkv_b=f32(torch.cat([g(b + "attn.to_q.bias"), g(b + "attn.to_k.bias"), g(b + "attn.to_v.bias")])),
                o_w=bf(g(b + "attn.to_out.0.weight")), o_b=f32(g(b + "attn.to_out.0.bias")),
                f1_w=bf(g(b + "ff.ff.0.0.weight")), f1_b=f32(g(b + "ff.ff.0.0.bias")),
                f2_w=bf(g(b + "ff.ff.2.weight")), f2_b=f32(g(b + "ff.ff.2.bias"))))
        po = torch.zeros(MELP, D)
        po[:mel] = g("proj_out.weight")
        pb = torch.zeros(MELP)
        pb[:mel] = g("proj_out.bias")
        self.out_w, self.out_b = bf(po), f32(pb)
        inv = 1.0 / (10000.0 ** (torch.arange(0, 64, 2).float() / 64))                        # x-transformers RotaryEmbedding
        ra = torch.arange(cfg.max_pos).float()[:, None] * inv[None]
        self.rope = f32(torch.stack((ra.cos(), ra.sin()), dim=-1).reshape(cfg.max_pos, 64))


@dataclass
class UtteranceInput:
    """One utterance after the `CFM.sample` prologue (cfm.py:100-149)."""
    cond: torch.Tensor        # fp32 [F, mel] prompt mel (CPU or CUDA)
    text_ids: torch.Tensor    # int64 [nt], vocabulary ids (no padding)
    n: int                    # total frames (duration after the max/clamp rules)
    cond_len: int             # frames where cond_mask is true (lens after max with text_lens)
    y0: torch.Tensor | None   # fp32 [n, mel] initial noise injected by the caller (parity tests); None: drawn on the device
    edit_mask: torch.Tensor | None = None   # bool [>= cond_len]
    noise_seed: int = 0       # 64-bit Philox key of the device draw when y0 is None (f5_randn_rows)


class Workspace:
    """Device buffers for one packed layout size; reused across batches of the same size."""

    def __init__(self, cfg: DiTConfig, R: int, steps_pad: int, device, x3: bool = False):
        D, TD, TI, FF, L = cfg.dim, cfg.text_dim, cfg.text_inner, cfg.ff_inner, cfg.depth
        P = 2 if x3 else 1                     # fp32 mode: every bf16 GEMM operand is two planes, hi | lo, side by side
        z = lambda r, c, dt: torch.zeros(r, c, device=device, dtype=dt)  # noqa: E731
        self.R = R
        self.x = z(R, MELP, F32)               # ODE state
        self.x0 = z(R, MELP, F32)              # initial noise y0 (kept so a staged batch can be re-run)
        self.cond = z(R, MELP, F32)            # step_cond
        self.xb = z(2 * R, P * MELP, BF16)     # bf16 copy of x for both CFG halves
        self.pred = z(2 * R, MELP, F32)
        self.xres = z(2 * R, D, F32)           # residual stream
        self.inv = z(2 * R, D, F32)            # W_c cond + W_t text + b
        self.hb = z(2 * R, P * D, BF16)
        self.ab = z(2 * R, P * D, BF16)        # conv-1 output / attention output
        if x3:                                 # fp32 scratch for GEMM outputs that feed another GEMM (QKV, FF1, conv-1, pointwise-1)
            self.s32 = z(2 * R, max(3 * D, FF, TI), F32)
        else:
            self.qkv = z(2 * R, 3 * D, BF16)
        self.fb = z(2 * R, P * FF, BF16)
        self.te = z(2 * R, TD, F32)
        self.tb = z(2 * R, P * TD, BF16)
        self.gb = z(2 * R, P * TI, BF16)
        self.act = z(2 * R, P * (MELP + TD), BF16)
        self.ids = torch.zeros(2 * R, device=device, dtype=I32)
        self.row_pos = torch.full((2 * R,), -1, device=device, dtype=I32)
        self.cond_flag = torch.zeros(R, device=device, dtype=I32)
        self.tsin = z(steps_pad, P * cfg.freq_embed_dim, BF16)
        self.th = z(steps_pad, D, F32)
        self.thb = z(steps_pad, P * D, BF16)
        self.mod = z(steps_pad, (6 * L + 2) * D, F32)
        self.tgrid = torch.zeros(steps_pad + 1, device=device, dtype=F32)
        self.dts = torch.zeros(steps_pad, device=device, dtype=F32)
        # attention work items live in ONE device buffer per workspace: the captured step graph holds its address and the
        # (padded) item count, so batches of other utterance lengths that fit the same rows replay the same graph
        self.tiles_buf = torch.zeros(2 * (R // 128 + 64), 4, device=device, dtype=I32)
        self.row_utt = torch.full((R,), -1, device=device, dtype=I32)
        self.graphs: dict[tuple, torch.cuda.CUDAGraph] = {}
        self.graph_launches: dict[tuple, int] = {}
        # Pinned host mirrors of the per-batch inputs, allocated once per workspace (cudaHostAlloc of tens of MB per batch was
        # a measurable part of the end-to-end step, and a fresh pinned buffer per batch keeps the host allocator busy while
        # the GPU runs).  `upload_done` marks the end of the last batch's H2D copies: the mirrors are rewritten only after it.
        self._pinned: dict[str, torch.Tensor] = {}
        self.upload_done: torch.cuda.Event | None = None
        self.generation = 0                     # bumped by every upload: a Staged batch is valid for ONE generation

    def pinned(self, name: str, shape: tuple, dtype) -> torch.Tensor:
        t = self._pinned.get(name)
        if t is None or t.shape != tuple(shape) or t.dtype != dtype:
            t = torch.zeros(*shape, dtype=dtype).pin_memory()
            self._pinned[name] = t
        return t


class F5Engine:
    """CUDA sampler for one set of DiT weights.  `sample_packed` is the device-resident hot path."""

    def __init__(self, sd: dict, cfg: DiTConfig, device="cuda", use_graphs: bool = True, precision: str = "bf16"):
        """precision: "bf16" (served path: bf16 tensor-core operands, fp32 everything else) or "fp32" (the reference as
        deployed is fp32, core/managers.py:76): every GEMM runs in the split-operand mode — operands carried as hi + lo bf16
        planes, three tcgen05 products per k-block — and attention runs in fp32 on the CUDA cores (attn_f32.cu)."""
        from ._lib import lib
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.precision, self.x3 = precision, precision == "fp32"
        if not torch.cuda.is_available():
            raise RuntimeError("F5Engine needs a CUDA device (sm_100a); there is no CPU fallback")
        dev = torch.device(device)
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        torch.cuda.set_device(dev)
        rc = lib.f5_device_check()
        if rc != 0:
            raise RuntimeError("libf5b200.so targets sm_100a (B200) only")
        self.cfg, self.device = cfg, dev
        self.w = DiTWeights(sd, cfg, self.device, self.x3)
        self.use_graphs = use_graphs
        self._ws: dict[tuple, Workspace] = {}        # LRU over (row count, slot) (insertion order = recency)
        self.max_workspaces = 4                       # ~30 KB of buffers per row: C2 (159 k rows) is 4.7 GB
        self.max_graphs_per_workspace = 8
        self.pdl_max_rows = 16384                     # packed rows (incl. the CFG duplicate) up to which PDL is switched on
        self.pdl_auto = os.environ.get("F5_PDL") is None   # an explicit F5_PDL=0/1 (A/B runs) or a test overrides the policy
        self.graph_captures = 0                       # statistics: captures vs replays of the step graph
        self.graph_replays = 0

    # ------------------------------------------------------------------------------------------ batch set-up
    def workspace(self, R: int, slot: int = 0) -> Workspace:
        """Buffers for a packed batch of R rows.  `slot` separates batches of the SAME size that must be resident at the same
        time (several staged packs of one request batch); a stage into an occupied (R, slot) takes the workspace over."""
        ws = self._ws.pop((R, slot), None)
        if ws is None:
            while len(self._ws) >= self.max_workspaces:            # evict the least recently used size (and its graphs)
                self._ws.pop(next(iter(self._ws)))
            ws = Workspace(self.cfg, R, 128, self.device, self.x3)
        self._ws[(R, slot)] = ws
        return ws

    def upload(self, utts: list[UtteranceInput], layout: PackedLayout, steps: int, sway: float | None, slot: int = 0) -> Workspace:
        """Stage the per-batch inputs through the workspace's pinned host mirrors and copy them to the device (H2D on the
        current stream).  The initial noise is drawn on the device (Philox, `f5_randn_rows`) unless the caller injects `y0`;
        a prompt mel that is already on the device never visits the host."""
        cfg, R = self.cfg, layout.half_rows
        mel = cfg.mel_dim
        if steps > 128:
            raise ValueError(f"steps={steps}: the hoisted time/AdaLN tables hold at most 128 Euler steps")
        if max(layout.lengths) > cfg.max_pos:
            raise ValueError(f"an utterance of {max(layout.lengths)} frames exceeds max_pos={cfg.max_pos} (cfm.py:137 clamps at 4096)")
        emb_rows = self.w.emb.shape[0]
        for u in utts:                                               # nn.Embedding would raise IndexError (dit.py:56)
            if u.text_ids.numel() and int(u.text_ids.max()) + 1 >= emb_rows:
                raise IndexError(f"token id {int(u.text_ids.max())} is outside the text embedding ({emb_rows - 1} entries)")
        ws = self.workspace(R, slot)
        if ws.upload_done is not None:
            ws.upload_done.synchronize()                             # the previous batch's H2D copies have left the mirrors
        ws.generation += 1
        inject = any(u.y0 is not None for u in utts)
        if inject and not all(u.y0 is not None for u in utts):
            raise ValueError("y0 must be given for every utterance of a batch or for none")
        host_cond = any(not (u.cond.is_cuda and u.edit_mask is None) for u in utts)
        h_ids, h_flag = ws.pinned("ids", (2 * R,), I32), ws.pinned("flag", (R,), I32)
        h_ids.zero_()
        h_flag.zero_()
        h_x = h_cond = None
        if inject:
            h_x = ws.pinned("x0", (R, MELP), F32)
            h_x.zero_()
        if host_cond:
            h_cond = ws.pinned("cond", (R, MELP), F32)
            h_cond.zero_()
        dev_conds = []
        for u, s in zip(utts, layout.starts):
            n = u.n
            if inject:
                h_x[s:s + n, :mel] = u.y0[:n].float().cpu()
            F_ = min(u.cond.shape[0], n)
            mask = torch.zeros(n, dtype=torch.bool)
            mask[: min(u.cond_len, n)] = True
            if u.edit_mask is not None:
                em = u.edit_mask.cpu().bool()[:n]
                mask[: em.numel()] &= em
            h_flag[s:s + n] = mask.to(I32)
            if u.cond.is_cuda and u.edit_mask is None:
                dev_conds.append((s, min(F_, u.cond_len), u.cond))       # prompt mel already on the device: no host round trip
            else:
                c = torch.zeros(n, mel)
                c[:F_] = u.cond[:F_].float().cpu()
                h_cond[s:s + n, :mel] = torch.where(mask[:, None], c, torch.zeros_like(c))
            nt = min(u.text_ids.numel(), n)                      # dit.py:48-51: +1, truncate to n, filler 0
            h_ids[s:s + nt] = (u.text_ids[:nt] + 1).to(I32)
        h_pos = ws.pinned("row_pos", (2 * R,), I32)
        h_pos.copy_(layout.row_pos)
        h_utt = ws.pinned("row_utt", (R,), I32)
        h_utt.copy_(layout.row_utt)
        ws.row_pos.copy_(h_pos, non_blocking=True)
        ws.row_utt.copy_(h_utt, non_blocking=True)
        if inject:
            ws.x0.copy_(h_x, non_blocking=True)
        else:
            h_seed = ws.pinned("seeds", (max(len(utts), 1),), torch.int64)
            h_seed[: len(utts)] = torch.tensor([((u.noise_seed + (1 << 63)) % (1 << 64)) - (1 << 63) for u in utts], dtype=torch.int64)
            ws.seeds = h_seed.to(self.device, non_blocking=True)
            ops.randn_rows(ws.x0, mel, ws.row_pos, ws.row_utt, ws.seeds, M=R)      # cfm.py:181-186 on the device
        if host_cond:
            ws.cond.copy_(h_cond, non_blocking=True)
        else:
            ws.cond.zero_()
        for s, f, c in dev_conds:
            ws.cond[s:s + f, :mel].copy_(c[:f])
        ws.ids.copy_(h_ids, non_blocking=True)
        ws.cond_flag.copy_(h_flag, non_blocking=True)
        n_items = layout.attn_tiles.shape[0]
        padded = (n_items + 7) // 8 * 8                # item count is baked into the step graph: pad it to a multiple of 8 ...
        if padded > ws.tiles_buf.shape[0]:
            ws.tiles_buf = torch.zeros(padded + 64, 4, device=self.device, dtype=I32)
            ws.graphs.clear()
        h_tiles = ws.pinned("tiles", (ws.tiles_buf.shape[0], 4), I32)
        h_tiles.zero_()                                            # ... with items of zero query rows, which the kernel skips
        h_tiles[:n_items] = layout.attn_tiles
        ws.tiles_buf[:padded].copy_(h_tiles[:padded], non_blocking=True)
        ws.tiles = ws.tiles_buf[:padded]
        nseg = layout.seg_rows.shape[0]
        h_seg = ws.pinned("segs", (nseg, 2), I32)
        h_seg.copy_(layout.seg_rows)
        ws.segs = h_seg.to(self.device, non_blocking=True)
        ws.sumsq = torch.zeros(nseg, cfg.text_inner, device=self.device, dtype=F32)
        t = sway_time_grid(steps, sway)
        h_t = ws.pinned("tgrid", (ws.tgrid.shape[0],), F32)
        h_t.zero_()
        h_t[: steps + 1] = t
        h_dt = ws.pinned("dts", (ws.dts.shape[0],), F32)
        h_dt.zero_()
        h_dt[:steps] = t[1:] - t[:-1]                          # torchdiffeq Euler: dt = t1 - t0 in the grid dtype
        ws.tgrid.copy_(h_t, non_blocking=True)
        ws.dts.copy_(h_dt, non_blocking=True)
        ws.upload_done = torch.cuda.Event()
        ws.upload_done.record()
        moved = [h_ids, h_flag, h_pos, h_utt, h_tiles[:padded], h_seg, h_t, h_dt] + ([h_x] if inject else []) + ([h_cond] if host_cond else [])
        ws.h2d_bytes = sum(v.numel() * v.element_size() for v in moved) + (0 if inject else 8 * len(utts))
        return ws

    # ------------------------------------------------------------------------------------------ precision-aware building blocks
    def _gemm_to_operand(self, ws: Workspace, A, W, out_b, N: int, *, act=ops.F5_ACT_NONE, bias=None, mask_rows=False, **kw) -> None:
        """out_b = act(A W^T + bias) as the NEXT GEMM's operand.  bf16 mode: the tensor-core GEMM stores bf16 itself.  fp32 mode:
        the split-operand GEMM stores fp32 into the scratch and `f5_pack_bf16` splits it into the hi | lo planes."""
        if not self.x3:
            ops.gemm(A, W, N=N, mode=ops.F5_EPI_STORE_BF16, act=act, bias=bias, out=out_b, row_pos=ws.row_pos if mask_rows else None,
                     mask_rows=mask_rows, **kw)
            return
        s32 = ws.s32[:, :N]
        ops.gemm(A, W, N=N, mode=ops.F5_EPI_STORE_F32, act=act, bias=bias, out=s32, split=True, **kw)
        ops.pack_bf16(s32, out_b, 0, N, N, row_pos=ws.row_pos if mask_rows else None, lo_off=N)

    def _lo(self, width: int) -> int:
        return width if self.x3 else 0

    # ------------------------------------------------------------------------------------------ hoisted work
    def hoist(self, ws: Workspace, steps: int) -> None:
        """Step-invariant work: time/AdaLN vectors for all steps, text embedding (both CFG variants), W_c cond + W_t text + b."""
        cfg, w, R, x3 = self.cfg, self.w, ws.R, self.x3
        D, TD, TI = cfg.dim, cfg.text_dim, cfg.text_inner
        # --- time embedding + all modulation vectors (modules.py:648-658, :286, :307)
        ops.time_sinus(ws.tgrid[:128], w.t_freqs, ws.tsin, lo_off=self._lo(cfg.freq_embed_dim))
        ops.gemm(ws.tsin, w.t0_w, mode=ops.F5_EPI_STORE_F32, bias=w.t0_b, out=ws.th, split=x3)
        ops.silu_bf16(ws.th, ws.thb, split=x3)
        ops.gemm(ws.thb, w.t2_w, mode=ops.F5_EPI_STORE_F32, bias=w.t2_b, out=ws.th, split=x3)
        ops.silu_bf16(ws.th, ws.thb, split=x3)
        ops.gemm(ws.thb, w.mod_w, mode=ops.F5_EPI_STORE_F32, bias=w.mod_b, out=ws.mod, split=x3)
        # --- text embedding, conditional half = real tokens, unconditional half = all filler (dit.py:47-69)
        ops.text_gather_pos(ws.ids, ws.row_pos, w.emb, w.pos_table, ws.te)
        for blk in w.text_blocks:                                            # ConvNeXtV2Block, modules.py:259-269
            ops.dwconv7_ln(ws.te, ws.tb, ws.row_pos, blk["dw_w"], blk["dw_b"], blk["ln_w"], blk["ln_b"], lo_off=self._lo(TD))
            if not x3:
                ops.gemm(ws.tb, blk["pw1_w"], mode=ops.F5_EPI_STORE_BF16, act=ops.F5_ACT_GELU_ERF, bias=blk["pw1_b"], out=ws.gb)
                ops.grn(ws.gb, ws.segs, ws.sumsq, blk["grn_g"], blk["grn_b"])
            else:                                                            # GELU and GRN in fp32, then split for pwconv2
                s32 = ws.s32[:, :TI]
                ops.gemm(ws.tb, blk["pw1_w"], mode=ops.F5_EPI_STORE_F32, act=ops.F5_ACT_GELU_ERF, bias=blk["pw1_b"], out=s32, split=True)
                ops.grn_f32(s32, ws.segs, ws.sumsq, blk["grn_g"], blk["grn_b"], TI)
                ops.pack_bf16(s32, ws.gb, 0, TI, TI, lo_off=TI)
            ops.gemm(ws.gb, blk["pw2_w"], mode=ops.F5_EPI_RESID_F32, bias=blk["pw2_b"], resid=ws.te, split=x3)
        # --- A = [cond | text] (cond is zero in the unconditional half: drop_audio_cond, dit.py:82-83)
        ops.pack_bf16(ws.cond, ws.act, 0, cfg.mel_dim, MELP, M=R, lo_off=self._lo(MELP + TD))
        ops.pack_bf16(ws.te, ws.act, MELP, TD, TD, row_pos=ws.row_pos, lo_off=self._lo(MELP + TD))
        ops.gemm(ws.act, w.wct, mode=ops.F5_EPI_STORE_F32, bias=w.proj_b, out=ws.inv, split=x3)
        # --- bf16 copy of the initial state for both halves
        self._pack_state(ws)

    def _pack_state(self, ws: Workspace) -> None:
        R = ws.R
        ops.pack_bf16(ws.x, ws.xb, 0, self.cfg.mel_dim, MELP, row_pos=ws.row_pos, M=R, lo_off=self._lo(MELP))
        ops.pack_bf16(ws.x, ws.xb[R:], 0, self.cfg.mel_dim, MELP, row_pos=ws.row_pos, M=R, lo_off=self._lo(MELP))

    # ------------------------------------------------------------------------------------------ one Euler step
    def step(self, ws: Workspace, s: int, cfg_strength: float) -> None:
        cfg, w, R, x3 = self.cfg, self.w, ws.R, self.x3
        D, L, H, FF = cfg.dim, cfg.depth, cfg.heads, cfg.ff_inner
        mod = ws.mod[s]
        conv = dict(M=2 * R, block_n=64, num_taps=w.conv_k, kc_per_tap=1, tap_pad=w.conv_k // 2, a_grouped=True, b_tap_rows=D)
        # input embedding: W_x x + (W_c cond + W_t text + b); conv position embedding + residual (dit.py:85-86)
        if not x3:
            ops.gemm(ws.xb, w.wx, mode=ops.F5_EPI_STORE_F32, out=ws.xres, addend=ws.inv, out2=ws.hb, row_pos=ws.row_pos,
                     mask_rows=True)
        else:
            ops.gemm(ws.xb, w.wx, mode=ops.F5_EPI_STORE_F32, out=ws.xres, addend=ws.inv, split=True)
            ops.pack_bf16(ws.xres, ws.hb, 0, D, D, row_pos=ws.row_pos, lo_off=D)
        self._gemm_to_operand(ws, ws.hb, w.c1_w, ws.ab, D, act=ops.F5_ACT_MISH, bias=w.c1_b, mask_rows=True, **conv)
        ops.gemm(ws.ab, w.c2_w, N=D, mode=ops.F5_EPI_RESID_F32, act=ops.F5_ACT_MISH, bias=w.c2_b, resid=ws.xres, split=x3, **conv)
        for l, blk in enumerate(w.blocks):                                   # DiTBlock, modules.py:558-572
            m = mod[l * 6 * D:(l + 1) * 6 * D]
            shift_msa, scale_msa, gate_msa = m[0:D], m[D:2 * D], m[2 * D:3 * D]
            shift_mlp, scale_mlp, gate_mlp = m[3 * D:4 * D], m[4 * D:5 * D], m[5 * D:6 * D]
            ops.layernorm_mod(ws.xres, ws.hb, scale_msa, shift_msa, 1.0, lo_off=self._lo(D))
            if not x3:
                ops.gemm(ws.hb, blk["qkv_w"], mode=ops.F5_EPI_STORE_BF16, bias=blk["qkv_b"], out=ws.qkv, row_pos=ws.row_pos,
                         rope=w.rope, rope_period=D, rope_tiles=2)
                ops.attention(ws.qkv, ws.tiles, ws.ab, H, 0, D, 2 * D, 0.125)
            else:                                                            # fp32 Q / K / V, RoPE inside the fp32 attention kernel
                qkv32 = ws.s32[:, :3 * D]
                ops.gemm(ws.hb, blk["qkv_w"], mode=ops.F5_EPI_STORE_F32, bias=blk["qkv_b"], out=qkv32, split=True)
                ops.attention_f32(qkv32, ws.tiles, ws.ab, H, 0, D, 2 * D, 0.125, rope=w.rope, lo_off=D)
            ops.gemm(ws.ab, blk["o_w"], mode=ops.F5_EPI_RESID_F32, bias=blk["o_b"], gate=gate_msa, resid=ws.xres, split=x3)
            ops.layernorm_mod(ws.xres, ws.hb, scale_mlp, shift_mlp, 1.0, lo_off=self._lo(D))
            self._gemm_to_operand(ws, ws.hb, blk["f1_w"], ws.fb, FF, act=ops.F5_ACT_GELU_TANH, bias=blk["f1_b"])
            ops.gemm(ws.fb, blk["f2_w"], mode=ops.F5_EPI_RESID_F32, bias=blk["f2_b"], gate=gate_mlp, resid=ws.xres, split=x3)
        mf = mod[6 * L * D:]
        ops.layernorm_mod(ws.xres, ws.hb, mf[0:D], mf[D:2 * D], 1.0, lo_off=self._lo(D))          # AdaLayerNormZero_Final: (scale, shift)
        ops.gemm(ws.hb, w.out_w, mode=ops.F5_EPI_STORE_F32, bias=w.out_b, out=ws.pred, split=x3)
        ops.cfg_euler(ws.x, ws.pred, R, cfg.mel_dim, ws.row_pos, ws.dts, s, cfg_strength, ws.xb, MELP)
        if x3:
            self._pack_state(ws)                                             # both planes of the next step's W_x operand

    def run_steps(self, ws: Workspace, steps: int, cfg_strength: float) -> None:
        # The graph bakes in addresses of this workspace's buffers and the launch parameters: row count (the workspace), padded
        # attention item count, steps, cfg.  Everything else a batch changes — lengths, positions, tile table, noise, prompt,
        # text — is CONTENT of those buffers, so any batch that packs into the same rows replays the same graph.
        key = (ws.tiles.shape[0], steps, float(cfg_strength), ws.tiles.data_ptr())
        if not self.use_graphs:
            for s in range(steps):
                self.step(ws, s, cfg_strength)
            return
        g = ws.graphs.pop(key, None)
        if g is None:
            x_saved = ws.x.clone()
            self.step(ws, 0, cfg_strength)            # eager warm-up of every kernel variant (sets func attributes) ...
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count
            with torch.cuda.graph(g):                 # capture records, it does not execute
                for s in range(steps):
                    self.step(ws, s, cfg_strength)
            ws.graph_launches[key] = _lib.launch_count - n0
            _lib.launch_count = n0
            ws.x.copy_(x_saved)                       # ... then restore the state the warm-up step advanced
            self._pack_state(ws)
            while len(ws.graphs) >= self.max_graphs_per_workspace:
                ws.graphs.pop(next(iter(ws.graphs)))
            self.graph_captures += 1
        else:
            self.graph_replays += 1
        ws.graphs[key] = g                            # most recently used last
        g.replay()
        _lib.launch_count += ws.graph_launches[key]

    # ------------------------------------------------------------------------------------------ public
    @torch.inference_mode()
    def stage(self, utts: list[UtteranceInput], steps: int = 32, sway_sampling_coef: float | None = -1.0, slot: int = 0):
        """Host -> device: build the packed layout and copy this batch's inputs (pinned H2D).  No arithmetic."""
        layout = build_layout([u.n for u in utts])
        ws = self.upload(utts, layout, steps, sway_sampling_coef, slot)
        return ws, layout

    @torch.inference_mode()
    def compute(self, ws: Workspace, steps: int = 32, cfg_strength: float = 2.0, generation: int | None = None,
                on_state=None) -> None:
        """Device-resident hot path on a staged batch: hoisted work, the Euler loop, prompt re-insert (cfm.py:160-204).
        `on_state(k, ws)` is called with the ODE state after k = 0 .. steps Euler steps (the reference's trajectory,
        cfm.py:200); it forces eager launches (the served path never asks for it and replays the step graph)."""
        if cfg_strength < 1e-5:
            raise NotImplementedError("cfg_strength < 1e-5 (single-branch sampling, cfm.py:170-171) is not on the served path")
        # Programmatic dependent launch pays where kernels are short (a single request: 5120 graph nodes of ~10 us, -3 % latency);
        # on a full batch the same edges cost 4 % (measured same-box, profiles/README.md), so it is switched per batch size.
        # The setting is baked into a step graph when it is captured, and graphs are keyed per workspace (= per size).
        if self.pdl_auto:
            _lib.lib.f5_set_pdl(1 if 2 * ws.R <= self.pdl_max_rows else 0)
        if generation is not None and generation != ws.generation:
            raise RuntimeError("this staged batch was overwritten: another batch of the same packed size was staged into its "
                               "workspace before it ran (stage -> run must not interleave with another stage of the same size)")
        ws.x.copy_(ws.x0)
        self.hoist(ws, steps)
        if on_state is None:
            self.run_steps(ws, steps, cfg_strength)
        else:
            on_state(0, ws)
            for s in range(steps):
                self.step(ws, s, cfg_strength)
                on_state(s + 1, ws)
        ops.where_rows(ws.x, ws.cond, ws.cond_flag, self.cfg.mel_dim)        # cfm.py:204

    def sample_packed(self, utts: list[UtteranceInput], steps: int = 32, cfg_strength: float = 2.0,
                      sway_sampling_coef: float | None = -1.0, on_state=None) -> tuple[Workspace, PackedLayout]:
        """Run the sampler for a batch; the result stays on the device in `ws.x` (rows per `layout`)."""
        ws, layout = self.stage(utts, steps, sway_sampling_coef)
        self.compute(ws, steps, cfg_strength, on_state=on_state)
        return ws, layout

    @torch.inference_mode()
    def forward_flow(self, utts: list[UtteranceInput], t: float) -> list[torch.Tensor]:
        """One CFG velocity evaluation pair at time t (test hook): returns per utterance [2, n, mel] (cond, null)."""
        layout = build_layout([u.n for u in utts])
        ws = self.upload(utts, layout, 1, None)
        ws.tgrid[0] = t
        ws.x.copy_(ws.x0)
        self.hoist(ws, 1)
        graphs, self.use_graphs = self.use_graphs, False
        ws.dts.zero_()
        self.step(ws, 0, 2.0)
        self.use_graphs = graphs
        torch.cuda.synchronize()
        R, mel = ws.R, self.cfg.mel_dim
        return [torch.stack((ws.pred[s:s + n, :mel], ws.pred[R + s:R + s + n, :mel])).clone()
                for s, n in zip(layout.starts, layout.lengths)]
