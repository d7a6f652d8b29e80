"""B200-native Vocos decoder (vocos 0.1.0 `Vocos.decode` = ISTFTHead(VocosBackbone(mel)); reference call site
`f5_tts/infer/utils_infer.py:472`, loader `:92-115`).

Packed layout like the DiT engine (one row per mel frame, utterances separated by GAP zero rows so the k=7 convolutions
see the reference's zero padding).  The embed conv runs as a 7-tap implicit GEMM on tcgen05, the pointwise convs as
tcgen05 GEMMs with fused GELU / gamma*y+residual epilogues, dwconv+LayerNorm fused, ISTFT as an fp32 shared-memory FFT
with fused window / overlap-add / envelope normalisation.  The residual stream and the ISTFT are fp32.
"""
from __future__ import annotations

import torch

from . import ops
from .weights import VocosConfig

BF16, F32, I32 = torch.bfloat16, torch.float32, torch.int32
MELP = 128
VGAP = 8


class VocosEngine:
    def __init__(self, vsd: dict, vcfg: VocosConfig, device="cuda", precision: str = "bf16"):
        """precision "fp32": split-operand GEMMs (hi + lo bf16 planes, three products per k-block), see engine.F5Engine."""
        from ._lib import lib
        from .engine import split_planes
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.precision, self.x3 = precision, precision == "fp32"
        x3 = self.x3
        if not torch.cuda.is_available() or lib.f5_device_check() != 0:
            raise RuntimeError("VocosEngine needs a B200 (sm_100a); there is no CPU fallback")
        device = torch.device(device)
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.cfg, self.device = vcfg, device
        C, I = vcfg.dim, vcfg.intermediate_dim
        assert vcfg.n_fft == 1024 and vcfg.hop == 256 and C % 128 == 0 and C <= 512 and vcfg.n_mels <= MELP
        bf = lambda t: (split_planes(t).to(device).contiguous() if x3 else t.to(device=device, dtype=BF16).contiguous())  # noqa: E731
        f32 = lambda t: t.to(device=device, dtype=F32).contiguous()      # noqa: E731
        w = vsd["backbone.embed.weight"].float()                          # [C, n_mels, 7]
        wt = torch.zeros(7, C, MELP)
        wt[:, :, : vcfg.n_mels] = w.permute(2, 0, 1)
        self.emb_w, self.emb_b = bf(wt.reshape(7 * C, MELP)), f32(vsd["backbone.embed.bias"])
        self.norm_w, self.norm_b = f32(vsd["backbone.norm.weight"]), f32(vsd["backbone.norm.bias"])
        self.blocks = []
        for i in range(vcfg.num_layers):
            p = f"backbone.convnext.{i}."
            self.blocks.append(dict(
                dw_w=f32(vsd[p + "dwconv.weight"].reshape(C, 7)), dw_b=f32(vsd[p + "dwconv.bias"]),
                ln_w=f32(vsd[p + "norm.weight"]), ln_b=f32(vsd[p + "norm.bias"]),
                pw1_w=bf(vsd[p + "pwconv1.weight"]), pw1_b=f32(vsd[p + "pwconv1.bias"]),
                pw2_w=bf(vsd[p + "pwconv2.weight"]), pw2_b=f32(vsd[p + "pwconv2.bias"]), gamma=f32(vsd[p + "gamma"])))
        self.fin_w, self.fin_b = f32(vsd["backbone.final_layer_norm.weight"]), f32(vsd["backbone.final_layer_norm.bias"])
        nout = vcfg.n_fft + 2
        self.nout_pad = (nout + 7) // 8 * 8                                  # 1026 -> 1032 (N % 8 == 0)
        hw = torch.zeros(self.nout_pad, C)
        hw[:nout] = vsd["head.out.weight"].float()
        hb = torch.zeros(self.nout_pad)
        hb[:nout] = vsd["head.out.bias"].float()
        self.head_w, self.head_b = bf(hw), f32(hb)
        self.window = f32(vsd["head.istft.window"])
        self.spec_ld = (self.nout_pad + 127) // 128 * 128
        self._bufs: dict[int, dict] = {}
        self.max_buffer_sets = 4

    def _buffers(self, Rv: int) -> dict:
        """Scratch for Rv vocoder rows; an LRU of a few sizes (a server sees a handful of length buckets: re-allocating and
        zero-filling ~13 KB per row on every new length was a measurable part of a single-utterance request)."""
        b = self._bufs.pop(Rv, None)
        if b is None:
            C, I = self.cfg.dim, self.cfg.intermediate_dim
            z = lambda r, c, dt: torch.zeros(r, c, device=self.device, dtype=dt)  # noqa: E731
            P = 2 if self.x3 else 1                                          # fp32 mode: hi | lo planes side by side
            b = dict(melb=z(Rv, P * MELP, BF16), h=z(Rv, C, F32), v=z(Rv, C, F32), hb=z(Rv, P * C, BF16), ib=z(Rv, P * I, BF16),
                     spec=z(Rv, self.spec_ld, F32), frames=z(Rv, self.cfg.n_fft, F32))
            if self.x3:
                b["s32"] = z(Rv, I, F32)
            while len(self._bufs) >= self.max_buffer_sets:
                self._bufs.pop(next(iter(self._bufs)))
        self._bufs[Rv] = b                                                  # most recently used last
        return b

    @staticmethod
    def plan(frames: list[int]):
        """Row layout for utterances of `frames` mel frames each: starts, padded rows, row_pos, wav offsets."""
        starts, r = [], VGAP
        for T in frames:
            starts.append(r)
            r += T + VGAP
        Rv = (r + 127) // 128 * 128
        pos = torch.full((Rv,), -1, dtype=I32)
        for s, T in zip(starts, frames):
            pos[s:s + T] = torch.arange(T, dtype=I32)
        offs, tot = [], 0
        for T in frames:
            offs.append(tot)
            tot += 256 * max(T - 1, 0)
        return starts, Rv, pos, offs, tot

    @torch.inference_mode()
    def decode_rows(self, src: torch.Tensor, src_rows: torch.Tensor, row_pos: torch.Tensor, seg: torch.Tensor,
                    frames: list[int], total: int, gains: torch.Tensor | None = None) -> torch.Tensor:
        """src fp32 [*, >=n_mels] device mel rows; src_rows int32 [Rv] maps vocoder rows to src rows (-1 = zero row);
        seg int32 [S,4] = {row0, frames, wav_offset, 0} on the device.  Returns the flat fp32 waveform buffer
        (utterance i at its offset, 256*(frames[i]-1) samples).  Device-resident: no host copies."""
        cfg, Rv, x3 = self.cfg, src_rows.shape[0], self.x3
        b = self._buffers(Rv)
        C, I = cfg.dim, cfg.intermediate_dim
        lo = (lambda w: w) if x3 else (lambda w: 0)                          # low-plane offset of a w-wide operand
        ops.pack_bf16(src, b["melb"], 0, cfg.n_mels, MELP, src_rows=src_rows, lo_off=lo(MELP))
        ops.gemm(b["melb"], self.emb_w, M=Rv, N=C, mode=ops.F5_EPI_STORE_F32, bias=self.emb_b, out=b["h"], num_taps=7,
                 kc_per_tap=MELP // 64, tap_pad=3, b_tap_rows=C, split=x3)
        ops.layernorm_mod(b["h"], None, self.norm_w, self.norm_b, 0.0, y32=b["v"])
        for blk in self.blocks:
            ops.dwconv7_ln(b["v"], b["hb"], row_pos, blk["dw_w"], blk["dw_b"], blk["ln_w"], blk["ln_b"], lo_off=lo(C))
            if not x3:
                ops.gemm(b["hb"], blk["pw1_w"], mode=ops.F5_EPI_STORE_BF16, act=ops.F5_ACT_GELU_ERF, bias=blk["pw1_b"], out=b["ib"])
            else:
                ops.gemm(b["hb"], blk["pw1_w"], mode=ops.F5_EPI_STORE_F32, act=ops.F5_ACT_GELU_ERF, bias=blk["pw1_b"], out=b["s32"], split=True)
                ops.pack_bf16(b["s32"], b["ib"], 0, I, I, lo_off=I)
            ops.gemm(b["ib"], blk["pw2_w"], mode=ops.F5_EPI_RESID_F32, bias=blk["pw2_b"], gate=blk["gamma"], resid=b["v"], split=x3)
        ops.layernorm_mod(b["v"], b["hb"], self.fin_w, self.fin_b, 0.0, lo_off=lo(C))
        ops.gemm(b["hb"], self.head_w, mode=ops.F5_EPI_STORE_F32, bias=self.head_b, out=b["spec"], block_n=128, split=x3)
        wav = torch.empty(max(total, 1), device=self.device, dtype=F32)
        ops.istft(b["spec"], self.window, b["frames"], seg, 256 * max(max(frames) - 1, 1), wav, gains)
        return wav

    @torch.inference_mode()
    def decode(self, mel: torch.Tensor) -> torch.Tensor:
        """`Vocos.decode` surface: mel [b, n_mels, T] (any device) -> wav [b, 256*(T-1)] fp32 on the engine's device."""
        assert mel.dim() == 3 and mel.shape[1] == self.cfg.n_mels
        B, _, T = mel.shape
        src = mel.to(self.device, F32).permute(0, 2, 1).reshape(B * T, self.cfg.n_mels).contiguous()
        starts, Rv, pos, offs, tot = self.plan([T] * B)
        src_rows = torch.full((Rv,), -1, dtype=I32)
        for i, s in enumerate(starts):
            src_rows[s:s + T] = torch.arange(i * T, (i + 1) * T, dtype=I32)
        seg = torch.tensor([[s, T, o, 0] for s, o in zip(starts, offs)], dtype=I32).to(self.device)
        wav = self.decode_rows(src, src_rows.to(self.device), pos.to(self.device), seg, [T] * B, tot)
        return wav[:tot].view(B, 256 * (T - 1))
