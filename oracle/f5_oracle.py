"""ORACLE — test infrastructure only.  CPU fp32 restatement of the reference's F5-TTS inference path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s baseline legs (`cpu_baseline`, `--impl reference`, and the
same-box torch-eager `gpu_eager_baseline` VERDICT r01 asked for) may import this module — as the checker or as the thing a
baseline times, never as part of the product; `tts_indic_server_f5_b200/` never imports it.

Each function follows the reference file:line it cites (paths relative to
`/root/reference/src/server/f5_tts/`).  Three pieces of arithmetic live in third-party packages that are
NOT vendored in the reference tree and are not installed here; they are restated from their published
algorithms and are "parity unpinned" by any reference test (the reference has no tests at all):
  * x-transformers==2.2.8  RotaryEmbedding / apply_rotary_pos_emb   (call sites model/modules.py:418-419)
  * torchdiffeq==0.2.5     odeint(method="euler")                   (call site  model/cfm.py:200)
  * vocos==0.1.0           VocosBackbone + ISTFTHead                (call site  infer/utils_infer.py:472)
Everything that IS in the tree (CFM.sample, DiT, modules) is pinned by `tests/test_oracle_golden.py::test_oracle_vs_real_reference_modules`,
which imports the real reference modules in place (through `oracle/ref_shims.py`) when `/root/reference`
exists, and by the golden vectors under `tests/golden/` generated from the real reference by
`oracle/make_golden.py`.

Semantics: batch-1 per utterance (what the server computes, `infer/utils_infer.py:441-466`; the
reference's padded batched mode is not self-consistent).
All functions are functional over a state dict with the reference's key names.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# third-party restatements
# ----------------------------------------------------------------------------------------------


def rotary_freqs(seq_len: int, dim_head: int = 64) -> torch.Tensor:
    """x-transformers 2.2.8 `RotaryEmbedding(dim).forward_from_seq_len(n)` -> freqs [1, n, dim]
    (interleaved duplicate: [f0,f0,f1,f1,...]); xpos scale is 1.0.  Call site model/backbones/dit.py:117,149."""
    inv_freq = 1.0 / (10000.0 ** (torch.arange(0, dim_head, 2).float() / dim_head))
    t = torch.arange(seq_len).float()
    freqs = torch.einsum("i,j->ij", t, inv_freq)
    freqs = torch.stack((freqs, freqs), dim=-1).reshape(seq_len, dim_head)
    return freqs[None]


def _rotate_half(x: torch.Tensor) -> torch.Tensor:
    x = x.reshape(*x.shape[:-1], x.shape[-1] // 2, 2)
    x1, x2 = x.unbind(dim=-1)
    return torch.stack((-x2, x1), dim=-1).reshape(*x.shape[:-2], -1)


def apply_rotary_pos_emb(t: torch.Tensor, freqs: torch.Tensor, scale=1.0) -> torch.Tensor:
    """x-transformers 2.2.8 `apply_rotary_pos_emb`: rotate the first `freqs.shape[-1]` channels of `t`
    (adjacent-pair convention), pass the rest through, fp32 math, cast back."""
    rot_dim, seq_len, orig_dtype = freqs.shape[-1], t.shape[-2], t.dtype
    freqs = freqs[:, -seq_len:, :]
    t, t_unrot = t[..., :rot_dim], t[..., rot_dim:]
    t = t.float()
    t = (t * freqs.cos() * scale) + (_rotate_half(t) * freqs.sin() * scale)
    return torch.cat((t, t_unrot.float()), dim=-1).to(orig_dtype)


def odeint_euler(fn, y0: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """torchdiffeq 0.2.5 fixed-grid Euler on the given grid; returns all len(t) states (model/cfm.py:200)."""
    ys = [y0]
    y = y0
    for t0, t1 in zip(t[:-1], t[1:]):
        dt = t1 - t0
        y = y + dt * fn(t0.to(y.dtype), y)
        ys.append(y)
    return torch.stack(ys)


# ----------------------------------------------------------------------------------------------
# in-tree modules (model/modules.py, model/backbones/dit.py)
# ----------------------------------------------------------------------------------------------


def _lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def sinus_position_embedding(t: torch.Tensor, dim: int = 256, scale: float = 1000.0) -> torch.Tensor:
    """model/modules.py:149-161: [sin | cos], denominator half_dim-1, x1000."""
    half = dim // 2
    emb = math.log(10000) / (half - 1)
    emb = torch.exp(torch.arange(half).float() * -emb)
    emb = scale * t.unsqueeze(1) * emb.unsqueeze(0)
    return torch.cat((emb.sin(), emb.cos()), dim=-1)


def timestep_embedding(sd, t: torch.Tensor, freq_dim: int = 256) -> torch.Tensor:
    """model/modules.py:648-658. t: [b] -> [b, dim]."""
    h = sinus_position_embedding(t, freq_dim).to(t.dtype)
    h = _lin(sd, "transformer.time_embed.time_mlp.0", h)
    h = F.silu(h)
    return _lin(sd, "transformer.time_embed.time_mlp.2", h)


def precompute_freqs_cis(dim: int, end: int, theta: float = 10000.0) -> torch.Tensor:
    """model/modules.py:196-207: [cos | sin] absolute position table [end, dim]."""
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
    t = torch.arange(end)
    freqs = torch.outer(t, freqs).float()
    return torch.cat([torch.cos(freqs), torch.sin(freqs)], dim=-1)


def grn(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor) -> torch.Tensor:
    """model/modules.py:231-234: L2 norm over the SEQUENCE axis (dim=1)."""
    gx = torch.norm(x, p=2, dim=1, keepdim=True)
    nx = gx / (gx.mean(dim=-1, keepdim=True) + 1e-6)
    return gamma * (x * nx) + beta + x


def convnext_v2_block(sd, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """model/modules.py:259-269 (erf-GELU)."""
    res = x
    dim = x.shape[-1]
    h = F.conv1d(x.transpose(1, 2), sd[prefix + "dwconv.weight"], sd[prefix + "dwconv.bias"], padding=3, groups=dim)
    h = h.transpose(1, 2)
    h = F.layer_norm(h, (dim,), sd[prefix + "norm.weight"], sd[prefix + "norm.bias"], eps=1e-6)
    h = _lin(sd, prefix + "pwconv1", h)
    h = F.gelu(h)
    h = grn(h, sd[prefix + "grn.gamma"], sd[prefix + "grn.beta"])
    h = _lin(sd, prefix + "pwconv2", h)
    return res + h


def text_embedding(sd, cfg, text: torch.Tensor, seq_len: int, drop_text: bool) -> torch.Tensor:
    """model/backbones/dit.py:47-69. text int64 [b, nt] (pad -1) -> [b, seq_len, text_dim]."""
    text = text + 1
    text = text[:, :seq_len]
    text = F.pad(text, (0, seq_len - text.shape[1]), value=0)
    if drop_text:
        text = torch.zeros_like(text)
    h = F.embedding(text, sd["transformer.text_embed.text_embed.weight"])
    if cfg.conv_layers > 0:
        table = precompute_freqs_cis(cfg.text_dim, cfg.max_pos)
        pos = torch.arange(seq_len).clamp(max=cfg.max_pos - 1)  # modules.py:210-219 with start=0, scale=1
        h = h + table[pos][None]
        for i in range(cfg.conv_layers):
            h = convnext_v2_block(sd, f"transformer.text_embed.text_blocks.{i}.", h)
    return h


def conv_position_embedding(sd, cfg, x: torch.Tensor) -> torch.Tensor:
    """model/modules.py:178-190 with mask=None (dit.py:86 passes none)."""
    p = "transformer.input_embed.conv_pos_embed.conv1d."
    pad = cfg.conv_pos_kernel // 2
    h = x.permute(0, 2, 1)
    h = F.mish(F.conv1d(h, sd[p + "0.weight"], sd[p + "0.bias"], padding=pad, groups=cfg.conv_pos_groups))
    h = F.mish(F.conv1d(h, sd[p + "2.weight"], sd[p + "2.bias"], padding=pad, groups=cfg.conv_pos_groups))
    return h.permute(0, 2, 1)


def input_embedding(sd, cfg, x, cond, text_embed, drop_audio_cond: bool) -> torch.Tensor:
    """model/backbones/dit.py:81-87."""
    if drop_audio_cond:
        cond = torch.zeros_like(cond)
    h = _lin(sd, "transformer.input_embed.proj", torch.cat((x, cond, text_embed), dim=-1))
    return conv_position_embedding(sd, cfg, h) + h


def attention(sd, cfg, prefix: str, x: torch.Tensor, rope_freqs: torch.Tensor) -> torch.Tensor:
    """model/modules.py:399-449 with mask=None: RoPE on the un-split q,k => only head 0 rotates."""
    b = x.shape[0]
    q = _lin(sd, prefix + "to_q", x)
    k = _lin(sd, prefix + "to_k", x)
    v = _lin(sd, prefix + "to_v", x)
    q = apply_rotary_pos_emb(q, rope_freqs, 1.0)
    k = apply_rotary_pos_emb(k, rope_freqs, 1.0)
    H = cfg.heads
    hd = q.shape[-1] // H
    q = q.view(b, -1, H, hd).transpose(1, 2)
    k = k.view(b, -1, H, hd).transpose(1, 2)
    v = v.view(b, -1, H, hd).transpose(1, 2)
    o = F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False)
    o = o.transpose(1, 2).reshape(b, -1, H * hd).to(q.dtype)
    return _lin(sd, prefix + "to_out.0", o)


def dit_block(sd, cfg, l: int, x: torch.Tensor, t: torch.Tensor, rope_freqs: torch.Tensor) -> torch.Tensor:
    """model/modules.py:558-572 + AdaLayerNormZero :285-290 (chunk order shift,scale,gate x2) + tanh-GELU FFN :556."""
    p = f"transformer.transformer_blocks.{l}."
    D = x.shape[-1]
    emb = _lin(sd, p + "attn_norm.linear", F.silu(t))
    shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = torch.chunk(emb, 6, dim=1)
    norm = F.layer_norm(x, (D,), eps=1e-6) * (1 + scale_msa[:, None]) + shift_msa[:, None]
    a = attention(sd, cfg, p + "attn.", norm, rope_freqs)
    x = x + gate_msa.unsqueeze(1) * a
    norm = F.layer_norm(x, (D,), eps=1e-6) * (1 + scale_mlp[:, None]) + shift_mlp[:, None]
    f = _lin(sd, p + "ff.ff.0.0", norm)
    f = F.gelu(f, approximate="tanh")
    f = _lin(sd, p + "ff.ff.2", f)
    return x + gate_mlp.unsqueeze(1) * f


def dit_forward(sd, cfg, x, cond, text, time, drop_audio_cond: bool, drop_text: bool) -> torch.Tensor:
    """model/backbones/dit.py:130-163 (mask=None, no long skip). x, cond [b,n,mel]; text int64 [b,nt]; time 0-d or [b]."""
    b, n = x.shape[0], x.shape[1]
    if time.ndim == 0:
        time = time.repeat(b)
    t = timestep_embedding(sd, time, cfg.freq_embed_dim)
    te = text_embedding(sd, cfg, text, n, drop_text)
    h = input_embedding(sd, cfg, x, cond, te, drop_audio_cond)
    rope = rotary_freqs(n, cfg.dim_head)
    for l in range(cfg.depth):
        h = dit_block(sd, cfg, l, h, t, rope)
    D = h.shape[-1]
    emb = _lin(sd, "transformer.norm_out.linear", F.silu(t))
    scale, shift = torch.chunk(emb, 2, dim=1)  # modules.py:308 — (scale, shift) order, unlike the blocks
    h = F.layer_norm(h, (D,), eps=1e-6) * (1 + scale)[:, None, :] + shift[:, None, :]
    return _lin(sd, "transformer.proj_out", h)


# ----------------------------------------------------------------------------------------------
# sampler (model/cfm.py:81-210), batch-1
# ----------------------------------------------------------------------------------------------


def sway_time_grid(steps: int, sway_sampling_coef, dtype=torch.float32) -> torch.Tensor:
    """model/cfm.py:196-198."""
    t = torch.linspace(0, 1, steps + 1, dtype=dtype)
    if sway_sampling_coef is not None:
        t = t + sway_sampling_coef * (torch.cos(torch.pi / 2 * t) - 1 + t)
    return t


def list_str_to_idx(text, vocab_char_map, padding_value=-1) -> torch.Tensor:
    """model/utils.py:88-95."""
    from torch.nn.utils.rnn import pad_sequence
    tensors = [torch.tensor([vocab_char_map.get(c, 0) for c in t], dtype=torch.long) for t in text]
    return pad_sequence(tensors, padding_value=padding_value, batch_first=True)


def cfm_sample(sd, cfg, cond_mel: torch.Tensor, text_ids: torch.Tensor, duration: int, *, y0: torch.Tensor | None = None,
               steps=32, cfg_strength=2.0, sway_sampling_coef=-1.0, seed=None, max_duration=4096, lens=None,
               return_trajectory=False):
    """model/cfm.py:100-210 for ONE utterance.  cond_mel [1,F,mel] fp32; text_ids int64 [1,nt] (pad -1).
    `y0` [n,mel] replaces the `torch.randn` draw at cfm.py:181-186 when given (same CPU noise for every
    implementation); otherwise the draw is reproduced exactly (optional manual_seed, then randn)."""
    assert cond_mel.shape[0] == 1 and text_ids.shape[0] == 1
    cond = cond_mel.float()
    cond_seq_len = cond.shape[1]
    lens_t = torch.full((1,), cond_seq_len, dtype=torch.long) if lens is None else torch.as_tensor(lens).view(1).long()
    text_lens = (text_ids != -1).sum(dim=-1)
    lens_t = torch.maximum(text_lens, lens_t)
    cond_mask = (torch.arange(int(lens_t.amax()))[None, :] < lens_t[:, None])
    dur = torch.full((1,), int(duration), dtype=torch.long)
    dur = torch.maximum(lens_t + 1, dur).clamp(max=max_duration)
    n = int(dur.amax())
    cond = F.pad(cond, (0, 0, 0, n - cond_seq_len), value=0.0)
    cond_mask = F.pad(cond_mask, (0, n - cond_mask.shape[-1]), value=False).unsqueeze(-1)
    step_cond = torch.where(cond_mask, cond, torch.zeros_like(cond))

    def fn(t, x):
        pred = dit_forward(sd, cfg, x, step_cond, text_ids, t, False, False)
        if cfg_strength < 1e-5:
            return pred
        null_pred = dit_forward(sd, cfg, x, step_cond, text_ids, t, True, True)
        return pred + (pred - null_pred) * cfg_strength

    if y0 is None:
        if seed is not None:
            torch.manual_seed(seed)
        y0 = torch.randn(n, cfg.mel_dim, dtype=step_cond.dtype)
    y0 = y0[None, :n]
    t = sway_time_grid(steps, sway_sampling_coef, step_cond.dtype)
    traj = odeint_euler(fn, y0, t)
    out = torch.where(cond_mask, cond, traj[-1])
    return (out, traj) if return_trajectory else out


# ----------------------------------------------------------------------------------------------
# prompt mel (model/modules.py:75-101, torchaudio MelSpectrogram restated)
# ----------------------------------------------------------------------------------------------


def _hz_to_mel_htk(f):
    return 2595.0 * math.log10(1.0 + f / 700.0)


def mel_filterbank_htk(n_freqs=513, n_mels=100, sample_rate=24000, f_min=0.0, f_max=None) -> torch.Tensor:
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale='htk') -> [n_freqs, n_mels]."""
    f_max = f_max if f_max is not None else sample_rate / 2
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_pts = torch.linspace(_hz_to_mel_htk(f_min), _hz_to_mel_htk(f_max), n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up))


def mel_spectrogram(wave: torch.Tensor, n_fft=1024, hop=256, n_mels=100, sample_rate=24000) -> torch.Tensor:
    """model/modules.py:75-101: wave [b, nw] -> log-mel [b, n_mels, 1 + nw//hop]."""
    window = torch.hann_window(n_fft)
    spec = torch.stft(wave, n_fft, hop_length=hop, win_length=n_fft, window=window, center=True, pad_mode="reflect",
                      normalized=False, onesided=True, return_complex=True).abs()  # power=1
    fb = mel_filterbank_htk(n_fft // 2 + 1, n_mels, sample_rate)
    mel = torch.matmul(spec.transpose(-1, -2), fb).transpose(-1, -2)
    return mel.clamp(min=1e-5).log()


# ----------------------------------------------------------------------------------------------
# Vocos (third-party, restated)
# ----------------------------------------------------------------------------------------------


def vocos_backbone(vsd, vcfg, mel: torch.Tensor) -> torch.Tensor:
    """mel [b, n_mels, T] -> features [b, T, dim]."""
    C = vcfg.dim
    x = F.conv1d(mel, vsd["backbone.embed.weight"], vsd["backbone.embed.bias"], padding=3)
    x = F.layer_norm(x.transpose(1, 2), (C,), vsd["backbone.norm.weight"], vsd["backbone.norm.bias"], eps=1e-6)
    x = x.transpose(1, 2)
    for i in range(vcfg.num_layers):
        p = f"backbone.convnext.{i}."
        res = x
        h = F.conv1d(x, vsd[p + "dwconv.weight"], vsd[p + "dwconv.bias"], padding=3, groups=C)
        h = h.transpose(1, 2)
        h = F.layer_norm(h, (C,), vsd[p + "norm.weight"], vsd[p + "norm.bias"], eps=1e-6)
        h = F.linear(h, vsd[p + "pwconv1.weight"], vsd[p + "pwconv1.bias"])
        h = F.gelu(h)
        h = F.linear(h, vsd[p + "pwconv2.weight"], vsd[p + "pwconv2.bias"])
        h = vsd[p + "gamma"] * h
        x = res + h.transpose(1, 2)
    return F.layer_norm(x.transpose(1, 2), (C,), vsd["backbone.final_layer_norm.weight"],
                        vsd["backbone.final_layer_norm.bias"], eps=1e-6)


def vocos_istft_head(vsd, vcfg, feats: torch.Tensor) -> torch.Tensor:
    """ISTFTHead(padding='center'): Linear -> (mag, phase) -> exp/clip(1e2) -> cos/sin -> torch.istft(center=True)."""
    x = F.linear(feats, vsd["head.out.weight"], vsd["head.out.bias"]).transpose(1, 2)
    mag, p = x.chunk(2, dim=1)
    mag = torch.exp(mag)
    mag = torch.clip(mag, max=1e2)
    S = mag * (torch.cos(p) + 1j * torch.sin(p))
    return torch.istft(S, vcfg.n_fft, vcfg.hop, vcfg.n_fft, vsd["head.istft.window"], center=True)


def vocos_decode(vsd, vcfg, mel: torch.Tensor) -> torch.Tensor:
    """`Vocos.decode` (call site infer/utils_infer.py:472): mel [b, n_mels, T] -> wav [b, hop*(T-1)]."""
    return vocos_istft_head(vsd, vcfg, vocos_backbone(vsd, vcfg, mel.float()))


# ----------------------------------------------------------------------------------------------
# driver (infer/utils_infer.py:406-524), tensor part only
# ----------------------------------------------------------------------------------------------


def rms_normalise(audio: torch.Tensor, target_rms=0.1):
    """infer/utils_infer.py:427-429."""
    rms = torch.sqrt(torch.mean(torch.square(audio)))
    if rms < target_rms:
        audio = audio * target_rms / rms
    return audio, rms


def estimate_duration(ref_audio_len: int, ref_text: str, gen_text: str, speed=1.0) -> int:
    """infer/utils_infer.py:446-453 (byte-length ratio rule)."""
    return ref_audio_len + int(ref_audio_len / len(ref_text.encode("utf-8")) * len(gen_text.encode("utf-8")) / speed)


def infer_one(sd, cfg, vsd, vcfg, audio: torch.Tensor, text_ids: torch.Tensor, duration: int, *, y0=None,
              steps=32, cfg_strength=2.0, sway_sampling_coef=-1.0, target_rms=0.1, hop=256):
    """One chunk of infer/utils_infer.py:441-482: audio fp32 [1, nw] @24 kHz -> (wave [S], mel [mel, F_gen])."""
    audio, rms = rms_normalise(audio, target_rms)
    ref_audio_len = audio.shape[-1] // hop
    cond = mel_spectrogram(audio).permute(0, 2, 1)
    out = cfm_sample(sd, cfg, cond, text_ids, duration, y0=y0, steps=steps, cfg_strength=cfg_strength,
                     sway_sampling_coef=sway_sampling_coef)
    gen = out.float()[:, ref_audio_len:, :].permute(0, 2, 1)
    wave = vocos_decode(vsd, vcfg, gen)
    if rms < target_rms:
        wave = wave * rms / target_rms
    return wave.squeeze(0), gen[0]


def cross_fade(waves, cross_fade_duration=0.15, sample_rate=24000):
    """infer/utils_infer.py:485-519 (numpy)."""
    import numpy as np
    if cross_fade_duration <= 0:
        return np.concatenate(waves)
    final = waves[0]
    for nxt in waves[1:]:
        n = min(int(cross_fade_duration * sample_rate), len(final), len(nxt))
        if n <= 0:
            final = np.concatenate([final, nxt])
            continue
        mixed = final[-n:] * np.linspace(1, 0, n) + nxt[:n] * np.linspace(0, 1, n)
        final = np.concatenate([final[:-n], mixed, nxt[n:]])
    return final
