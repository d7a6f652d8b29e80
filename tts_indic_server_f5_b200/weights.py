"""Architecture record + deterministic synthetic weights for the IndicF5 DiT and the Vocos vocoder.

There is no network in the build/bench environment, so throughput and parity are measured on
random-init weights of the named architecture (BASELINE.json).  The state-dict KEY NAMES and SHAPES
are the reference's own (reference `f5_tts/model/backbones/dit.py:93-128`,
`f5_tts/model/modules.py:167-176,241-257,276-283,297-304,317-325,361-373,542-556,648-652`, and
vocos 0.1.0 `VocosBackbone`/`ISTFTHead`), so a real IndicF5 / `charactr/vocos-mel-24khz` checkpoint
loads through the same path (`load_checkpoint` key rules: `f5_tts/infer/utils_infer.py:175-218`).

Every tensor is drawn from its own `torch.Generator` seeded by (seed, crc32(key)), so values do not
depend on module construction order and are identical in the oracle, the reference (loaded through
`load_state_dict`) and the CUDA engine.
"""
from __future__ import annotations

import math
import zlib
from dataclasses import dataclass

import torch


@dataclass(frozen=True)
class DiTConfig:
    """`model_cfg` of the reference (`f5_tts/infer/infer_cli.py:136`) + fixed constants."""
    dim: int = 1024
    depth: int = 22
    heads: int = 16
    dim_head: int = 64
    ff_mult: int = 2
    text_dim: int = 512
    conv_layers: int = 4
    conv_mult: int = 2          # dit.py:33
    mel_dim: int = 100          # utils_infer.py:41
    vocab_size: int = 263       # text_num_embeds; embedding has vocab_size+1 rows (dit.py:35)
    conv_pos_kernel: int = 31   # modules.py:168
    conv_pos_groups: int = 16
    freq_embed_dim: int = 256   # modules.py:649
    max_pos: int = 4096         # dit.py:39

    @property
    def ff_inner(self) -> int:
        return int(self.dim * self.ff_mult)

    @property
    def text_inner(self) -> int:
        return self.text_dim * self.conv_mult


@dataclass(frozen=True)
class VocosConfig:
    """`charactr/vocos-mel-24khz/config.yaml`."""
    n_mels: int = 100
    dim: int = 512
    intermediate_dim: int = 1536
    num_layers: int = 8
    n_fft: int = 1024
    hop: int = 256


INDICF5 = DiTConfig()
VOCOS_24K = VocosConfig()


def tiny_dit_config(vocab_size: int = 263) -> DiTConfig:
    """Small DiT used by CPU-speed parity tests and the committed golden vectors."""
    return DiTConfig(dim=256, depth=3, heads=4, ff_mult=2, text_dim=128, conv_layers=2, vocab_size=vocab_size)


def tiny_vocos_config() -> VocosConfig:
    return VocosConfig(dim=128, intermediate_dim=384, num_layers=2)


def _gen(seed: int, key: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1_000_003 + zlib.crc32(key.encode())) % (2**63 - 1))
    return g


def _uniform(seed, key, shape, bound):
    return (torch.rand(shape, generator=_gen(seed, key), dtype=torch.float32) * 2 - 1) * bound


def _normal(seed, key, shape, std, mean=0.0):
    return torch.randn(shape, generator=_gen(seed, key), dtype=torch.float32) * std + mean


def _linear(sd, seed, name, out_f, in_f, fan_in=None, w_shape=None):
    fan_in = fan_in or in_f
    b = 1.0 / math.sqrt(fan_in)
    sd[name + ".weight"] = _uniform(seed, name + ".weight", w_shape or (out_f, in_f), b)
    sd[name + ".bias"] = _uniform(seed, name + ".bias", (out_f,), b)


def _layernorm(sd, seed, name, dim):
    sd[name + ".weight"] = _normal(seed, name + ".weight", (dim,), 0.1, 1.0)
    sd[name + ".bias"] = _normal(seed, name + ".bias", (dim,), 0.1)


def make_dit_state_dict(cfg: DiTConfig = INDICF5, seed: int = 0) -> dict[str, torch.Tensor]:
    """CFM state dict (keys prefixed `transformer.` exactly like the reference's `CFM.state_dict()`)."""
    sd: dict[str, torch.Tensor] = {}
    p = "transformer."
    D, TD = cfg.dim, cfg.text_dim
    _linear(sd, seed, p + "time_embed.time_mlp.0", D, cfg.freq_embed_dim)
    _linear(sd, seed, p + "time_embed.time_mlp.2", D, D)
    sd[p + "text_embed.text_embed.weight"] = _normal(seed, p + "text_embed.text_embed.weight", (cfg.vocab_size + 1, TD), 1.0)
    for i in range(cfg.conv_layers):
        b = f"{p}text_embed.text_blocks.{i}."
        _linear(sd, seed, b + "dwconv", TD, 1, fan_in=7, w_shape=(TD, 1, 7))
        _layernorm(sd, seed, b + "norm", TD)
        _linear(sd, seed, b + "pwconv1", cfg.text_inner, TD)
        sd[b + "grn.gamma"] = _normal(seed, b + "grn.gamma", (1, 1, cfg.text_inner), 0.5)
        sd[b + "grn.beta"] = _normal(seed, b + "grn.beta", (1, 1, cfg.text_inner), 0.5)
        _linear(sd, seed, b + "pwconv2", TD, cfg.text_inner)
    _linear(sd, seed, p + "input_embed.proj", D, 2 * cfg.mel_dim + TD)
    cpg = D // cfg.conv_pos_groups
    for j in (0, 2):
        _linear(sd, seed, f"{p}input_embed.conv_pos_embed.conv1d.{j}", D, cpg,
                fan_in=cpg * cfg.conv_pos_kernel, w_shape=(D, cpg, cfg.conv_pos_kernel))
    for l in range(cfg.depth):
        b = f"{p}transformer_blocks.{l}."
        _linear(sd, seed, b + "attn_norm.linear", 6 * D, D)
        for n in ("to_q", "to_k", "to_v"):
            _linear(sd, seed, b + "attn." + n, D, D)
        _linear(sd, seed, b + "attn.to_out.0", D, D)
        _linear(sd, seed, b + "ff.ff.0.0", cfg.ff_inner, D)
        _linear(sd, seed, b + "ff.ff.2", D, cfg.ff_inner)
    _linear(sd, seed, p + "norm_out.linear", 2 * D, D)
    _linear(sd, seed, p + "proj_out", cfg.mel_dim, D)
    return sd


def make_vocos_state_dict(cfg: VocosConfig = VOCOS_24K, seed: int = 0) -> dict[str, torch.Tensor]:
    sd: dict[str, torch.Tensor] = {}
    C, I = cfg.dim, cfg.intermediate_dim
    _linear(sd, seed, "backbone.embed", C, cfg.n_mels, fan_in=cfg.n_mels * 7, w_shape=(C, cfg.n_mels, 7))
    _layernorm(sd, seed, "backbone.norm", C)
    for i in range(cfg.num_layers):
        b = f"backbone.convnext.{i}."
        _linear(sd, seed, b + "dwconv", C, 1, fan_in=7, w_shape=(C, 1, 7))
        _layernorm(sd, seed, b + "norm", C)
        _linear(sd, seed, b + "pwconv1", I, C)
        _linear(sd, seed, b + "pwconv2", C, I)
        sd[b + "gamma"] = _normal(seed, b + "gamma", (C,), 0.02, 1.0 / cfg.num_layers)
    _layernorm(sd, seed, "backbone.final_layer_norm", C)
    _linear(sd, seed, "head.out", cfg.n_fft + 2, C)
    # keep log-magnitudes modest so exp() stays far from the 1e2 clip most of the time, but not always
    sd["head.out.weight"][: cfg.n_fft // 2 + 1] *= 0.5
    sd["head.istft.window"] = torch.hann_window(cfg.n_fft)
    return sd


def strip_checkpoint(ckpt: dict, use_ema: bool = True) -> dict[str, torch.Tensor]:
    """Key rules of the reference loader (`f5_tts/infer/utils_infer.py:195-213`): pick the EMA dict,
    strip the `ema_model.` prefix, drop `initted`/`step` and the two legacy mel buffers."""
    if use_ema and "ema_model_state_dict" in ckpt:
        sd = {k.replace("ema_model.", ""): v for k, v in ckpt["ema_model_state_dict"].items()
              if k not in ("initted", "step")}
    elif "model_state_dict" in ckpt:
        sd = dict(ckpt["model_state_dict"])
    else:
        sd = {k.replace("ema_model.", ""): v for k, v in ckpt.items() if k not in ("initted", "step")}
    for key in ("mel_spec.mel_stft.mel_scale.fb", "mel_spec.mel_stft.spectrogram.window"):
        sd.pop(key, None)
    return sd


def infer_dit_config(sd: dict[str, torch.Tensor]) -> DiTConfig:
    """Recover the architecture from a state dict's shapes (real-weights path)."""
    p = "transformer."
    D = sd[p + "proj_out.weight"].shape[1]
    depth = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith(p + "transformer_blocks."))
    TD = sd[p + "text_embed.text_embed.weight"].shape[1]
    conv_layers = 1 + max((int(k.split(".")[3]) for k in sd if k.startswith(p + "text_embed.text_blocks.")), default=-1)
    ff_inner = sd[p + "transformer_blocks.0.ff.ff.0.0.weight"].shape[0]
    return DiTConfig(dim=D, depth=depth, heads=D // 64, ff_mult=ff_inner // D, text_dim=TD,
                     conv_layers=conv_layers, mel_dim=sd[p + "proj_out.weight"].shape[0],
                     vocab_size=sd[p + "text_embed.text_embed.weight"].shape[0] - 1)
