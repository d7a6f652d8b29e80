"""ORACLE — test infrastructure only.  Import the REAL reference `f5_tts` package in place.

Runs only where `/root/reference` exists (the build container); nothing that runs on the GPU box
(`-m gpu` tests, `smoke()`, `bench.py`) may call `load_reference()`.  It is used to
  (a) validate `oracle/f5_oracle.py` against the reference's own modules on identical weights/inputs
      (`tests/test_oracle_golden.py::test_oracle_vs_real_reference_modules`), and
  (b) mint the golden vectors committed under `tests/golden/` (`oracle/make_golden.py`).

The reference imports nine packages that are not installed here.  Six are irrelevant to the arithmetic
(librosa, jieba, pypinyin, matplotlib, pydub, the trainer) and get inert fakes.  Three carry arithmetic
(x_transformers RoPE, torchdiffeq Euler, vocos); their fakes delegate to the restatements in
`oracle/f5_oracle.py` (published-algorithm restatements, "parity unpinned", see that file's header).
Recipe and gotchas: SURVEY.md Appendix D.
"""
from __future__ import annotations

import importlib
import importlib.machinery
import os
import sys
import types

import torch
from torch import nn

from . import f5_oracle as O

REFERENCE_ROOT = os.environ.get("F5_REFERENCE_ROOT", "/root/reference")
_SERVER = os.path.join(REFERENCE_ROOT, "src", "server")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(_SERVER, "f5_tts", "model"))


def _fake(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _RotaryEmbedding(nn.Module):
    """x-transformers 2.2.8 RotaryEmbedding surface used at dit.py:117,149."""

    def __init__(self, dim, **_):
        super().__init__()
        self.dim = dim
        self.register_buffer("inv_freq", 1.0 / (10000.0 ** (torch.arange(0, dim, 2).float() / dim)))

    def forward_from_seq_len(self, seq_len):
        return O.rotary_freqs(seq_len, self.dim).to(self.inv_freq.device), 1.0


class _RMSNorm(nn.Module):  # UNetT only; never executed on the hot path
    def __init__(self, dim):
        super().__init__()
        self.scale = dim ** 0.5
        self.g = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        return torch.nn.functional.normalize(x, dim=-1) * self.scale * self.g


class _VocosBackbone(nn.Module):
    def __init__(self, vcfg):
        super().__init__()
        C, I = vcfg.dim, vcfg.intermediate_dim
        self.embed = nn.Conv1d(vcfg.n_mels, C, 7, padding=3)
        self.norm = nn.LayerNorm(C, eps=1e-6)
        self.convnext = nn.ModuleList()
        for _ in range(vcfg.num_layers):
            blk = nn.Module()
            blk.dwconv = nn.Conv1d(C, C, 7, padding=3, groups=C)
            blk.norm = nn.LayerNorm(C, eps=1e-6)
            blk.pwconv1 = nn.Linear(C, I)
            blk.pwconv2 = nn.Linear(I, C)
            blk.gamma = nn.Parameter(torch.full((C,), 1.0 / vcfg.num_layers))
            self.convnext.append(blk)
        self.final_layer_norm = nn.LayerNorm(C, eps=1e-6)


class _ISTFT(nn.Module):
    def __init__(self, n_fft):
        super().__init__()
        self.register_buffer("window", torch.hann_window(n_fft))


class _ISTFTHead(nn.Module):
    def __init__(self, vcfg):
        super().__init__()
        self.out = nn.Linear(vcfg.dim, vcfg.n_fft + 2)
        self.istft = _ISTFT(vcfg.n_fft)


class FakeVocos(nn.Module):
    """Stand-in for `vocos.Vocos` (vocos==0.1.0 is not installed): same state-dict keys, `decode()` from the
    oracle's restatement.  There is no independent pin for the vocoder — parity unpinned."""

    def __init__(self, vcfg):
        super().__init__()
        self.vcfg = vcfg
        self.feature_extractor = nn.Module()
        self.backbone = _VocosBackbone(vcfg)
        self.head = _ISTFTHead(vcfg)

    @torch.inference_mode()
    def decode(self, mel):
        return O.vocos_decode({k: v for k, v in self.state_dict().items()}, self.vcfg, mel)


_loaded = None


def load_reference():
    """Returns a namespace with the reference's own `CFM`, `DiT`, `utils_infer`, `model_utils` modules."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    from transformers import pipeline  # noqa: F401  (must precede the fake librosa, Appendix D.1)

    xt = _fake("x_transformers", RMSNorm=_RMSNorm, RotaryEmbedding=_RotaryEmbedding)
    xtx = _fake("x_transformers.x_transformers", RotaryEmbedding=_RotaryEmbedding,
                apply_rotary_pos_emb=O.apply_rotary_pos_emb, RMSNorm=_RMSNorm)
    xt.x_transformers = xtx
    _fake("torchdiffeq", odeint=lambda fn, y0, t, **kw: O.odeint_euler(fn, y0, t))
    lib = _fake("librosa")
    lib.filters = _fake("librosa.filters", mel=lambda **kw: None)
    _fake("jieba", initialize=lambda: None, cut=lambda text: [text])
    _fake("pypinyin", lazy_pinyin=lambda seg, **kw: list(seg), Style=types.SimpleNamespace(TONE3=3))
    voc = _fake("vocos", Vocos=FakeVocos)
    voc.feature_extractors = _fake("vocos.feature_extractors", EncodecFeatures=type("EncodecFeatures", (), {}))
    mpl = _fake("matplotlib", use=lambda *a, **k: None)
    mpl.pylab = _fake("matplotlib.pylab")
    _fake("pydub", AudioSegment=type("AudioSegment", (), {}), silence=types.SimpleNamespace())
    _fake("f5_tts.model.trainer", Trainer=type("Trainer", (), {}))
    if _SERVER not in sys.path:
        sys.path.insert(0, _SERVER)
    model = importlib.import_module("f5_tts.model")
    utils_infer = importlib.import_module("f5_tts.infer.utils_infer")
    model_utils = importlib.import_module("f5_tts.model.utils")
    modules = importlib.import_module("f5_tts.model.modules")
    _loaded = types.SimpleNamespace(CFM=model.CFM, DiT=model.DiT, utils_infer=utils_infer,
                                    model_utils=model_utils, modules=modules)
    return _loaded


def build_reference_cfm(sd: dict, cfg, vocab_char_map=None):
    """Instantiate the reference's own CFM(DiT) with `cfg`, load `sd` (reference key names), eval, CPU fp32."""
    ref = load_reference()
    dit = ref.DiT(dim=cfg.dim, depth=cfg.depth, heads=cfg.heads, dim_head=cfg.dim_head, ff_mult=cfg.ff_mult,
                  mel_dim=cfg.mel_dim, text_num_embeds=cfg.vocab_size, text_dim=cfg.text_dim,
                  conv_layers=cfg.conv_layers)
    cfm = ref.CFM(transformer=dit,
                  mel_spec_kwargs=dict(n_fft=1024, hop_length=256, win_length=1024, n_mel_channels=cfg.mel_dim,
                                       target_sample_rate=24000, mel_spec_type="vocos"),
                  odeint_kwargs=dict(method="euler"), vocab_char_map=vocab_char_map)
    missing, unexpected = cfm.load_state_dict(sd, strict=False)
    missing = [k for k in missing if "inv_freq" not in k]
    assert not missing and not unexpected, (missing, unexpected)
    return cfm.eval()


def build_reference_vocos(vsd: dict, vcfg) -> FakeVocos:
    v = FakeVocos(vcfg)
    v.load_state_dict(vsd)
    return v.eval()
