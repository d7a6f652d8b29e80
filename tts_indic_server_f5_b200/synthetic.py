"""Deterministic synthetic inputs: prompt audio, Indic text, initial noise, workload configs.

Benchmarks, tests and the golden-vector generator all draw their inputs from here so that the CUDA engine,
the oracle and the reference see byte-identical tensors.  Everything is generated on the CPU with explicit
`torch.Generator`s (device RNGs differ between CPU and CUDA, so noise is never drawn on the device).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch

from . import text as T

SAMPLE_RATE = 24000
HOP = 256
N_MELS = 100


def prompt_audio(seconds: float = 5.0, seed: int = 0, rms: float = 0.1) -> torch.Tensor:
    """Seeded speech-like prompt: harmonic stack with slow pitch/amplitude drift + a little noise. [1, nw] fp32."""
    g = torch.Generator("cpu").manual_seed(10_000 + seed)
    nw = int(round(seconds * SAMPLE_RATE))
    t = torch.arange(nw, dtype=torch.float64) / SAMPLE_RATE
    f0 = 140.0 + 40.0 * torch.sin(2 * math.pi * 0.7 * t + float(torch.rand(1, generator=g)) * 6.28)
    phase = 2 * math.pi * torch.cumsum(f0, 0) / SAMPLE_RATE
    x = torch.zeros(nw, dtype=torch.float64)
    for h in range(1, 25):
        a = float(torch.rand(1, generator=g)) / h
        x += a * torch.sin(h * phase + float(torch.rand(1, generator=g)) * 6.28)
    env = 0.6 + 0.4 * torch.sin(2 * math.pi * 3.1 * t) ** 2
    x = x * env + 0.02 * torch.randn(nw, generator=g, dtype=torch.float64)
    x = x * (rms / x.pow(2).mean().sqrt())
    return x.float()[None]


def initial_noise(n_frames: int, index: int = 0, seed: int = 1234) -> torch.Tensor:
    """`y0_i = randn(dur_i, 100)` from a CPU generator seeded `seed + index` (replaces cfm.py:181-186)."""
    g = torch.Generator("cpu").manual_seed(seed + index)
    return torch.randn(n_frames, N_MELS, generator=g, dtype=torch.float32)


def reference_noise(specs) -> list[torch.Tensor]:
    """The injected noise of every parity comparison (golden vectors, oracle runs, smoke): one CPU draw per utterance,
    seed 1234 + noise_index.  The product's own default is a fresh device draw per request (api.fresh_noise_seed)."""
    return [initial_noise(4096, s.noise_index) for s in specs]


def forward_inputs(n: int, vocab_size: int, prompt_frames: int = 469, nt: int = 300):
    """Seeded inputs of one full-size DiT forward pair (tests/golden/full_fwd.npz stores only the reference's OUTPUTS):
    x [1, n, 100] noise-like state, cond [1, n, 100] log-mel-like prompt (zero past `prompt_frames`), text ids [1, nt]."""
    g = torch.Generator("cpu").manual_seed(1000 + n)
    x = torch.randn(1, n, N_MELS, generator=g)
    cond = torch.randn(1, n, N_MELS, generator=g) * 1.5 - 2.0
    cond[:, prompt_frames:] = 0
    text = torch.randint(0, vocab_size, (1, nt), generator=g)
    return x, cond, text


@dataclass
class UtteranceSpec:
    """One synthesis request at boundary #2 level: prompt wave + texts + (optional) explicit duration."""
    audio: torch.Tensor            # [1, nw] fp32 @ 24 kHz
    ref_text: str
    gen_text: str
    duration: int | None = None    # total frames (prompt + generated); None => byte-ratio rule
    noise_index: int = 0
    meta: dict = field(default_factory=dict)


def _texts_for(ref_cp: int, gen_frames: int, ref_len: int, seed: int, script: str):
    ref_text = T.finish_ref_text(T.synthetic_indic_text(ref_cp, seed, script))
    # size gen_text so that the byte-ratio duration rule lands near gen_frames (3 bytes / code point)
    ref_bytes = len(ref_text.encode("utf-8"))
    gen_cp = max(4, int(round(gen_frames * ref_bytes / ref_len / 3)))
    gen_text = T.synthetic_indic_text(gen_cp, seed + 7919, script)
    return ref_text, gen_text


def workload(name: str, seed: int = 0) -> list[UtteranceSpec]:
    """Named workloads of BASELINE.json `configs` plus small test cases."""
    g = torch.Generator("cpu").manual_seed(777 + seed)
    specs: list[UtteranceSpec] = []

    def add(prompt_s, gen_frames, i, script="kannada", ref_cp=150):
        audio = prompt_audio(prompt_s, seed=i % 4)
        ref_len = audio.shape[-1] // HOP
        rt, gt = _texts_for(ref_cp, gen_frames, ref_len, 31 * i + seed, script)
        specs.append(UtteranceSpec(audio, rt, gt, duration=ref_len + gen_frames, noise_index=i,
                                   meta=dict(gen_frames=gen_frames, ref_len=ref_len)))

    if name == "c1":            # single Kannada utterance, 5 s prompt, server-example length (N = 468 + 317)
        add(5.0, 317, 0)
    elif name == "c1_8s":       # 8 s variant
        add(5.0, 750, 0)
    elif name == "c2":          # 64 mixed-length utterances, ~8 s each (6-10 s)
        for i in range(64):
            gen = int(torch.randint(560, 941, (1,), generator=g))
            add(5.0, gen, i, script="kannada" if i % 2 == 0 else "hindi")
    elif name == "c3":          # long-form 30 s, batch 16
        for i in range(16):
            add(5.0, 2600, i)
    elif name == "c4":          # 512 utterances of the C2 distribution
        for i in range(512):
            gen = int(torch.randint(560, 941, (1,), generator=g))
            add(5.0, gen, i, script="kannada" if i % 2 == 0 else "hindi")
    elif name == "tiny":        # CPU-speed parity case
        add(0.6, 64, 0, ref_cp=24)
    elif name == "tiny3":       # three ragged utterances (packed var-len path)
        for i, (ps, gen) in enumerate(((0.6, 64), (0.9, 23), (0.45, 150))):
            add(ps, gen, i, script="kannada" if i != 1 else "hindi", ref_cp=20 + 9 * i)
    elif name == "small8":      # 8 short utterances for smoke / quick GPU checks
        for i in range(8):
            gen = int(torch.randint(100, 301, (1,), generator=g))
            add(2.0, gen, i, ref_cp=60)
    else:
        raise ValueError(f"unknown workload {name!r}")
    return specs


def generated_audio_seconds(specs_or_frames) -> float:
    """Σ (F_gen − 1)·256/24000 — Vocos' centred ISTFT yields hop·(T−1) samples."""
    tot = 0.0
    for s in specs_or_frames:
        f = s if isinstance(s, int) else s.meta["gen_frames"]
        tot += (f - 1) * HOP / SAMPLE_RATE
    return tot
