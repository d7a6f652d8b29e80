"""Prompt log-mel front-end (`MelSpec` / `get_vocos_mel_spectrogram`, reference f5_tts/model/modules.py:75-143).

Runs once per request on the prompt only.  The arithmetic is the `f5_mel_frames` kernel
(`csrc/mel_frontend.cu`: reflect-padded frames, hann window, 1024-point FFT in shared memory, magnitude, HTK mel bands,
log(clamp 1e-5)); this module only builds the constant tables — the HTK filterbank restated from
torchaudio.functional.melscale_fbanks(norm=None, mel_scale="htk") and the periodic hann window — and a small per-voice
cache so that a server that always sends the same prompt (`tts_utils.py:31-36`) computes its mel once.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

from . import ops

N_FFT, HOP = 1024, 256
_tables: dict = {}


def htk_fbank(n_freqs: int, n_mels: int, sample_rate: int) -> torch.Tensor:
    """fp32 [n_freqs, n_mels] triangular HTK filters, f_min 0, f_max sample_rate/2, no area normalisation."""
    hz2mel = lambda f: 2595.0 * math.log10(1.0 + f / 700.0)  # noqa: E731
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_pts = torch.linspace(hz2mel(0.0), hz2mel(sample_rate / 2), n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up)).contiguous()


def band_ranges(fb: torch.Tensor) -> torch.Tensor:
    """int32 [n_mels, 2]: half-open range of frequency bins where each filter is non-zero (empty filters -> [0, 0))."""
    nz = fb > 0
    out = torch.zeros(fb.shape[1], 2, dtype=torch.int32)
    for m in range(fb.shape[1]):
        idx = torch.nonzero(nz[:, m]).flatten()
        if idx.numel():
            out[m, 0], out[m, 1] = int(idx[0]), int(idx[-1]) + 1
    return out


def _get_tables(n_mels: int, sample_rate: int, device):
    key = (n_mels, sample_rate, str(device))
    if key not in _tables:
        fb = htk_fbank(N_FFT // 2 + 1, n_mels, sample_rate)
        _tables[key] = (fb.to(device), band_ranges(fb).to(device), torch.hann_window(N_FFT).to(device))
    return _tables[key]


def mel_rows(waves: list[torch.Tensor], n_mels=100, sample_rate=24000) -> list[torch.Tensor]:
    """Device fp32 1-D prompts (any lengths) -> list of log-mel [frames_i, n_mels] views of ONE output buffer; one launch."""
    dev = waves[0].device
    if dev.type != "cuda":
        raise RuntimeError("the prompt mel front-end is a CUDA kernel (f5_mel_frames); there is no CPU path")
    fb, band, window = _get_tables(n_mels, sample_rate, dev)
    lens = [int(w.numel()) for w in waves]
    if min(lens) <= N_FFT // 2:
        raise ValueError("prompt shorter than n_fft/2 samples: reflect padding is undefined (torch.stft rejects it too)")
    frames = [1 + n // HOP for n in lens]
    flat = torch.cat([w.reshape(-1).float() for w in waves]) if len(waves) > 1 else waves[0].reshape(-1).float().contiguous()
    seg, o, r = [], 0, 0
    for n, f in zip(lens, frames):
        seg.append([o, n, r, f])
        o += n
        r += f
    mel = torch.empty(r, n_mels, device=dev, dtype=torch.float32)
    ops.mel_frames(flat, torch.tensor(seg, dtype=torch.int32).to(dev), max(frames), window, fb, band, mel)
    return [mel[s[2]:s[2] + s[3]] for s in seg]


def mel_spectrogram(wave: torch.Tensor, n_fft=N_FFT, hop=HOP, n_mels=100, sample_rate=24000) -> torch.Tensor:
    """wave fp32 [b, nw] (CUDA) -> log-mel [b, n_mels, 1 + nw // hop] — the reference's call shape (modules.py:104-143)."""
    assert n_fft == N_FFT and hop == HOP, "the kernel is built for n_fft 1024 / hop 256 (vocos-mel-24khz)"
    if wave.ndim == 1:
        wave = wave[None]
    rows = mel_rows([wave[i] for i in range(wave.shape[0])], n_mels, sample_rate)
    return torch.stack(rows).permute(0, 2, 1)


class PromptCache:
    """LRU of prompt log-mels keyed by the prompt's samples (after mono / RMS / resample).  The reference server sends ONE
    fixed voice prompt with every request and recomputes its mel each time (`utils_infer.py:424-433`)."""

    def __init__(self, capacity: int = 64):
        self.capacity, self.hits, self.misses = capacity, 0, 0
        self._d: OrderedDict = OrderedDict()

    @staticmethod
    def key(audio: torch.Tensor) -> tuple:
        """Content key of a prompt: length + CRC-32 of all sample bytes + two 64-bit integer checksums of coprime-strided subsets
        of the sample bit patterns + the first and last 8 samples: ~0.25 ms per 5 s prompt.  Not an adversarial hash (a
        cryptographic one over the 480 KB costs more than the mel kernel it saves); for an accidental collision a different
        waveform of the same length would have to match the CRC and 128 bits of sums at once."""
        import zlib
        a = audio.detach().reshape(-1)
        if a.device.type != "cpu" or a.dtype != torch.float32:
            a = a.float().cpu()
        a = a.contiguous()
        w = a.view(torch.int32)
        return (int(w.numel()), zlib.crc32(a.numpy().data), int(w[::7].sum(dtype=torch.int64)), int(w[3::11].sum(dtype=torch.int64)),
                a[:8].numpy().tobytes(), a[-8:].numpy().tobytes())

    def get(self, k, device):
        v = self._d.get((k, str(device)))
        if v is not None:
            self._d.move_to_end((k, str(device)))
            self.hits += 1
        return v

    def put(self, k, device, mel: torch.Tensor) -> None:
        self.misses += 1
        self._d[(k, str(device))] = mel
        while len(self._d) > self.capacity:
            self._d.popitem(last=False)
