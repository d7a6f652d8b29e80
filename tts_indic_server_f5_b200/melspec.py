"""Prompt log-mel front-end (`MelSpec` / `get_vocos_mel_spectrogram`, reference f5_tts/model/modules.py:75-143).

Runs once per request on the prompt only (SURVEY.md §8f "next" row 2): STFT via torch.stft (cuFFT plumbing), HTK mel
filterbank restated from torchaudio.functional.melscale_fbanks(norm=None, mel_scale="htk"), log(clamp(1e-5)).
"""
from __future__ import annotations

import math

import torch

_cache: dict = {}


def _fbank(n_freqs: int, n_mels: int, sample_rate: int, device) -> torch.Tensor:
    key = (n_freqs, n_mels, sample_rate, str(device))
    if key not in _cache:
        hz2mel = lambda f: 2595.0 * math.log10(1.0 + f / 700.0)  # noqa: E731
        all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
        m_pts = torch.linspace(hz2mel(0.0), hz2mel(sample_rate / 2), n_mels + 2)
        f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
        f_diff = f_pts[1:] - f_pts[:-1]
        slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
        down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
        up = slopes[:, 2:] / f_diff[1:]
        _cache[key] = (torch.max(torch.zeros(1), torch.min(down, up)).to(device), torch.hann_window(1024).to(device))
    return _cache[key]


def mel_spectrogram(wave: torch.Tensor, n_fft=1024, hop=256, n_mels=100, sample_rate=24000) -> torch.Tensor:
    """wave fp32 [b, nw] -> log-mel [b, n_mels, 1 + nw // hop] on wave's device."""
    fb, window = _fbank(n_fft // 2 + 1, n_mels, sample_rate, wave.device)
    spec = torch.stft(wave, n_fft, hop_length=hop, win_length=n_fft, window=window, center=True, pad_mode="reflect",
                      normalized=False, onesided=True, return_complex=True).abs()
    mel = torch.matmul(spec.transpose(-1, -2), fb).transpose(-1, -2)
    return mel.clamp(min=1e-5).log()
