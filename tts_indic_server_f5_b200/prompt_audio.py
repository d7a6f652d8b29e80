"""Reference-prompt conditioning, audio half of `preprocess_ref_audio_text` (reference
f5_tts/infer/utils_infer.py:262-320): clip the prompt to <= 15 s on silences, trim silent edges, append 50 ms of silence,
re-export as WAV.  The reference does this with pydub 0.25.1 (`AudioSegment`, `silence.split_on_silence`,
`silence.detect_leading_silence`; pyproject.toml:38 — third party, source not in the reference tree); this module restates
exactly the pydub semantics those calls rely on, on numpy integer PCM, without pydub / ffmpeg:

  * a segment's length is `round(1000 * frames / rate)` ms; slicing is in ms with `int(ms * rate / 1000.0)` frame positions,
    the end clamped to the length and up to 2 ms of missing frames zero-filled (AudioSegment.__getitem__);
  * `rms` is audioop's truncated integer RMS over all samples of all channels; `dBFS = 20 log10(rms / 2^(bits-1))`;
  * `detect_silence` slides a `min_silence_len` window in `seek_step` steps (plus the last possible start), marks windows with
    `rms <= 10^(thresh/20) * 2^(bits-1)`, and merges them into ranges with pydub's continuity rule;
  * `split_on_silence` widens the non-silent ranges by `keep_silence` and splits overlaps at the midpoint;
  * `a + b` brings both sides to max(channels), max(rate), max(width) first, so `silent(50)` (551 frames at pydub's default
    11 025 Hz) becomes `floor(out' (551 - 1) / in') + 1` frames at the prompt's rate (audioop.ratecv's output count).

The sliding-window RMS is evaluated from one prefix sum of squares (exact integers), so a 15 s prompt costs a few
milliseconds instead of pydub's ~1500 slice objects per pass.  `oracle/pydub_port.py` holds the literal (slow, audioop-based)
restatement the tests compare this against.
"""
from __future__ import annotations

import math
import os
import tempfile
import wave
from dataclasses import dataclass

import numpy as np

SILENT_DEFAULT_RATE = 11025          # AudioSegment.silent(frame_rate=11025)


@dataclass
class PcmSegment:
    """Integer PCM: `data` int64 [frames, channels], values in the range of `sample_width` bytes."""
    data: np.ndarray
    sample_width: int
    frame_rate: int

    # ---------------------------------------------------------------- construction / export
    @staticmethod
    def from_wav(path: str) -> "PcmSegment":
        with wave.open(path, "rb") as w:
            sw, ch, rate, n = w.getsampwidth(), w.getnchannels(), w.getframerate(), w.getnframes()
            raw = w.readframes(n)
        if sw == 2:
            a = np.frombuffer(raw, dtype="<i2").astype(np.int64)
        elif sw == 4:
            a = np.frombuffer(raw, dtype="<i4").astype(np.int64)
        elif sw == 3:                                   # pydub widens 24-bit PCM to 32-bit on load
            b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int64)
            a = ((b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)) << 8)
            a = np.where(a >= 1 << 31, a - (1 << 32), a)
            sw = 4
        else:
            raise NotImplementedError(f"{sw * 8}-bit WAV prompts are not supported (16/24/32-bit PCM only)")
        if rate < SILENT_DEFAULT_RATE:
            raise NotImplementedError("prompts below 11 025 Hz would be resampled by pydub's `+`; not supported")
        return PcmSegment(a.reshape(-1, ch), sw, rate)

    @staticmethod
    def from_file(path: str) -> "PcmSegment":
        try:
            return PcmSegment.from_wav(path)
        except (wave.Error, EOFError) as e:
            raise NotImplementedError(f"only PCM WAV prompts are decoded without ffmpeg ({e})") from e

    def export_wav(self, path: str) -> None:
        dt = {2: "<i2", 4: "<i4"}[self.sample_width]
        with wave.open(path, "wb") as w:
            w.setnchannels(self.channels)
            w.setsampwidth(self.sample_width)
            w.setframerate(self.frame_rate)
            w.writeframes(self.data.astype(dt).tobytes())

    def spawn(self, data: np.ndarray) -> "PcmSegment":
        return PcmSegment(data, self.sample_width, self.frame_rate)

    @staticmethod
    def silent(duration_ms: int, channels: int, sample_width: int, frame_rate: int) -> "PcmSegment":
        """`AudioSegment.silent(duration)` as it arrives on the right-hand side of `+` with a segment of this format."""
        n = int(SILENT_DEFAULT_RATE * (duration_ms / 1000.0))
        if frame_rate != SILENT_DEFAULT_RATE and n > 0:
            g = math.gcd(SILENT_DEFAULT_RATE, frame_rate)
            n = (frame_rate // g) * (n - 1) // (SILENT_DEFAULT_RATE // g) + 1      # audioop.ratecv output count
        return PcmSegment(np.zeros((n, channels), dtype=np.int64), sample_width, frame_rate)

    # ---------------------------------------------------------------- pydub's view of a segment
    @property
    def channels(self) -> int:
        return self.data.shape[1]

    @property
    def frames(self) -> int:
        return self.data.shape[0]

    @property
    def max_possible_amplitude(self) -> float:
        return (1 << (8 * self.sample_width)) / 2

    @property
    def duration_seconds(self) -> float:
        return self.frames / self.frame_rate if self.frame_rate else 0.0

    def __len__(self) -> int:
        return round(1000 * (self.frames / self.frame_rate))

    def pos(self, ms) -> int:
        return int(ms * (self.frame_rate / 1000.0))

    def slice_ms(self, start, end) -> "PcmSegment":
        n = len(self)
        start = 0 if start is None else start
        end = n if end is None else end
        if start < 0:
            start = n - abs(start)
        if end < 0:
            end = n - abs(end)
        a, b = self.pos(min(start, n)), self.pos(min(end, n))
        d = self.data[a:b]
        missing = max(b - a, 0) - d.shape[0]
        if missing > 0:
            if missing > self.pos(2):
                raise ValueError("too many missing frames while slicing (pydub TooManyMissingFrames)")
            d = np.concatenate([d, np.zeros((missing, self.channels), dtype=np.int64)])
        return self.spawn(d)

    @property
    def rms(self) -> int:
        if self.data.size == 0:
            return 0
        return _isqrt_mean(_sumsq(self.data), self.data.size)

    @property
    def dBFS(self) -> float:
        r = self.rms
        return -float("inf") if r == 0 else 20 * math.log(r / self.max_possible_amplitude, 10)

    def __add__(self, other: "PcmSegment") -> "PcmSegment":
        ch, rate, sw = max(self.channels, other.channels), max(self.frame_rate, other.frame_rate), max(self.sample_width, other.sample_width)
        a, b = self._as(ch, rate, sw), other._as(ch, rate, sw)
        return PcmSegment(np.concatenate([a.data, b.data]), sw, rate)

    def _as(self, ch: int, rate: int, sw: int) -> "PcmSegment":
        if (self.channels, self.frame_rate, self.sample_width) == (ch, rate, sw):
            return self
        if self.frames == 0 or not self.data.any():      # empty or all-zero (the `silent()` operands): only the frame count converts
            n = self.frames
            if rate != self.frame_rate and n > 0:
                g = math.gcd(self.frame_rate, rate)
                n = (rate // g) * (n - 1) // (self.frame_rate // g) + 1
            return PcmSegment(np.zeros((n, ch), dtype=np.int64), sw, rate)
        raise NotImplementedError("concatenating prompts of different formats needs pydub's resampler; not on the served path")


def _sumsq(a: np.ndarray):
    if np.abs(a).max(initial=0) < (1 << 15) + 1:
        return int(np.sum(a * a))
    return sum(int(v) * int(v) for v in a.reshape(-1))       # 32-bit samples: exact Python integers


def _isqrt_mean(sum_squares: int, count: int) -> int:
    """audioop.rms: (unsigned int) sqrt(sum_squares / (double) count) — a double division, a double sqrt, truncation."""
    return int(math.sqrt(sum_squares / float(count)))


def db_to_float(db: float) -> float:
    return 10 ** (db / 20)


# --------------------------------------------------------------------------------------------- pydub.silence
def _window_rms(seg: PcmSegment, starts: list[int], win_ms: int) -> list[int]:
    """RMS of seg[s : s + win_ms] for every s in starts, from one prefix sum of squares (exact integers for 16-bit PCM)."""
    n_ms = len(seg)
    per_frame = seg.data * seg.data if seg.sample_width <= 2 else None
    if per_frame is None:
        return [seg.slice_ms(s, s + win_ms).rms for s in starts]
    csum = np.concatenate([[0], np.cumsum(per_frame.sum(axis=1))])
    out = []
    for s in starts:
        a, b = seg.pos(min(s, n_ms)), seg.pos(min(s + win_ms, n_ms))
        frames = max(b - a, 0)
        if frames == 0:
            out.append(0)
            continue
        ss = int(csum[min(b, seg.frames)] - csum[min(a, seg.frames)])   # frames past the data are the zero fill of __getitem__
        out.append(_isqrt_mean(ss, frames * seg.channels))
    return out


def detect_silence(seg: PcmSegment, min_silence_len=1000, silence_thresh=-16, seek_step=1) -> list[list[int]]:
    seg_len = len(seg)
    if seg_len < min_silence_len:
        return []
    thresh = db_to_float(silence_thresh) * seg.max_possible_amplitude
    last_start = seg_len - min_silence_len
    starts = list(range(0, last_start + 1, seek_step))
    if last_start % seek_step:
        starts.append(last_start)
    rms = _window_rms(seg, starts, min_silence_len)
    silence_starts = [s for s, r in zip(starts, rms) if r <= thresh]
    if not silence_starts:
        return []
    ranges = []
    prev_i = silence_starts.pop(0)
    cur = prev_i
    for s in silence_starts:
        continuous = s == prev_i + seek_step
        has_gap = s > prev_i + min_silence_len
        if not continuous and has_gap:
            ranges.append([cur, prev_i + min_silence_len])
            cur = s
        prev_i = s
    ranges.append([cur, prev_i + min_silence_len])
    return ranges


def detect_nonsilent(seg: PcmSegment, min_silence_len=1000, silence_thresh=-16, seek_step=1) -> list[list[int]]:
    silent = detect_silence(seg, min_silence_len, silence_thresh, seek_step)
    n = len(seg)
    if not silent:
        return [[0, n]]
    if silent[0][0] == 0 and silent[0][1] == n:
        return []
    prev_end, out = 0, []
    for s, e in silent:
        out.append([prev_end, s])
        prev_end = e
    if e != n:
        out.append([prev_end, n])
    if out[0] == [0, 0]:
        out.pop(0)
    return out


def split_on_silence(seg: PcmSegment, min_silence_len=1000, silence_thresh=-16, keep_silence=100, seek_step=1) -> list[PcmSegment]:
    if isinstance(keep_silence, bool):
        keep_silence = len(seg) if keep_silence else 0
    ranges = [[s - keep_silence, e + keep_silence] for s, e in detect_nonsilent(seg, min_silence_len, silence_thresh, seek_step)]
    for r0, r1 in zip(ranges, ranges[1:]):
        if r1[0] < r0[1]:
            r0[1] = (r0[1] + r1[0]) // 2
            r1[0] = r0[1]
    return [seg.slice_ms(max(s, 0), min(e, len(seg))) for s, e in ranges]


def detect_leading_silence(seg: PcmSegment, silence_threshold=-50.0, chunk_size=10) -> int:
    trim = 0
    while seg.slice_ms(trim, trim + chunk_size).dBFS < silence_threshold and trim < len(seg):
        trim += chunk_size
    return min(trim, len(seg))


def remove_silence_edges(seg: PcmSegment, silence_threshold=-42) -> PcmSegment:
    """utils_infer.py:262-276."""
    seg = seg.slice_ms(detect_leading_silence(seg, silence_threshold=silence_threshold), None)
    end = seg.duration_seconds
    for i in reversed(range(len(seg))):                  # `for ms in reversed(audio)`: one-millisecond slices from the end
        if seg.slice_ms(i, i + 1).dBFS > silence_threshold:
            break
        end -= 0.001
    return seg.slice_ms(None, int(end * 1000))


def clip_reference(seg: PcmSegment, clip_short=True, show_info=print) -> PcmSegment:
    """utils_infer.py:288-318: the <= 15 s clipping passes, edge trim and the 50 ms tail."""
    if clip_short:
        def gather(min_silence_len, silence_thresh, label):
            out = PcmSegment.silent(0, 1, 2, 11025)      # AudioSegment.silent(duration=0): pydub's defaults; `+` lifts it to the source format
            for s in split_on_silence(seg, min_silence_len=min_silence_len, silence_thresh=silence_thresh, keep_silence=1000, seek_step=10):
                if len(out) > 6000 and len(out + s) > 15000:
                    show_info(f"Audio is over 15s, clipping short. ({label})")
                    break
                out = out + s
            return out
        wave_ = gather(1000, -50, 1)
        if len(wave_) > 15000:
            wave_ = gather(100, -40, 2)
        seg = wave_
        if len(seg) > 15000:
            seg = seg.slice_ms(None, 15000)
            show_info("Audio is over 15s, clipping short. (3)")
    seg = remove_silence_edges(seg)
    return seg + PcmSegment.silent(50, seg.channels, seg.sample_width, seg.frame_rate)


def preprocess_ref_audio(ref_audio_orig: str, clip_short=True, show_info=print) -> str:
    """-> path of a temporary WAV holding the conditioned prompt (what the reference hands to `infer_process`)."""
    seg = clip_reference(PcmSegment.from_file(os.fspath(ref_audio_orig)), clip_short, show_info)
    with tempfile.NamedTemporaryFile(delete=False, suffix=".wav") as f:
        name = f.name
    seg.export_wav(name)
    return name


def remove_silence_segment(seg: PcmSegment) -> PcmSegment:
    """utils_infer.py:530-539 on a segment: keep the non-silent stretches (>= 1 s below -50 dBFS counts as silence, 500 ms of it
    kept on each side) and join them."""
    out = PcmSegment.silent(0, 1, 2, 11025)              # as above: a prompt with no non-silent stretch stays an empty 11 025 Hz segment
    for s in split_on_silence(seg, min_silence_len=1000, silence_thresh=-50, keep_silence=500, seek_step=10):
        out = out + s
    return out


def remove_silence_for_generated_wav(filename: str) -> None:
    """`remove_silence_for_generated_wav(filename)` (utils_infer.py:530-539; the `remove_sil` option of the CLI / gradio paths, off on
    the served path): rewrites the WAV file in place."""
    remove_silence_segment(PcmSegment.from_file(filename)).export_wav(filename)
