"""ORACLE — test infrastructure only (never imported by the product).

Literal restatement of the pydub 0.25.1 calls behind the reference's prompt conditioning
(f5_tts/infer/utils_infer.py:262-320; pydub pinned at pyproject.toml:38, source NOT in the reference tree — parity with
pydub itself is therefore unpinned, this file follows its published code): `AudioSegment` over raw little-endian PCM bytes
with CPython's own `audioop` (the module pydub calls) for rms / ratecv / mul, `pydub.silence.detect_silence`,
`detect_nonsilent`, `split_on_silence`, `detect_leading_silence`, and the reference's `remove_silence_edges` /
clipping passes on top.  Slow on purpose: one slice object per probe, exactly like pydub.
"""
from __future__ import annotations

import itertools
import math
import warnings
import wave

with warnings.catch_warnings():
    warnings.simplefilter("ignore", DeprecationWarning)
    import audioop  # CPython <= 3.12 (pydub 0.25.1 imports the same module)


def db_to_float(db):
    return 10 ** (db / 20)


def ratio_to_db(ratio):
    if ratio == 0:
        return -float("inf")
    return 20 * math.log(ratio, 10)


class Seg:
    """pydub.AudioSegment subset: raw bytes + sample_width + frame_rate + channels."""

    def __init__(self, data: bytes, sample_width: int, frame_rate: int, channels: int):
        self._data, self.sample_width, self.frame_rate, self.channels = data, sample_width, frame_rate, channels
        self.frame_width = sample_width * channels

    @classmethod
    def from_wav(cls, path):
        with wave.open(path, "rb") as w:
            return cls(w.readframes(w.getnframes()), w.getsampwidth(), w.getframerate(), w.getnchannels())

    @classmethod
    def silent(cls, duration=1000, frame_rate=11025):
        frames = int(frame_rate * (duration / 1000.0))
        return cls(b"\0\0" * frames, 2, frame_rate, 1)

    def _spawn(self, data):
        return Seg(data, self.sample_width, self.frame_rate, self.channels)

    def frame_count(self, ms=None):
        if ms is not None:
            return ms * (self.frame_rate / 1000.0)
        return float(len(self._data) // self.frame_width)

    def __len__(self):
        return round(1000 * (self.frame_count() / self.frame_rate))

    @property
    def duration_seconds(self):
        return self.frame_rate and self.frame_count() / self.frame_rate or 0.0

    @property
    def rms(self):
        return audioop.rms(self._data, self.sample_width)

    @property
    def max_possible_amplitude(self):
        return (2 ** (self.sample_width * 8)) / 2

    @property
    def dBFS(self):
        rms = self.rms
        if not rms:
            return -float("inf")
        return ratio_to_db(self.rms / self.max_possible_amplitude)

    def _parse_position(self, val):
        if val < 0:
            val = len(self) - abs(val)
        return int(self.frame_count(ms=val))

    def __getitem__(self, millisecond):
        if isinstance(millisecond, slice):
            start = millisecond.start if millisecond.start is not None else 0
            end = millisecond.stop if millisecond.stop is not None else len(self)
            start, end = min(start, len(self)), min(end, len(self))
        else:
            start, end = millisecond, millisecond + 1
        start = self._parse_position(start) * self.frame_width
        end = self._parse_position(end) * self.frame_width
        data = self._data[start:end]
        expected_length = end - start
        missing_frames = (expected_length - len(data)) // self.frame_width
        if missing_frames:
            if missing_frames > self.frame_count(ms=2):
                raise ValueError("TooManyMissingFrames")
            silence = audioop.mul(data[: self.frame_width], self.sample_width, 0)
            data += silence * missing_frames
        return self._spawn(data)

    def set_frame_rate(self, frame_rate):
        if frame_rate == self.frame_rate:
            return self
        converted = audioop.ratecv(self._data, self.sample_width, self.channels, self.frame_rate, frame_rate, None)[0] if self._data else self._data
        return Seg(converted, self.sample_width, frame_rate, self.channels)

    def set_channels(self, channels):
        if channels == self.channels:
            return self
        if channels == 2 and self.channels == 1:
            return Seg(audioop.tostereo(self._data, self.sample_width, 1, 1), self.sample_width, self.frame_rate, 2)
        raise NotImplementedError

    def set_sample_width(self, sample_width):
        if sample_width == self.sample_width:
            return self
        return Seg(audioop.lin2lin(self._data, self.sample_width, sample_width), sample_width, self.frame_rate, self.channels)

    def __add__(self, other):
        channels, rate, width = max(self.channels, other.channels), max(self.frame_rate, other.frame_rate), max(self.sample_width, other.sample_width)
        a = self.set_channels(channels).set_frame_rate(rate).set_sample_width(width)
        b = other.set_channels(channels).set_frame_rate(rate).set_sample_width(width)
        return Seg(a._data + b._data, width, rate, channels)


def detect_silence(audio_segment, min_silence_len=1000, silence_thresh=-16, seek_step=1):
    seg_len = len(audio_segment)
    if seg_len < min_silence_len:
        return []
    silence_thresh = db_to_float(silence_thresh) * audio_segment.max_possible_amplitude
    silence_starts = []
    last_slice_start = seg_len - min_silence_len
    slice_starts = range(0, last_slice_start + 1, seek_step)
    if last_slice_start % seek_step:
        slice_starts = itertools.chain(slice_starts, [last_slice_start])
    for i in slice_starts:
        if audio_segment[i:i + min_silence_len].rms <= silence_thresh:
            silence_starts.append(i)
    if not silence_starts:
        return []
    silent_ranges = []
    prev_i = silence_starts.pop(0)
    current_range_start = prev_i
    for silence_start_i in silence_starts:
        continuous = silence_start_i == prev_i + seek_step
        silence_has_gap = silence_start_i > (prev_i + min_silence_len)
        if not continuous and silence_has_gap:
            silent_ranges.append([current_range_start, prev_i + min_silence_len])
            current_range_start = silence_start_i
        prev_i = silence_start_i
    silent_ranges.append([current_range_start, prev_i + min_silence_len])
    return silent_ranges


def detect_nonsilent(audio_segment, min_silence_len=1000, silence_thresh=-16, seek_step=1):
    silent_ranges = detect_silence(audio_segment, min_silence_len, silence_thresh, seek_step)
    len_seg = len(audio_segment)
    if not silent_ranges:
        return [[0, len_seg]]
    if silent_ranges[0][0] == 0 and silent_ranges[0][1] == len_seg:
        return []
    prev_end_i = 0
    nonsilent_ranges = []
    for start_i, end_i in silent_ranges:
        nonsilent_ranges.append([prev_end_i, start_i])
        prev_end_i = end_i
    if end_i != len_seg:
        nonsilent_ranges.append([prev_end_i, len_seg])
    if nonsilent_ranges[0] == [0, 0]:
        nonsilent_ranges.pop(0)
    return nonsilent_ranges


def split_on_silence(audio_segment, min_silence_len=1000, silence_thresh=-16, keep_silence=100, seek_step=1):
    if isinstance(keep_silence, bool):
        keep_silence = len(audio_segment) if keep_silence else 0
    output_ranges = [[start - keep_silence, end + keep_silence]
                     for (start, end) in detect_nonsilent(audio_segment, min_silence_len, silence_thresh, seek_step)]
    for range_i, range_ii in zip(output_ranges, output_ranges[1:]):
        last_end, next_start = range_i[1], range_ii[0]
        if next_start < last_end:
            range_i[1] = (last_end + next_start) // 2
            range_ii[0] = range_i[1]
    return [audio_segment[max(start, 0):min(end, len(audio_segment))] for start, end in output_ranges]


def detect_leading_silence(sound, silence_threshold=-50.0, chunk_size=10):
    trim_ms = 0
    while sound[trim_ms:trim_ms + chunk_size].dBFS < silence_threshold and trim_ms < len(sound):
        trim_ms += chunk_size
    return min(trim_ms, len(sound))


def remove_silence_edges(audio, silence_threshold=-42):
    """utils_infer.py:262-276."""
    audio = audio[detect_leading_silence(audio, silence_threshold=silence_threshold):]
    non_silent_end_duration = audio.duration_seconds
    for i in reversed(range(len(audio))):
        if audio[i].dBFS > silence_threshold:
            break
        non_silent_end_duration -= 0.001
    return audio[: int(non_silent_end_duration * 1000)]


def clip_reference(aseg, clip_short=True, show_info=print):
    """utils_infer.py:288-318."""
    if clip_short:
        non_silent_segs = split_on_silence(aseg, min_silence_len=1000, silence_thresh=-50, keep_silence=1000, seek_step=10)
        non_silent_wave = Seg.silent(duration=0)
        for non_silent_seg in non_silent_segs:
            if len(non_silent_wave) > 6000 and len(non_silent_wave + non_silent_seg) > 15000:
                show_info("Audio is over 15s, clipping short. (1)")
                break
            non_silent_wave += non_silent_seg
        if len(non_silent_wave) > 15000:
            non_silent_segs = split_on_silence(aseg, min_silence_len=100, silence_thresh=-40, keep_silence=1000, seek_step=10)
            non_silent_wave = Seg.silent(duration=0)
            for non_silent_seg in non_silent_segs:
                if len(non_silent_wave) > 6000 and len(non_silent_wave + non_silent_seg) > 15000:
                    show_info("Audio is over 15s, clipping short. (2)")
                    break
                non_silent_wave += non_silent_seg
        aseg = non_silent_wave
        if len(aseg) > 15000:
            aseg = aseg[:15000]
            show_info("Audio is over 15s, clipping short. (3)")
    return remove_silence_edges(aseg) + Seg.silent(duration=50)


def remove_silence_for_generated_wav_seg(aseg):
    """utils_infer.py:530-539 on a segment."""
    non_silent_segs = split_on_silence(aseg, min_silence_len=1000, silence_thresh=-50, keep_silence=500, seek_step=10)
    non_silent_wave = Seg.silent(duration=0)
    for non_silent_seg in non_silent_segs:
        non_silent_wave += non_silent_seg
    return non_silent_wave
