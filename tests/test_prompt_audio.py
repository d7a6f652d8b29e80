"""Prompt conditioning (audio half of `preprocess_ref_audio_text`, utils_infer.py:262-320): the numpy implementation the
product uses against the literal audioop-based restatement of pydub 0.25.1 in oracle/pydub_port.py, on synthetic prompts that
exercise every branch — no clipping, clipping on long silences (1), on short silences (2), the hard 15 s cut (3), edge
trimming, stereo, 48 kHz, 32-bit PCM.  Bit-exact: the output is integer PCM."""
import os
import wave

import numpy as np
import pytest

from oracle import pydub_port as P
from tts_indic_server_f5_b200 import prompt_audio as A


def synth(rate, plan, seed=0, channels=1, noise_db=-70.0):
    """plan: list of (seconds, 'v' | 's'): voiced bursts (harmonics, -12 dBFS) and near-silence (noise at noise_db)."""
    rng = np.random.default_rng(seed)
    parts = []
    for sec, kind in plan:
        n = int(round(sec * rate))
        t = np.arange(n) / rate
        if kind == "v":
            x = sum(np.sin(2 * np.pi * f * t + rng.uniform(0, 6.28)) / (k + 1) for k, f in enumerate((140, 280, 420, 700)))
            x = 0.25 * x / np.abs(x).max() * (0.7 + 0.3 * np.sin(2 * np.pi * 2.3 * t))
        else:
            x = rng.standard_normal(n) * 10 ** (noise_db / 20)
        parts.append(x)
    x = np.concatenate(parts)
    pcm = np.clip(np.round(x * 32767), -32768, 32767).astype("<i2")
    if channels == 2:
        pcm = np.stack([pcm, np.roll(pcm, 7)], axis=1).reshape(-1)
    return pcm


def write_wav(path, pcm, rate, channels=1, width=2):
    with wave.open(str(path), "wb") as w:
        w.setnchannels(channels)
        w.setsampwidth(width)
        w.setframerate(rate)
        w.writeframes(pcm.tobytes())


CASES = {
    "short_edges": (24000, 1, [(0.31, "s"), (4.2, "v"), (0.27, "s")]),
    "long_silences": (24000, 1, [(0.2, "s"), (4.0, "v"), (1.4, "s"), (5.0, "v"), (1.3, "s"), (6.5, "v"), (1.2, "s"), (4.0, "v")]),
    "short_silences": (24000, 1, [(3.0, "v"), (0.25, "s"), (4.0, "v"), (0.3, "s"), (5.0, "v"), (0.2, "s"), (6.0, "v"), (0.2, "s"), (3.0, "v")]),
    "no_silence": (16000, 1, [(19.3, "v")]),
    "stereo_48k": (48000, 2, [(0.4, "s"), (3.0, "v"), (1.1, "s"), (2.0, "v"), (0.6, "s")]),
    "odd_rate": (22050, 1, [(0.05, "s"), (2.7183, "v"), (0.033, "s")]),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_clip_reference_matches_pydub_port(tmp_path, name):
    rate, ch, plan = CASES[name]
    path = tmp_path / f"{name}.wav"
    write_wav(path, synth(rate, plan, seed=len(name), channels=ch), rate, ch)
    msgs_a, msgs_p = [], []
    got = A.clip_reference(A.PcmSegment.from_wav(str(path)), True, msgs_a.append)
    ref = P.clip_reference(P.Seg.from_wav(str(path)), True, msgs_p.append)
    assert (got.frame_rate, got.channels, got.sample_width) == (ref.frame_rate, ref.channels, ref.sample_width)
    assert got.data.astype("<i2").tobytes() == ref._data, (name, got.frames, len(ref._data) // ref.frame_width)
    assert msgs_a == msgs_p
    if name == "long_silences":
        assert msgs_a == ["Audio is over 15s, clipping short. (1)"] and len(got) <= 15100
    if name == "short_silences":
        assert msgs_a[-1].endswith("(2)")
    if name == "no_silence":
        assert msgs_a[-1].endswith("(3)") and len(got) == 15050
    # without clipping only the edges change
    got2 = A.clip_reference(A.PcmSegment.from_wav(str(path)), False, print)
    ref2 = P.clip_reference(P.Seg.from_wav(str(path)), False, print)
    assert got2.data.astype("<i2").tobytes() == ref2._data


def test_silence_primitives_match(tmp_path):
    rate = 24000
    path = tmp_path / "p.wav"
    write_wav(path, synth(rate, [(0.5, "s"), (1.0, "v"), (1.234, "s"), (0.8, "v"), (0.111, "s"), (0.7, "v"), (1.0, "s")], 3), rate)
    a, p = A.PcmSegment.from_wav(str(path)), P.Seg.from_wav(str(path))
    assert len(a) == len(p) and a.rms == p.rms and a.dBFS == p.dBFS and a.duration_seconds == p.duration_seconds
    for kw in (dict(min_silence_len=1000, silence_thresh=-50, seek_step=10), dict(min_silence_len=100, silence_thresh=-40, seek_step=10),
               dict(min_silence_len=333, silence_thresh=-45, seek_step=7), dict(min_silence_len=100, silence_thresh=-16, seek_step=1)):
        assert A.detect_silence(a, **kw) == P.detect_silence(p, **kw), kw
        assert A.detect_nonsilent(a, **kw) == P.detect_nonsilent(p, **kw), kw
        sa, sp = A.split_on_silence(a, keep_silence=1000, **kw), P.split_on_silence(p, keep_silence=1000, **kw)
        assert [s.data.astype("<i2").tobytes() for s in sa] == [s._data for s in sp]
    assert A.detect_leading_silence(a, -42) == P.detect_leading_silence(p, -42) == 500
    for lo, hi in ((0, 10), (495, 505), (5339, None), (-20, None), (None, 17)):
        assert a.slice_ms(lo, hi).data.astype("<i2").tobytes() == p[lo:hi]._data


def test_silent_tail_frame_count_is_ratecv():
    for rate in (11025, 16000, 22050, 24000, 44100, 48000):
        ours = A.PcmSegment.silent(50, 1, 2, rate).frames
        theirs = (P.Seg(b"", 2, rate, 1) + P.Seg.silent(duration=50))
        assert ours == len(theirs._data) // 2, rate


def test_32bit_and_24bit_prompts(tmp_path):
    rate = 24000
    pcm16 = synth(rate, [(0.2, "s"), (1.5, "v"), (0.2, "s")], 5)
    path = tmp_path / "p32.wav"
    write_wav(path, (pcm16.astype(np.int64) << 16).astype("<i4"), rate, 1, 4)
    got = A.clip_reference(A.PcmSegment.from_wav(str(path)), True, print)
    ref = P.clip_reference(P.Seg.from_wav(str(path)), True, print)
    assert got.sample_width == ref.sample_width == 4 and got.data.astype("<i4").tobytes() == ref._data


def test_preprocess_ref_audio_text_writes_conditioned_wav(tmp_path):
    from tts_indic_server_f5_b200 import api
    rate = 24000
    src = tmp_path / "prompt.wav"
    write_wav(src, synth(rate, [(0.4, "s"), (3.0, "v"), (0.5, "s")], 9), rate)
    out_path, text = api.preprocess_ref_audio_text(str(src), "ನಮಸ್ಕಾರ", show_info=lambda *_: None)
    assert text == "ನಮಸ್ಕಾರ. " and out_path != str(src) and os.path.exists(out_path)
    ref = P.clip_reference(P.Seg.from_wav(str(src)), True, lambda *_: None)
    with wave.open(out_path, "rb") as w:
        assert (w.getframerate(), w.getnchannels(), w.getsampwidth()) == (rate, 1, 2)
        assert w.readframes(w.getnframes()) == ref._data
    os.unlink(out_path)
    with pytest.raises(NotImplementedError):
        api.preprocess_ref_audio_text(str(src), "   ")


def test_remove_silence_for_generated_wav(tmp_path):
    """utils_infer.py:530-539 (`remove_sil`): in-place rewrite equals the pydub port on a wave with two long pauses."""
    import shutil
    rate = 24000
    src = tmp_path / "gen.wav"
    write_wav(src, synth(rate, [(0.3, "s"), (2.0, "v"), (1.8, "s"), (1.5, "v"), (2.4, "s"), (1.0, "v"), (0.2, "s")], 12), rate)
    ref = P.remove_silence_for_generated_wav_seg(P.Seg.from_wav(str(src)))
    work = tmp_path / "work.wav"
    shutil.copy(src, work)
    A.remove_silence_for_generated_wav(str(work))
    with wave.open(str(work), "rb") as w:
        got = w.readframes(w.getnframes())
    assert got == ref._data and len(got) < os.path.getsize(src) - 2 * rate * 2       # well over a second of pause removed


def test_clip_reference_fuzz_matches_pydub_port(tmp_path):
    """40 seeded random prompts (2 - 30 s; bursts and silences of random lengths from 20 ms to 7 s, noise floors either side
    of the -50 / -42 dBFS thresholds, mono / stereo, four sample rates): the numpy implementation and the audioop restatement
    of the pydub calls must agree byte for byte and print the same messages, with and without clipping."""
    rng = np.random.default_rng(99)
    for case in range(40):
        rate = int(rng.choice([16000, 22050, 24000, 44100]))
        ch = int(rng.choice([1, 1, 2]))
        plan, total = [], 0.0
        target = float(rng.uniform(2.0, 30.0))
        while total < target:
            sec = float(np.exp(rng.uniform(np.log(0.02), np.log(7.0))))
            plan.append((sec, "v" if (len(plan) % 2 == int(case % 2)) else "s"))
            total += sec
        noise_db = float(rng.choice([-75.0, -55.0, -48.0, -44.0, -40.0]))
        path = tmp_path / f"fuzz{case}.wav"
        write_wav(path, synth(rate, plan, seed=1000 + case, channels=ch, noise_db=noise_db), rate, ch)
        for clip in (True, False):
            ma, mp = [], []
            got = A.clip_reference(A.PcmSegment.from_wav(str(path)), clip, ma.append)
            ref = P.clip_reference(P.Seg.from_wav(str(path)), clip, mp.append)
            assert (got.frame_rate, got.channels, got.sample_width) == (ref.frame_rate, ref.channels, ref.sample_width), case
            assert got.data.astype("<i2").tobytes() == ref._data, (case, clip, rate, ch, noise_db, plan)
            assert ma == mp, (case, clip)
        got = A.remove_silence_segment(A.PcmSegment.from_wav(str(path)))             # utils_infer.py:530-539 on the same file
        ref = P.remove_silence_for_generated_wav_seg(P.Seg.from_wav(str(path)))
        assert (got.frame_rate, got.channels, got.sample_width) == (ref.frame_rate, ref.channels, ref.sample_width), case
        assert got.data.astype("<i2").tobytes() == ref._data, (case, "remove_sil")


def test_the_real_preprocess_ref_audio_text_on_the_port_primitives(tmp_path, monkeypatch):
    """The REAL `preprocess_ref_audio_text` (utils_infer.py:282-351) executed in place, with `pydub.AudioSegment` / `pydub.silence`
    bound to the audioop restatement of those primitives (oracle/pydub_port.py): its control flow — the two clipping passes, the
    15 s cut, edge trim, 50 ms tail, WAV export, the '. ' rule on the reference text — then runs as written by the reference's
    authors, and both restatements of it (`P.clip_reference`, the product's `api.preprocess_ref_audio_text`) must produce its
    file, byte for byte, and its text, on every prompt of CASES plus ten random ones."""
    from oracle import ref_shims as R
    if not R.reference_available():
        pytest.skip("reference tree only exists in the build container")
    import types
    from tts_indic_server_f5_b200 import api
    ui = R.load_reference().utils_infer

    def export(self, path, format="wav"):                       # pydub's export(format="wav") is the wave module on the raw data
        assert format == "wav"
        with wave.open(path, "wb") as w:
            w.setnchannels(self.channels)
            w.setsampwidth(self.sample_width)
            w.setframerate(self.frame_rate)
            w.writeframesraw(self._data)

    monkeypatch.setattr(P.Seg, "from_file", classmethod(lambda cls, path: cls.from_wav(path)), raising=False)
    monkeypatch.setattr(P.Seg, "export", export, raising=False)
    monkeypatch.setattr(ui, "AudioSegment", P.Seg)
    monkeypatch.setattr(ui, "silence", types.SimpleNamespace(split_on_silence=P.split_on_silence,
                                                            detect_leading_silence=P.detect_leading_silence))
    rng = np.random.default_rng(5)
    cases = dict(CASES)
    for k in range(10):
        plan, total, target = [], 0.0, float(rng.uniform(2.0, 28.0))
        while total < target:
            sec = float(np.exp(rng.uniform(np.log(0.03), np.log(6.0))))
            plan.append((sec, "v" if len(plan) % 2 == k % 2 else "s"))
            total += sec
        cases[f"rand{k}"] = (int(rng.choice([16000, 24000, 44100])), int(rng.choice([1, 2])), plan)
    texts = ["ನಮಸ್ಕಾರ", "ನಮಸ್ಕಾರ.", "ನಮಸ್ಕಾರ. ", "नमस्ते。", "abc?"]
    for i, (name, (rate, ch, plan)) in enumerate(sorted(cases.items())):
        src = tmp_path / f"{name}.wav"
        write_wav(src, synth(rate, plan, seed=200 + i, channels=ch), rate, ch)
        ref_text = texts[i % len(texts)]
        msgs_r, msgs_a = [], []
        ref_path, ref_out_text = ui.preprocess_ref_audio_text(str(src), ref_text, show_info=msgs_r.append)
        got_path, got_text = api.preprocess_ref_audio_text(str(src), ref_text, show_info=msgs_a.append)
        port = P.clip_reference(P.Seg.from_wav(str(src)), True, lambda *_: None)
        with wave.open(ref_path, "rb") as w:
            want = (w.getframerate(), w.getnchannels(), w.getsampwidth(), w.readframes(w.getnframes()))
        with wave.open(got_path, "rb") as w:
            got = (w.getframerate(), w.getnchannels(), w.getsampwidth(), w.readframes(w.getnframes()))
        os.unlink(ref_path)
        os.unlink(got_path)
        assert got == want, name
        assert (port.frame_rate, port.channels, port.sample_width, port._data) == want, name
        assert got_text == ref_out_text and msgs_a == msgs_r, name
        # the real `remove_silence_for_generated_wav` (utils_infer.py:529-538) and the product's, in place on two copies
        import shutil
        ca, cb = tmp_path / f"{name}_a.wav", tmp_path / f"{name}_b.wav"
        shutil.copy(src, ca)
        shutil.copy(src, cb)
        ui.remove_silence_for_generated_wav(str(ca))
        A.remove_silence_for_generated_wav(str(cb))
        with wave.open(str(ca), "rb") as wa, wave.open(str(cb), "rb") as wb:
            assert wa.getparams()[:3] == wb.getparams()[:3] and wa.readframes(wa.getnframes()) == wb.readframes(wb.getnframes()), name
