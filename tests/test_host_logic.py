"""CPU: host-side logic of the product (text front-end, packed layout, sharding, weight layouts, C ABI surface)."""
import ctypes
import os
import re
import subprocess
import tempfile

import pytest
import numpy as np
import torch
import torch.nn.functional as F

from tts_indic_server_f5_b200 import text as T
from tts_indic_server_f5_b200 import weights as W
from tts_indic_server_f5_b200.dist import lpt_partition, utterance_cost
from tts_indic_server_f5_b200.layout import GAP, build_layout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tokenizer_one_token_per_codepoint_and_vocab_roundtrip(tmp_path):
    toks = T.synthetic_indic_vocab()
    assert toks[0] == " " and len(toks) == len(set(toks)) == W.INDICF5.vocab_size
    p = tmp_path / "vocab.txt"
    T.write_vocab(str(p), toks)
    vmap, size = T.get_tokenizer(str(p), "custom")
    assert size == len(toks) and vmap[" "] == 0 and vmap["ಕ"] == toks.index("ಕ")
    s = "ಕನ್ನಡ ಪದ, हिन्दी."
    out = T.convert_char_to_pinyin([s])[0]
    assert out == list(s)                                   # virama / matras are separate tokens, spaces preserved
    ids = T.list_str_to_idx([out, out[:3]], vmap)
    assert ids.shape == (2, len(s)) and (ids[1, 3:] == -1).all() and (ids[0] >= 0).all()
    assert T.list_str_to_idx([["一x"]], vmap)[0, 0] == 0   # unknown -> 0
    with pytest.raises(ValueError):
        T.convert_char_to_pinyin(["中文"])
    assert T.convert_char_to_pinyin(["ಕ;ab"])[0] == ["ಕ", ",", " ", "a", "b"]   # ';'->',' and ASCII-run leading space


def test_tokenizer_fast_path_equals_character_scan(monkeypatch):
    """Pure-Indic text takes a fast path (one token per code point); it must equal the full character scan of
    utils.py:140-177 on arbitrary mixtures of Indic code points, punctuation, quotes, digits, decimals and percent signs."""
    import random
    import re
    from tts_indic_server_f5_b200 import text as T
    rng = random.Random(0)
    alphabet = [chr(c) for c in range(0x0C80, 0x0CA0)] + list(" .,;!?:'\"%") + list("ab7Z09") + ["\u201c", "\u2019"]
    strings = ["".join(rng.choice(alphabet) for _ in range(rng.randint(1, 30))) for _ in range(3000)]
    fast = T.convert_char_to_pinyin(strings)
    monkeypatch.setattr(T, "_NEEDS_SCAN", re.compile(""))          # always scan
    assert T.convert_char_to_pinyin(strings) == fast
    assert any(T._ASCII_RUN.search(s) for s in strings)


def test_duration_and_ref_text_rules():
    assert T.finish_ref_text("abc") == "abc. " and T.finish_ref_text("abc.") == "abc. " and T.finish_ref_text("abc. ") == "abc. "
    assert T.estimate_duration(468, "x" * 100, "y" * 50) == 468 + 234
    assert T.estimate_duration(468, "x" * 100, "y" * 50, speed=2.0) == 468 + 117
    assert T.estimate_duration(468, "r", "g", fix_duration=10.0) == int(10.0 * 24000 / 256)
    assert T.chunk_text("a. b. c.", 3) == ["a.", "b.", "c."]
    assert T.chunk_text("", 10) == []


def test_layout_invariants():
    lens = [120, 1, 300, 129]
    L = build_layout(lens)
    R = L.half_rows
    assert R % 128 == 0 and L.row_pos.numel() == 2 * R
    pos = L.row_pos[:R]
    assert torch.equal(L.row_pos[R:], pos)
    for s, n in zip(L.starts, lens):
        assert torch.equal(pos[s:s + n], torch.arange(n, dtype=torch.int32))
        assert (pos[s - GAP:s] == -1).all() and (pos[s + n:s + n + GAP] == -1).all()
    assert int((pos >= 0).sum()) == sum(lens)
    # every real row is covered by exactly one query tile per half; kv ranges are the utterance's own rows
    cover = torch.zeros(2 * R, dtype=torch.int32)
    for q0, kv0, kvl, qv in L.attn_tiles.tolist():
        cover[q0:q0 + qv] += 1
        assert (L.row_pos[kv0:kv0 + kvl] == torch.arange(kvl, dtype=torch.int32)).all()
        assert kv0 <= q0 < kv0 + kvl and 1 <= qv <= 256
    assert torch.equal(cover, (L.row_pos >= 0).to(torch.int32))
    assert L.seg_rows.tolist() == [[h + s, n] for h in (0, R) for s, n in zip(L.starts, lens)]


def test_lpt_partition_balances_and_covers():
    g = torch.Generator().manual_seed(0)
    lens = torch.randint(1029, 1410, (512,), generator=g).tolist()
    for w in (1, 2, 4, 8):
        parts = lpt_partition(lens, w)
        assert sorted(i for p in parts for i in p) == list(range(512))
        loads = [sum(utterance_cost(lens[i]) for i in p) for p in parts]
        assert max(loads) / (sum(loads) / w) < 1.02          # <2 % imbalance at 64 utterances / GPU


def test_scheduler_packing_and_cross_fade():
    """Request scheduler host logic: packs cover every utterance once, respect the row and count
    budgets, are length-bucketed; the cross-fade equals the reference's formula (utils_infer.py:485-519)."""
    from tts_indic_server_f5_b200.layout import build_layout
    from tts_indic_server_f5_b200.scheduler import cross_fade, pack_rows, plan_packs
    g = torch.Generator().manual_seed(3)
    lengths = [int(x) for x in torch.randint(300, 4097, (200,), generator=g)]
    packs = plan_packs(lengths, max_rows=65536, max_utts=24)
    assert sorted(i for p in packs for i in p) == list(range(200))
    for p in packs:
        assert len(p) <= 24
        rows = pack_rows([lengths[i] for i in p])
        assert rows == build_layout([lengths[i] for i in p]).rows          # the planner's row count is the engine's
        assert rows <= 65536 or len(p) == 1
    firsts = [lengths[p[0]] for p in packs]
    assert firsts == sorted(firsts, reverse=True)                          # longest-first buckets
    assert plan_packs([5000], max_rows=1024) == [[0]]                      # an oversize utterance still runs, alone
    assert plan_packs([], 1024) == []
    a, b, c = (np.random.RandomState(k).randn(n).astype(np.float32) for k, n in ((0, 9000), (1, 5000), (2, 100)))
    n = int(0.15 * 24000)
    want = np.concatenate([a[:-n], a[-n:] * np.linspace(1, 0, n) + b[:n] * np.linspace(0, 1, n), b[n:]])
    np.testing.assert_array_equal(cross_fade([a, b]), want)
    got = cross_fade([a, b, c])                                            # third chunk shorter than the fade window
    assert len(got) == len(want) + len(c) - min(n, len(c))
    np.testing.assert_array_equal(cross_fade([a, b], 0.0), np.concatenate([a, b]))


def test_conv_pos_weight_block_diagonal_is_exact():
    from tts_indic_server_f5_b200.engine import _conv_pos_weight
    D, G, K, n = 256, 16, 31, 50
    g = torch.Generator().manual_seed(1)
    w = torch.randn(D, D // G, K, generator=g)
    x = torch.randn(n, D, generator=g)
    ref = F.conv1d(x.t()[None], w, None, padding=K // 2, groups=G)[0].t()
    wt = _conv_pos_weight(w).reshape(K, D, 64)
    xp = F.pad(x, (0, 0, K // 2, K // 2))
    out = torch.zeros(n, D)
    for t in range(K):
        for sg in range(D // 64):
            out[:, sg * 64:(sg + 1) * 64] += xp[t:t + n, sg * 64:(sg + 1) * 64] @ wt[t, sg * 64:(sg + 1) * 64].t()
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-4)


def test_checkpoint_key_rules_and_config_inference():
    cfg = W.tiny_dit_config()
    sd = W.make_dit_state_dict(cfg, seed=2)
    ckpt = {"ema_model_state_dict": {**{"ema_model." + k: v for k, v in sd.items()}, "initted": torch.tensor(1), "step": torch.tensor(5),
                                     "ema_model.mel_spec.mel_stft.mel_scale.fb": torch.zeros(1)}}
    out = W.strip_checkpoint(ckpt)
    assert set(out) == set(sd)
    assert W.infer_dit_config(out) == cfg
    assert W.make_dit_state_dict(cfg, seed=2)["transformer.proj_out.weight"].equal(sd["transformer.proj_out.weight"])
    full = W.make_dit_state_dict(W.INDICF5, seed=0)
    assert full["transformer.transformer_blocks.21.attn_norm.linear.weight"].shape == (6144, 1024)
    assert full["transformer.input_embed.proj.weight"].shape == (1024, 712)
    assert full["transformer.input_embed.conv_pos_embed.conv1d.2.weight"].shape == (1024, 64, 31)


def test_sway_grid_matches_oracle_and_closed_form():
    from oracle import f5_oracle as O
    from tts_indic_server_f5_b200.engine import sway_time_grid
    t = sway_time_grid(32, -1.0)
    assert torch.equal(t, O.sway_time_grid(32, -1.0))
    i = torch.arange(33, dtype=torch.float64)
    torch.testing.assert_close(t.double(), 1 - torch.cos(torch.pi * i / 64), rtol=0, atol=2e-7)
    assert torch.equal(sway_time_grid(8, None), torch.linspace(0, 1, 9))


# ------------------------------------------------------------------------------------------------ C ABI surface
def _declared_symbols():
    h = open(os.path.join(ROOT, "include", "f5_b200.h")).read()
    return sorted(set(re.findall(r"^(?:int|const char\*)\s+(f5_\w+)\s*\(", h, flags=re.M)))


def test_library_exports_every_declared_symbol():
    from tts_indic_server_f5_b200 import _lib
    syms = _declared_symbols()
    assert len(syms) >= 16 and sorted(_lib.EXPORTS) == syms
    for s in syms:
        assert hasattr(_lib.lib, s)
    assert b"sm_100a" in _lib.lib.f5_version()


def test_gemm_args_struct_layout_matches_header():
    from tts_indic_server_f5_b200 import _lib
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "f5_b200.h"\nint main(){printf("%zu %zu %zu %zu", sizeof(f5_gemm_args),' \
          ' offsetof(f5_gemm_args, mode), offsetof(f5_gemm_args, resid), offsetof(f5_gemm_args, num_sms));return 0;}'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        size, o_mode, o_resid, o_sms = map(int, subprocess.check_output([exe]).split())
    G = _lib.GemmArgs
    assert (ctypes.sizeof(G), G.mode.offset, G.resid.offset, G.num_sms.offset) == (size, o_mode, o_resid, o_sms)


def test_launchers_reject_bad_arguments_without_touching_the_gpu():
    from tts_indic_server_f5_b200 import _lib
    a = _lib.GemmArgs()
    assert _lib.lib.f5_gemm_bf16(ctypes.byref(a), None) == -1            # F5_ERR_ARG: null operands
    a.A, a.B, a.M, a.N, a.num_taps, a.kc_per_tap, a.block_n = 1 << 20, 1 << 21, 128, 100, 1, 1, 256
    assert _lib.lib.f5_gemm_bf16(ctypes.byref(a), None) == -1            # N % 8 != 0
    assert _lib.lib.f5_layernorm_mod(None, 0, None, 0, None, 0, 1, 128, None, None, 1.0, 1e-6, 0, None) == -1
    assert _lib.lib.f5_attention_f32(None, 0, 0, 0, 0, 16, None, 0, None, None, 0, 0, 0.125, None) == -1
    a.taps_per_seg, a.num_taps, a.N = 2, 3, 128                            # split-operand mode needs num_taps == 3 * taps_per_seg
    assert _lib.lib.f5_gemm_bf16(ctypes.byref(a), None) == -1
    assert _lib.lib.f5_attention_d64(None, 0, 0, 0, 0, 0, 16, None, 0, None, 0, 0.125, None) == -1
    with pytest.raises(_lib.F5Error):
        _lib.check(-2, "x")


def test_engine_refuses_to_run_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tts_indic_server_f5_b200 import api
    with pytest.raises(RuntimeError):
        api.load_model(device="cpu")
    with pytest.raises(RuntimeError):
        api.load_model(device="cuda")


def test_wav_response_matches_a_wav_reader():
    """Response body format of `synthesize_speech` (tts_utils.py:60-65): 24 kHz mono 16-bit PCM WAV, int16 input rescaled
    by 1/32768 first.  Read back with the stdlib `wave` module and, when present, with soundfile (the reference's writer)."""
    import wave
    from tts_indic_server_f5_b200.api import wav_response_bytes
    rng = np.random.default_rng(0)
    x = (0.3 * rng.standard_normal(24000)).astype(np.float32)
    x[:4] = [1.5, -1.5, 1.0, -1.0]                                   # out-of-range samples clip, full scale maps to +-32767
    buf = wav_response_bytes(x)
    with wave.open(buf, "rb") as w:
        assert (w.getframerate(), w.getnchannels(), w.getsampwidth(), w.getnframes()) == (24000, 1, 2, 24000)
        pcm = np.frombuffer(w.readframes(24000), dtype="<i2")
    assert pcm[:4].tolist() == [32767, -32767, 32767, -32767]
    assert np.array_equal(pcm, np.rint(np.clip(x.astype(np.float64), -1, 1) * 32767).astype(np.int16))
    # the int16 return type of the model object goes through the same 1/32768 rescale as the server applies
    xi = (x[4:] * 32768).astype(np.int16)
    with wave.open(wav_response_bytes(xi), "rb") as w:
        pcm_i = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2")
    assert np.abs(pcm_i.astype(np.int32) - xi.astype(np.int32)).max() <= 1
    with pytest.raises(ValueError):
        wav_response_bytes(np.zeros((2, 10), np.float32))
    try:
        import soundfile as sf
    except Exception:
        return
    import io
    ref = io.BytesIO()
    sf.write(ref, x[4:], 24000, format="WAV")
    assert ref.getvalue() == wav_response_bytes(x[4:]).getvalue()


def test_sample_prologue_rules():
    """`CFM.sample` prologue (cfm.py:118-137, :181-186) as host logic: lens = max(text_lens, lens); duration =
    clamp(max(lens + 1, duration), <= max_duration); one noise draw per item re-seeded with the same seed (so items share
    a noise prefix); pad ids (-1) stripped.  Runs without a GPU on a stub of the engine-backed object."""
    from types import SimpleNamespace
    from tts_indic_server_f5_b200.api import CFM
    stub = SimpleNamespace(vocab_char_map={" ": 0, "a": 1, "b": 2}, num_channels=100, _device="cpu")
    cond = torch.randn(3, 40, 100)
    text = torch.tensor([[1, 2, 1, -1, -1, -1], [1] * 6, [2, 2, -1, -1, -1, -1]])
    utts = CFM._prepare(stub, cond, text, torch.tensor([30, 100, 5000]), torch.tensor([40, 3, 40]), 7, 4096, None, None)
    assert [u.n for u in utts] == [41, 100, 4096]                 # lens + 1 floor, as given, clamped to max_duration
    assert [u.cond_len for u in utts] == [40, 6, 40]              # item 1: text_lens (6) > lens (3)
    assert [u.text_ids.tolist() for u in utts] == [[1, 2, 1], [1] * 6, [2, 2]]
    # shared seed => shared noise prefix (torch's CPU normal_ recomputes the last 16 values of a draw, so up to those)
    assert torch.equal(utts[0].y0.flatten()[:-16], utts[1].y0.flatten()[:4100 - 16])
    assert torch.equal(utts[1].y0.flatten()[:-16], utts[2].y0.flatten()[:10000 - 16])
    torch.manual_seed(7)
    assert torch.equal(utts[2].y0, torch.randn(4096, 100))
    # an int duration broadcasts; strings go through the vocabulary (unknown -> 0)
    utts = CFM._prepare(stub, cond[:1], ["abz"], 64, None, None, 4096, None, [torch.ones(80, 100)])
    assert utts[0].n == 64 and utts[0].cond_len == 40 and utts[0].text_ids.tolist() == [1, 2, 0]
    assert utts[0].y0.shape == (64, 100) and bool((utts[0].y0 == 1).all())


# ------------------------------------------------------------------------------------------------ round 2: noise, loader
def test_philox_oracle_matches_random123_known_answers():
    """Pins oracle/philox.py (the checker of the device noise kernel) to the published Random123 philox4x32-10 vectors."""
    from oracle import philox as P
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = P.philox4x32_10(*[[c] for c in ctr], *key)
        assert tuple(int(g[0]) for g in got) == want
    z = P.randn_rows(0xDEADBEEFCAFEF00D, 3000)
    assert z.shape == (3000, 100) and abs(float(z.mean())) < 5e-3 and abs(float(z.std()) - 1.0) < 5e-3
    assert np.array_equal(P.randn_rows(7, 50)[:20], P.randn_rows(7, 20))          # counter-based: a prefix is a prefix
    assert not np.array_equal(P.randn_rows(7, 20), P.randn_rows(8, 20))


def test_noise_seeds_follow_the_global_generator():
    from tts_indic_server_f5_b200 import api
    torch.manual_seed(123)
    a = [api.fresh_noise_seed() for _ in range(3)]
    torch.manual_seed(123)
    assert a == [api.fresh_noise_seed() for _ in range(3)] and len(set(a)) == 3      # repeatable after manual_seed, fresh otherwise
    s = {api.utterance_seed(a[0], i) for i in range(1000)}
    assert len(s) == 1000 and all(0 <= x < 2 ** 64 for x in s)


def test_checkpoint_roundtrip_with_the_reference_modules_real_keys(tmp_path):
    """`utils_infer.py:175-218`: an EMA checkpoint written from the REAL reference module's state dict (shim-loaded CFM) as
    .safetensors and as .pt comes back through `strip_checkpoint` with exactly the keys the reference's own strict
    `load_state_dict` accepts, and `infer_dit_config` recovers the architecture.  Needs /root/reference (build container)."""
    from oracle import ref_shims as R
    if not R.reference_available():
        pytest.skip("reference tree not present (GPU box)")
    from safetensors.torch import load_file, save_file
    cfg = W.tiny_dit_config()
    vocab = {t: i for i, t in enumerate(T.synthetic_indic_vocab())}
    cfm = R.build_reference_cfm(W.make_dit_state_dict(cfg, seed=4), cfg, vocab)
    real = {k: v.clone() for k, v in cfm.state_dict().items()}
    assert any(k.startswith("mel_spec.") for k in real) or all(k.startswith("transformer.") for k in real)
    ema = {"ema_model." + k: v.contiguous() for k, v in real.items()}
    ema_pt = {**ema, "initted": torch.tensor(True), "step": torch.tensor(1200000)}
    save_file({**ema, "initted": torch.tensor([1]), "step": torch.tensor([1200000])}, str(tmp_path / "model.safetensors"))
    torch.save({"ema_model_state_dict": ema_pt}, str(tmp_path / "model.pt"))
    for sd in (W.strip_checkpoint(load_file(str(tmp_path / "model.safetensors"))),
               W.strip_checkpoint(torch.load(str(tmp_path / "model.pt"), map_location="cpu", weights_only=True))):
        # what the reference does with the same file (utils_infer.py:195-213) must accept it strictly
        legacy = ("mel_spec.mel_stft.mel_scale.fb", "mel_spec.mel_stft.spectrogram.window")
        assert set(sd) == {k for k in real if k not in legacy}
        missing, unexpected = cfm.load_state_dict(sd, strict=False)
        assert not unexpected and set(missing) <= set(legacy)
        assert W.infer_dit_config(sd) == cfg
        for k in sd:
            assert torch.equal(sd[k], real[k])
    # non-EMA branch (:214-217)
    torch.save({"model_state_dict": real}, str(tmp_path / "plain.pt"))
    sd = W.strip_checkpoint(torch.load(str(tmp_path / "plain.pt"), map_location="cpu", weights_only=True), use_ema=False)
    assert set(sd) >= {k for k in real if k.startswith("transformer.")}


def test_continuous_scheduler_batches_concurrent_requests():
    """Worker thread + bounded queue: requests submitted from several threads are drained together, every
    future gets ITS request's result, back-pressure raises queue.Full, an engine error reaches every waiter of the batch."""
    import queue
    import threading
    import time
    from tts_indic_server_f5_b200.scheduler import ContinuousScheduler

    class Prep:
        def __init__(self, d):
            self.duration = d

    class FakeSyn:
        def __init__(self):
            self.calls, self.fail = [], False

        def _prep(self, spec, speed, fixd):
            return Prep(40 + len(spec.gen_text))

        def generate(self, specs, nfe, cfg, sway, speed, fixd, return_mel=False):
            time.sleep(0.05)                                     # the "GPU" is busy: later submissions pile up in the queue
            if self.fail:
                raise RuntimeError("CUDA error: unspecified launch failure")
            self.calls.append(len(specs))
            waves = [np.full(256 * 4, float(len(s.gen_text)), dtype=np.float32) for s in specs]
            mels = [np.full((100, 5), float(len(s.gen_text)), dtype=np.float32) for s in specs]
            return waves, mels

    syn = FakeSyn()
    cs = ContinuousScheduler(syn, max_queue=8, max_batch_requests=16, max_wait_ms=30.0, nfe_step=4)
    audio = torch.zeros(1, 24000)
    futs, lock = {}, threading.Lock()

    def client(k):
        f = cs.submit((audio, 24000), "ref. ", "x" * (k + 1), seed=k)
        with lock:
            futs[k] = f

    ths = [threading.Thread(target=client, args=(k,)) for k in range(6)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    for k, f in futs.items():
        wave, sr, mel = f.result(timeout=10)
        assert sr == 24000 and float(wave[0]) == k + 1 and mel.shape == (100, 5)      # each future got its own request
    assert sum(cs.batches) == 6 and max(cs.batches) >= 2                               # concurrent requests shared a batch
    # back-pressure: a full queue refuses instead of growing
    syn2 = FakeSyn()
    cs2 = ContinuousScheduler(syn2, max_queue=1, max_batch_requests=1, max_wait_ms=0.0, nfe_step=4)
    held = [cs2.submit((audio, 24000), "ref. ", "abc")]
    with pytest.raises(queue.Full):
        for _ in range(50):
            held.append(cs2.submit((audio, 24000), "ref. ", "abc", timeout=0.0))
    # an engine failure is delivered to the waiters
    syn.fail = True
    f = cs.submit((audio, 24000), "ref. ", "boom")
    with pytest.raises(RuntimeError, match="launch failure"):
        f.result(timeout=10)
    cs.close()
    cs2.close()
    with pytest.raises(RuntimeError):
        cs.submit((audio, 24000), "ref. ", "late")


def test_out_of_line_wait_fits_the_low_register_warps():
    """`mbar_wait_warp_slow` is the one device function the tcgen05 kernels CALL; its callers are the producer / MMA warps, which
    run on 56 registers (attention; GEMM with eight epilogue warps) after `setmaxnreg.dec`.  ptxas allocates the callee's
    registers per kernel without knowing that budget, so the built library is checked: every register the callee touches must
    exist in the calling warp."""
    import shutil
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    from tts_indic_server_f5_b200 import _lib
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB], capture_output=True, text=True).stdout
    body, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            body[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur:
            body[cur].append((int(m.group(1), 16), m.group(2)))
    checked = 0
    for fn, ins in body.items():
        if "attn_d64_kernel" in fn:
            budget = 72 if "attn_d64" in fn else 56
        elif "gemm_tcgen05_kernel" in fn and fn.split("gemm_tcgen05_kernelI")[1].startswith(("Li256ELi1ELi1ELi8", "Li256ELi2ELi1ELi8", "Li256ELi3ELi1ELi8",
                                                                                             "Li256ELi1ELi2ELi8", "Li256ELi2ELi2ELi8", "Li256ELi3ELi2ELi8")):
            budget = 72 if "attn_d64" in fn else 56
        else:
            continue
        targets = sorted({int(x, 16) for _, i in ins for x in re.findall(r"CALL\.REL\.NOINC (0x[0-9a-f]+)", i)})
        for t in targets:
            callee = []
            for a, i in ins:
                if a < t:
                    continue
                callee.append(i)
                if "RET.REL" in i and not i.strip().startswith("@"):
                    break
            if not any("TRYWAIT" in i for i in callee):
                continue
            top = max(int(r) for i in callee for r in re.findall(r"\bR(\d+)\b", i))
            assert top < budget, f"{fn}: out-of-line wait uses R{top}, the calling warp has {budget} registers"
            checked += 1
    assert checked >= 7


def test_tile_width_rule():
    """ops.pick_block_n: 256 for a full batch always; 128 where 256-wide tiles leave whole waves of the 148 SMs idle (a single
    request: M = 1792 -> 56 tiles for N = 1024, two waves for QKV) — DESIGN.md §7, profiles/r02_launches_c1.csv."""
    from tts_indic_server_f5_b200 import ops
    for M in (158976, 98816, 40000):
        assert [ops.pick_block_n(N, M) for N in (1024, 2048, 3072, 512)] == [256] * 4
    assert ops.pick_block_n(1024, 1792) == 128 and ops.pick_block_n(3072, 1792) == 128 and ops.pick_block_n(2048, 1792) == 256
    assert ops.pick_block_n(1032, 1792) == 128 and ops.pick_block_n(104, 1792) == 64 and ops.pick_block_n(1024) == 256


def test_speech_route_mirrors_the_reference_and_batches_concurrent_requests(tmp_path):
    """`server.create_app` against the reference's route behaviour (routes/speech.py:19-41,
    utils/tts_utils.py:39-65): 503 while the model is not loaded, 400 for empty text / unknown voice / empty reference text,
    audio/wav attachment with a 16-bit 24 kHz PCM body that decodes to the engine's samples, X-Response-Time header; and the part
    the reference lacks: requests that arrive together are ONE engine batch.  CPU: a fake engine behind the real scheduler."""
    import io
    import threading
    import time
    import wave as wavmod
    from fastapi.testclient import TestClient
    from tts_indic_server_f5_b200 import api, server

    class Prep:
        def __init__(self, d):
            self.duration = d

    class FakeSyn:
        device = None

        def _prep(self, spec, speed, fixd):
            return Prep(40 + len(spec.gen_text))

        def generate(self, specs, nfe, cfg, sway, speed, fixd, return_mel=False):
            time.sleep(0.15)
            waves = [np.full(2400, 0.001 * len(s.gen_text), dtype=np.float32) for s in specs]
            return waves, [np.zeros((100, 5), dtype=np.float32) for _ in specs]

    class FakeModel:
        output_int16, ema_model, vocoder = True, object(), object()

        def __init__(self):
            self.prompt_calls = 0

        def _prompt(self, path, ref_text):
            self.prompt_calls += 1
            return (torch.zeros(1, 24000), 24000), ref_text

    class FakeManager:
        repo_id, failed_reason = "ai4bharat/IndicF5", None

        def __init__(self):
            self.model, self.loads = None, 0

        def load(self):
            self.loads += 1
            self.model = FakeModel()

    syn = FakeSyn()
    orig = api._synthesizer_for
    api._synthesizer_for = lambda m, v: syn
    try:
        mgr = FakeManager()
        voices = {"KAN_F (Happy)": server.Voice(str(tmp_path / "kan.wav"), "ref text. ")}
        with pytest.raises(ValueError):
            server.create_app(mgr, voices, default_voice="nobody")
        app = server.create_app(mgr, voices, max_wait_ms=60.0)
        no_lifespan = TestClient(app)                                   # without the lifespan the model is not loaded yet
        assert no_lifespan.post("/v1/audio/speech", json={"text": "x"}).status_code == 503
        with TestClient(app) as client:                                 # lifespan: manager.load(), like main.py:37-57
            assert mgr.loads == 1 and mgr.model
            assert client.get("/v1/health").json()["status"] == "healthy"
            assert client.post("/v1/audio/speech", json={"text": "   "}).status_code == 400
            assert client.post("/v1/audio/speech", json={}).status_code == 422
            r = client.post("/v1/audio/speech/voice", json={"text": "x", "ref_audio_name": "nobody"})
            assert r.status_code == 400 and r.json()["detail"] == "Invalid reference audio name."
            r = client.post("/v1/audio/speech/voice", json={"text": "x", "ref_audio_name": "KAN_F (Happy)", "ref_text": " "})
            assert r.status_code == 400 and r.json()["detail"] == "Reference text cannot be empty."
            r = client.post("/v1/audio/speech", json={"text": "abcde"})
            assert r.status_code == 200 and r.headers["content-type"] == "audio/wav"
            assert r.headers["content-disposition"] == "attachment; filename=synthesized_kannada_speech.wav"
            assert float(r.headers["x-response-time"]) >= 0.1
            with wavmod.open(io.BytesIO(r.content)) as w:
                assert (w.getframerate(), w.getnchannels(), w.getsampwidth(), w.getnframes()) == (24000, 1, 2, 2400)
                pcm = np.frombuffer(w.readframes(2400), dtype="<i2")
            want = np.rint(np.float32(np.int16(0.005 * 32768)) / 32768.0 * 32767.0)
            assert (pcm == want).all()
            # six clients at once: the first is picked up alone or with company, the rest ride together (never six batches)
            res = {}

            def call(k):
                res[k] = client.post("/v1/audio/speech", json={"text": "y" * (k + 1)})

            ths = [threading.Thread(target=call, args=(k,)) for k in range(6)]
            [t.start() for t in ths]
            [t.join() for t in ths]
            assert all(res[k].status_code == 200 for k in range(6))
            for k in range(6):                                          # every caller got ITS text's audio
                with wavmod.open(io.BytesIO(res[k].content)) as w:
                    v = np.frombuffer(w.readframes(1), dtype="<i2")[0]
                assert v == np.rint(np.float32(np.int16(0.001 * (k + 1) * 32768)) / 32768.0 * 32767.0)
            batches = client.get("/v1/health").json()["batches"]
            assert sum(batches) == 7 and len(batches) <= 4
        assert mgr.model.prompt_calls == 7                              # the fake does not cache; INF5Model._prompt does (md5)
    finally:
        api._synthesizer_for = orig


def test_shard_plan_fuzz_covers_every_utterance_once():
    """`dist.shard_plan` on 300 seeded random request batches, including the corners a server meets (no utterances, fewer
    utterances than ranks, one 4096-frame utterance, tiny row budgets): every utterance lands on exactly one rank and in
    exactly one pack of that rank; ranks without work get empty lists (SURVEY §8e partitioning)."""
    import random
    from tts_indic_server_f5_b200.dist import shard_plan
    rnd = random.Random(0)
    for _ in range(300):
        n, w = rnd.randint(0, 40), rnd.choice([1, 2, 3, 4, 8])
        lens = [rnd.randint(1, 4096) for _ in range(n)]
        plan = shard_plan(lens, w, max_rows=rnd.choice([1024, 65536, 131072]))
        assert len(plan["parts"]) == len(plan["packs"]) == w
        assert sorted(i for part in plan["parts"] for i in part) == list(range(n))
        for r in range(w):
            assert sorted(j for pk in plan["packs"][r] for j in pk) == list(range(len(plan["parts"][r])))


def test_load_vocoder_reads_the_local_checkpoint_layout(tmp_path, monkeypatch):
    """`load_vocoder("vocos", is_local=True, local_path=...)` (utils_infer.py:100-115): `pytorch_model.bin` of
    charactr/vocos-mel-24khz carries the feature extractor's buffers next to backbone / head; the architecture is inferred from the
    tensors (dim 512, intermediate 1536, 8 layers), extra keys are ignored, bigvgan is refused.  The engine itself needs a B200,
    so it is replaced by a recorder here."""
    from tts_indic_server_f5_b200 import api
    sd = W.make_vocos_state_dict(W.VOCOS_24K, seed=3)
    extra = {"feature_extractor.mel_spec.spectrogram.window": torch.hann_window(1024),
             "feature_extractor.mel_spec.mel_scale.fb": torch.zeros(513, 100)}
    torch.save({**sd, **extra}, str(tmp_path / "pytorch_model.bin"))
    seen = {}
    monkeypatch.setattr(api, "Vocos", lambda state_dict, cfg, device, precision="bf16": seen.update(sd=state_dict, cfg=cfg, device=device,
                                                                                                  precision=precision) or "engine")
    assert api.load_vocoder("vocos", is_local=True, local_path=str(tmp_path), device="cuda:0") == "engine"
    assert (seen["cfg"].dim, seen["cfg"].intermediate_dim, seen["cfg"].num_layers) == (512, 1536, 8)
    assert seen["cfg"] == W.VOCOS_24K and seen["device"] == "cuda:0" and seen["precision"] == "bf16"
    assert all(torch.equal(seen["sd"][k], v) for k, v in sd.items())
    with pytest.raises(NotImplementedError):
        api.load_vocoder("bigvgan")


def test_layout_fuzz():
    """`build_layout` on 200 seeded random batches (1 - 48 utterances of 1 - 4096 frames): rows padded to 128, both CFG halves
    identical, every live row covered by exactly one attention query tile whose key range is its own utterance, gaps of at least
    GAP dead rows around every utterance (the k = 31 conv halo and the depthwise window read zeros there), segment table in order."""
    import random
    rnd = random.Random(4)
    for _ in range(200):
        lens = [rnd.choice([1, 2, 127, 128, 129, 255, 256, 257, rnd.randint(1, 4096)]) for _ in range(rnd.randint(1, 48))]
        L = build_layout(lens)
        R = L.half_rows
        assert R % 128 == 0 and L.rows == 2 * R and L.row_pos.numel() == 2 * R and L.lengths == lens
        pos = L.row_pos[:R]
        assert torch.equal(L.row_pos[R:], pos) and int((pos >= 0).sum()) == sum(lens) == L.real_tokens
        prev_end = 0
        for u, (s, n) in enumerate(zip(L.starts, lens)):
            assert s - prev_end >= GAP and torch.equal(pos[s:s + n], torch.arange(n, dtype=torch.int32))
            assert (pos[prev_end:s] == -1).all() and (L.row_utt[s:s + n] == u).all()
            prev_end = s + n
        assert (pos[prev_end:] == -1).all() and R - prev_end >= 0
        cover = torch.zeros(2 * R, dtype=torch.int32)
        for q0, kv0, kvl, qv in L.attn_tiles.tolist():
            cover[q0:q0 + qv] += 1
            assert kv0 <= q0 and q0 + qv <= kv0 + kvl and 1 <= qv <= 256
            assert int(L.row_pos[kv0]) == 0 and int(L.row_pos[kv0 + kvl - 1]) == kvl - 1
            assert kv0 + kvl == 2 * R or int(L.row_pos[kv0 + kvl]) == -1
        assert torch.equal(cover, (L.row_pos >= 0).to(torch.int32))
        assert L.seg_rows.tolist() == [[h + s, n] for h in (0, R) for s, n in zip(L.starts, lens)]
