"""CPU: the oracle (oracle/f5_oracle.py) against the golden vectors minted from the REAL reference (oracle/make_golden.py),
and, where /root/reference exists, against the reference's own modules executed in place."""
import os

import numpy as np
import pytest
import torch

from oracle import f5_oracle as O
from oracle import ref_shims as R
from tts_indic_server_f5_b200 import synthetic as S
from tts_indic_server_f5_b200 import text as T
from tts_indic_server_f5_b200 import weights as W


@pytest.fixture(scope="module")
def tiny():
    cfg, vcfg = W.tiny_dit_config(), W.tiny_vocos_config()
    return cfg, vcfg, W.make_dit_state_dict(cfg, seed=1), W.make_vocos_state_dict(vcfg, seed=1)


def _ids(spec):
    vocab = {t: i for i, t in enumerate(T.synthetic_indic_vocab())}
    ref_text = spec.ref_text + (" " if len(spec.ref_text[-1].encode()) == 1 else "")
    return O.list_str_to_idx(T.convert_char_to_pinyin([ref_text + spec.gen_text]), vocab)


def test_oracle_forward_matches_golden(tiny, golden_dir):
    cfg, _, sd, _ = tiny
    g = np.load(os.path.join(golden_dir, "tiny.npz"))
    x, cond, text = (torch.from_numpy(g[k])[None] for k in ("fwd_x", "fwd_condin", "fwd_text"))
    with torch.inference_mode():
        a = O.dit_forward(sd, cfg, x, cond, text, torch.tensor(0.37), False, False)[0].numpy()
        b = O.dit_forward(sd, cfg, x, cond, text, torch.tensor(0.37), True, True)[0].numpy()
    np.testing.assert_allclose(a, g["fwd_cond"], rtol=0, atol=2e-5)   # fp32 CPU, same op order: round-off only
    np.testing.assert_allclose(b, g["fwd_null"], rtol=0, atol=2e-5)


@pytest.mark.parametrize("wl,i", [("tiny", 0), ("tiny3", 1), ("tiny3", 2)])
def test_oracle_end_to_end_matches_golden(tiny, golden_dir, wl, i):
    cfg, vcfg, sd, vsd = tiny
    g = np.load(os.path.join(golden_dir, "tiny.npz"))
    spec = S.workload(wl)[i]
    with torch.inference_mode():
        wave, mel = O.infer_one(sd, cfg, vsd, vcfg, spec.audio, _ids(spec), spec.duration,
                                y0=S.initial_noise(4096, spec.noise_index))
    ref_len = spec.meta["ref_len"]
    np.testing.assert_allclose(mel.numpy().T, g[f"{wl}_{i}_mel"][ref_len:], rtol=0, atol=1e-4)
    np.testing.assert_allclose(wave.numpy(), g[f"{wl}_{i}_wave"], rtol=0, atol=1e-5)


def test_oracle_prompt_mel_matches_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "tiny.npz"))
    m = O.mel_spectrogram(S.prompt_audio(0.6, 0))[0].numpy()
    np.testing.assert_allclose(m, g["prompt_mel"], rtol=0, atol=1e-5)


def test_product_mel_tables_and_prompt_cache():
    """Host side of the prompt front-end (the arithmetic is the f5_mel_frames CUDA kernel, tested with -m gpu): the HTK
    filterbank equals torchaudio's (the reference's MelSpectrogram, modules.py:83-93), the non-zero bin ranges cover
    every non-zero weight, the cache is LRU and keyed by the samples, and there is no CPU path."""
    from tts_indic_server_f5_b200 import melspec as M
    fb = M.htk_fbank(513, 100, 24000)
    try:
        import torchaudio
        ref = torchaudio.functional.melscale_fbanks(513, 0.0, 12000.0, 100, 24000, norm=None, mel_scale="htk")
        assert torch.equal(fb, ref)
    except ImportError:
        pass
    assert torch.equal(fb, O.mel_filterbank_htk(513, 100, 24000))
    band = M.band_ranges(fb)
    for m in range(100):
        k0, k1 = int(band[m, 0]), int(band[m, 1])
        assert k1 > k0 and float(fb[:k0, m].abs().sum()) == 0.0 and float(fb[k1:, m].abs().sum()) == 0.0
    c = M.PromptCache(capacity=2)
    a, b, d = S.prompt_audio(0.3, 1), S.prompt_audio(0.3, 2), S.prompt_audio(0.3, 3)
    ka, kb, kd = M.PromptCache.key(a), M.PromptCache.key(b), M.PromptCache.key(d)
    assert ka != kb and ka == M.PromptCache.key(a.clone())
    c.put(ka, "cuda:0", "A"); c.put(kb, "cuda:0", "B")
    assert c.get(ka, "cuda:0") == "A" and c.get(ka, "cuda:1") is None
    c.put(kd, "cuda:0", "D")                       # evicts B (least recently used)
    assert c.get(kb, "cuda:0") is None and c.get(ka, "cuda:0") == "A" and c.get(kd, "cuda:0") == "D"
    with pytest.raises(RuntimeError):
        M.mel_spectrogram(a)                       # CPU tensor: the product path has no CPU fallback


@pytest.mark.skipif(not R.reference_available(), reason="reference tree only exists in the build container")
def test_oracle_vs_real_reference_modules(tiny):
    """Same weights, same inputs, reference's own DiT / CFM.sample vs the restatement."""
    cfg, _, sd, _ = tiny
    cfm = R.build_reference_cfm(sd, cfg)
    g = torch.Generator().manual_seed(3)
    n, F_ = 70, 30
    cond = torch.randn(1, F_, 100, generator=g)
    text = torch.randint(0, cfg.vocab_size, (1, 41), generator=g)
    with torch.inference_mode():
        ref, _ = cfm.sample(cond=cond, text=text, duration=n, steps=6, cfg_strength=2.0, sway_sampling_coef=-1.0, seed=11)
        ours = O.cfm_sample(sd, cfg, cond, text, n, steps=6, seed=11)
    assert torch.equal(ref, ours)
    # text longer than the prompt: lens = max(text_lens, lens) (cfm.py:123-125), duration = lens + 1 floor (:136)
    text2 = torch.randint(0, cfg.vocab_size, (1, 50), generator=g)
    with torch.inference_mode():
        ref, _ = cfm.sample(cond=cond[:, :20], text=text2, duration=10, steps=3, cfg_strength=2.0, sway_sampling_coef=-1.0, seed=5)
        ours = O.cfm_sample(sd, cfg, cond[:, :20], text2, 10, steps=3, seed=5)
    assert ref.shape[1] == 51 and torch.equal(ref, ours)


@pytest.mark.skipif(not R.reference_available(), reason="reference tree only exists in the build container")
def test_text_frontend_vs_reference():
    ref = R.load_reference()
    texts = [S.workload("tiny3")[1].ref_text + "ನಮಸ್ಕಾರ, हिन्दी! ", "ಕನ್ನಡ; “quote” ‘x’ ಪದ?", "अ आ इ. ई"]
    assert T.convert_char_to_pinyin(texts) == ref.model_utils.convert_char_to_pinyin(texts)
    long = " ".join(T.synthetic_indic_text(40, i) + "." for i in range(12))
    for mc in (60, 135, 400):
        assert T.chunk_text(long, mc) == ref.utils_infer.chunk_text(long, mc)
    vocab = {t: i for i, t in enumerate(T.synthetic_indic_vocab())}
    toks = T.convert_char_to_pinyin(texts)
    assert torch.equal(T.list_str_to_idx(toks, vocab), ref.model_utils.list_str_to_idx(toks, vocab))


@pytest.mark.skipif(not R.reference_available(), reason="reference tree only exists in the build container")
def test_chunk_text_randomised_vs_reference():
    """`chunk_text` (utils_infer.py:61-88: sentence split on `; : , . ! ?` and their full-width forms, byte-length budget, the
    single-byte-ending space rule) is pure Python in the reference: 300 seeded random texts mixing Kannada / Devanagari code
    points, ASCII words, every delimiter and stray whitespace, at five budgets, must chunk identically."""
    import random
    ref = R.load_reference()
    rnd = random.Random(7)
    kan = [chr(c) for c in range(0x0C85, 0x0CB9)] + ["\u0ccd", "\u0cbe", "\u200c"]
    dev = [chr(c) for c in range(0x0905, 0x0939)] + ["\u094d", "\u093e"]
    delim = list(";:,.!?") + ["；", "：", "，", "。", "！", "？"]
    for case in range(300):
        parts = []
        for _ in range(rnd.randint(1, 40)):
            kind = rnd.random()
            if kind < 0.45:
                parts.append("".join(rnd.choice(kan) for _ in range(rnd.randint(1, 9))))
            elif kind < 0.75:
                parts.append("".join(rnd.choice(dev) for _ in range(rnd.randint(1, 9))))
            elif kind < 0.9:
                parts.append("".join(rnd.choice("abcXYZ019") for _ in range(rnd.randint(1, 6))))
            else:
                parts.append(rnd.choice(["", " ", "  ", "\n"]))
            r = rnd.random()
            parts.append(rnd.choice(delim) + rnd.choice(["", " ", "  "]) if r < 0.35 else (" " if r < 0.9 else ""))
        text = "".join(parts)
        for mc in (1, 17, 60, 135, 1000):
            assert T.chunk_text(text, mc) == ref.utils_infer.chunk_text(text, mc), (case, mc, text)


@pytest.mark.skipif(os.environ.get("F5_SKIP_SLOW") == "1", reason="~1 min of CPU")
def test_oracle_full_size_forward_is_finite_and_param_count():
    """Architecture pin: 337.10 M parameters with the vendored 2545-entry vocab."""
    cfg = W.DiTConfig(vocab_size=2545)
    sd = W.make_dit_state_dict(cfg, seed=0)
    n_params = sum(v.numel() for v in sd.values())
    assert abs(n_params - 337.10e6) < 0.02e6, n_params


def test_vocos_architecture_pin():
    """vocos 0.1.0 is not in the reference tree (parity unpinned), so what CAN be pinned is pinned: the published
    size of `charactr/vocos-mel-24khz` (13.5 M parameters) and the state-dict layout its
    `pytorch_model.bin` has (`backbone.embed`, `backbone.norm`, 8 x `backbone.convnext.{i}.{dwconv,norm,pwconv1,pwconv2,gamma}`,
    `backbone.final_layer_norm`, `head.out`, `head.istft.window`), which is what `load_vocoder(local_path=...)` consumes
    (utils_infer.py:104-114), plus the 26.99 MFLOP per frame the roofline uses."""
    vc = W.VOCOS_24K
    sd = W.make_vocos_state_dict(vc, seed=0)
    n_params = sum(v.numel() for k, v in sd.items() if k != "head.istft.window")
    assert n_params == 13_531_650
    want = {"backbone.embed.weight": (512, 100, 7), "backbone.embed.bias": (512,), "backbone.norm.weight": (512,),
            "backbone.norm.bias": (512,), "backbone.final_layer_norm.weight": (512,), "backbone.final_layer_norm.bias": (512,),
            "head.out.weight": (1026, 512), "head.out.bias": (1026,), "head.istft.window": (1024,)}
    for i in range(8):
        p = f"backbone.convnext.{i}."
        want.update({p + "dwconv.weight": (512, 1, 7), p + "dwconv.bias": (512,), p + "norm.weight": (512,), p + "norm.bias": (512,),
                     p + "pwconv1.weight": (1536, 512), p + "pwconv1.bias": (1536,), p + "pwconv2.weight": (512, 1536),
                     p + "pwconv2.bias": (512,), p + "gamma": (512,)})
    assert {k: tuple(v.shape) for k, v in sd.items()} == want
    assert torch.equal(sd["head.istft.window"], torch.hann_window(1024))
    macs = 100 * 512 * 7 + 8 * (512 * 7 + 2 * 512 * 1536) + 512 * 1026          # embed + 8 blocks + head, per frame
    assert abs(2 * macs - 26.99e6) < 0.01e6


def test_third_party_restatements_have_the_properties_of_what_they_restate():
    """x-transformers RoPE and torchdiffeq Euler are not in the reference tree (parity unpinned); these are the properties any
    faithful restatement must have, checked in float64 against INDEPENDENT formulations: RoPE on interleaved pairs is the complex
    rotation (x0 + i x1) e^{i n theta_j} with theta_j = 10000^(-2j/64), it leaves channels >= 64 alone, and
    q.k after rotation depends only on the position difference; fixed-grid Euler of y' = A y is the ordered product of
    (I + dt_k A) and returns every grid state (Appendix A.2)."""
    g = torch.Generator().manual_seed(3)
    n, d = 37, 64
    x = torch.randn(2, n, 80, generator=g)                                               # fp32, like the model's q / k
    y = O.apply_rotary_pos_emb(x, O.rotary_freqs(n, d))
    assert y.dtype == x.dtype and torch.equal(y[..., d:], x[..., d:])
    theta = 10000.0 ** (-torch.arange(0, d, 2, dtype=torch.float64) / d)
    ang = torch.arange(n, dtype=torch.float64)[:, None] * theta[None]                      # [n, 32]
    xd = x.double()
    z = torch.complex(xd[..., 0:d:2], xd[..., 1:d:2]) * torch.polar(torch.ones_like(ang), ang)
    assert (y[..., 0:d:2] - z.real).abs().max() < 1e-5 and (y[..., 1:d:2] - z.imag).abs().max() < 1e-5   # fp32 math (A.1)
    q, k = torch.randn(d, generator=g), torch.randn(d, generator=g)
    big = O.rotary_freqs(64, d)[0]

    def rot(v, pos):
        return O.apply_rotary_pos_emb(v[None, None, :], big[None, pos:pos + 1])[0, 0]

    dots = [float(rot(q, m) @ rot(k, m + 5)) for m in (0, 7, 31, 58)]
    assert max(dots) - min(dots) < 1e-4 * max(1.0, abs(dots[0]))
    A = torch.randn(4, 4, generator=g, dtype=torch.float64) * 0.3
    y0 = torch.randn(4, generator=g, dtype=torch.float64)
    t = O.sway_time_grid(8, -1.0, dtype=torch.float64)
    ys = O.odeint_euler(lambda t0, yy: A @ yy, y0, t)
    assert ys.shape == (9, 4) and torch.equal(ys[0], y0)
    want = y0.clone()
    for k_ in range(8):
        want = (torch.eye(4, dtype=torch.float64) + (t[k_ + 1] - t[k_]) * A) @ want
        assert (ys[k_ + 1] - want).abs().max() < 1e-12


@pytest.mark.skipif(not R.reference_available(), reason="reference tree only exists in the build container")
def test_driver_host_logic_vs_the_real_infer_batch_process():
    """Everything `infer_batch_process` (utils_infer.py:406-524) does AROUND `model.sample` / `vocoder.decode`, run through the
    REAL reference function with recording stubs for the two model objects, against the product's host side (`Synthesizer._prep`,
    the gain rule, `cross_fade`): mono mix, RMS boost of quiet prompts, torchaudio resample of non-24 kHz prompts, the
    single-byte-ending space rule, tokens, `ref_audio_len`, the byte-ratio / fix_duration / speed duration rules, the prompt
    strip, RMS un-scaling and the 0.15 s cross-fade of the chunk waves."""
    from tts_indic_server_f5_b200 import api
    from tts_indic_server_f5_b200.scheduler import cross_fade
    ref = R.load_reference()

    class StubModel:
        def __init__(self):
            self.calls = []

        def sample(self, cond, text, duration, steps, cfg_strength, sway_sampling_coef):
            self.calls.append(dict(cond=cond.clone(), text=text, duration=duration, steps=steps, cfg=cfg_strength, sway=sway_sampling_coef))
            n = max(int(duration), cond.shape[-1] // 256 + 3)
            return torch.randn(1, n, 100, generator=torch.Generator().manual_seed(len(self.calls))), None

    class StubVocoder:
        def decode(self, mel):
            F_ = mel.shape[-1]
            return torch.randn(1, 256 * (F_ - 1), generator=torch.Generator().manual_seed(F_)) * 0.1

    g = torch.Generator().manual_seed(11)
    cases = [   # (channels, rate, seconds, amplitude, ref_text, speed, fix_duration)
        (1, 24000, 3.0, 0.3, T.synthetic_indic_text(30, 1) + ".", 1.0, None),        # ends in a 1-byte char: the space rule
        (2, 16000, 2.2, 0.02, T.synthetic_indic_text(24, 2), 0.8, None),              # stereo, quiet (RMS boost), resampled
        (1, 44100, 1.7, 0.5, T.synthetic_indic_text(20, 3) + ". ", 1.3, None),
        (1, 24000, 2.5, 0.05, T.synthetic_indic_text(26, 4), 1.0, 9.5),               # fix_duration
    ]
    for ci, (ch, rate, sec, amp, ref_text, speed_, fixd) in enumerate(cases):
        audio = torch.randn(ch, int(sec * rate), generator=g) * amp
        chunks = [T.synthetic_indic_text(40 + 13 * k, 50 + 7 * ci + k) + "." for k in range(3)]
        m, v = StubModel(), StubVocoder()
        want_wave, want_sr, want_mel = ref.utils_infer.infer_batch_process((audio, rate), ref_text, chunks, m, v, speed=speed_,
                                                                           fix_duration=fixd, device=None)
        assert want_sr == 24000 and len(m.calls) == 3
        waves = []
        for k, (chunk, call) in enumerate(zip(chunks, m.calls)):
            spec = S.UtteranceSpec(audio=audio, ref_text=ref_text, gen_text=chunk, duration=None, noise_index=k, meta={"sr": rate})
            p = api.Synthesizer._prep(None, spec, speed_, fixd)
            assert torch.equal(p.audio, call["cond"]), (ci, k)                   # mono mix, RMS boost, resample: same samples
            assert p.tokens == call["text"][0] and p.duration == call["duration"], (ci, k)
            assert p.ref_len == call["cond"].shape[-1] // 256
            assert (call["steps"], call["cfg"], call["sway"]) == (api.nfe_step, api.cfg_strength, api.sway_sampling_coef)
            # what the engine does after sampling: strip the prompt frames, vocode, undo the RMS boost (gain fused into the ISTFT)
            n = max(p.duration, p.audio.shape[-1] // 256 + 3)
            mel = torch.randn(1, n, 100, generator=torch.Generator().manual_seed(k + 1))[:, p.ref_len:, :].permute(0, 2, 1)
            wave = v.decode(mel)
            gain = p.rms / api.target_rms if p.rms < api.target_rms else 1.0
            if gain != 1.0:
                wave = wave * torch.tensor(gain, dtype=torch.float32)
            waves.append(wave.squeeze().numpy())
            assert np.array_equal(want_mel[:, sum(w_.shape[0] // 256 + 1 for w_ in waves[:-1]):][:, :mel.shape[-1]], mel[0].numpy())
        got = cross_fade(waves, 0.15, 24000)
        assert got.shape == want_wave.shape
        np.testing.assert_allclose(got, want_wave, rtol=0, atol=2e-7)             # the gain is one fp32 multiply either way


@pytest.mark.skipif(not R.reference_available(), reason="reference tree only exists in the build container")
def test_sample_prologue_fuzz_vs_the_real_cfm_sample(tiny):
    """`api.CFM._prepare` (the host half of `CFM.sample`, cfm.py:100-186) against the REAL reference method with a recording stub
    in place of its transformer: 60 seeded random calls (batch 1-4, `lens` given or not, int / tensor durations incl. values
    below lens + 1 and above `max_duration`, -1-padded text rows longer or shorter than the prompt, optional `edit_mask`, seeded
    noise).  What the reference hands its transformer at the first evaluation — x = y0, the masked `step_cond`, the row mask — must
    be what the product's per-utterance inputs describe."""
    from types import SimpleNamespace
    from tts_indic_server_f5_b200.api import CFM
    cfg, sd = tiny[0], tiny[2]
    cfm = R.build_reference_cfm(sd, cfg, None)
    seen = {}

    class Recorder(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.p = torch.nn.Parameter(torch.zeros(1))            # CFM.sample reads the model dtype off its first parameter

        def forward(self, x, cond, text, time, mask, drop_audio_cond, drop_text):
            if not seen:
                seen.update(x=x.clone(), cond=cond.clone(), text=text.clone(), mask=None if mask is None else mask.clone())
            return torch.zeros_like(x)

    cfm.transformer = Recorder()
    stub = SimpleNamespace(vocab_char_map=None, num_channels=cfg.mel_dim, _device="cpu")
    g = torch.Generator().manual_seed(21)
    ri = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=g))  # noqa: E731
    ran = 0
    for case in range(60):
        b, F_ = ri(1, 4), ri(5, 40)
        cond = torch.randn(b, F_, cfg.mel_dim, generator=g)
        nt = ri(1, 60)
        text = torch.randint(0, cfg.vocab_size, (b, nt), generator=g)
        for i in range(b):
            text[i, ri(1, nt):] = -1
        lens = None if ri(0, 1) else torch.tensor([ri(1, F_) for _ in range(b)])
        max_dur = ri(30, 90)
        duration = ri(1, 120) if ri(0, 1) else torch.tensor([ri(1, 120) for _ in range(b)])
        edit_mask = None
        if ri(0, 2) == 0:
            edit_mask = torch.rand(b, F_ if lens is None else int(max(int(lens.max()), int((text != -1).sum(-1).max()))), generator=g) > 0.3
        seed = ri(0, 1000)
        seen.clear()
        try:
            out, _ = cfm.sample(cond=cond, text=text.clone(), duration=duration, lens=None if lens is None else lens.clone(), steps=1,
                                cfg_strength=2.0, sway_sampling_coef=-1.0, seed=seed, max_duration=max_dur, edit_mask=edit_mask)
        except RuntimeError:
            continue                      # shapes the reference itself cannot broadcast (edit_mask narrower than its cond_mask, cond longer than max_duration)
        ran += 1
        utts = CFM._prepare(stub, cond, text.clone(), duration, lens, seed, max_dur, edit_mask, None)
        assert len(utts) == b
        N = seen["x"].shape[1]
        for i, u in enumerate(utts):
            n_ref = N if seen["mask"] is None else int(seen["mask"][i].sum())
            assert u.n == n_ref, (case, i)
            assert torch.equal(u.y0, seen["x"][i, :u.n]), (case, i)                   # the seeded CPU draw, per item
            assert (seen["x"][i, u.n:] == 0).all()
            assert torch.equal(u.text_ids, text[i][text[i] != -1])
            # step_cond as the engine builds it from (cond, cond_len, edit_mask): engine.upload
            m = torch.zeros(u.n, dtype=torch.bool)
            m[: min(u.cond_len, u.n)] = True
            if u.edit_mask is not None:
                em = u.edit_mask.bool()[: u.n]
                m[: em.numel()] &= em
            c = torch.zeros(u.n, cfg.mel_dim)
            k = min(u.cond.shape[0], u.n)
            c[:k] = u.cond[:k]
            assert torch.equal(torch.where(m[:, None], c, torch.zeros_like(c)), seen["cond"][i, :u.n]), (case, i)
            assert (seen["cond"][i, u.n:] == 0).all()
    print(f"prologue fuzz: {ran} of 60 calls accepted by the reference")
    assert ran >= 40


@pytest.mark.skipif(not R.reference_available(), reason="reference tree only exists in the build container")
def test_get_tokenizer_on_the_vendored_vocab_file():
    """`get_tokenizer(path, "custom")` (model/utils.py:124-129) on the vocab file the reference ships
    (`f5_tts/infer/examples/vocab.txt`): same map and size as the reference's own function; then ids of mixed Indic text through
    both `list_str_to_idx` implementations with that map (characters outside the file fall back to 0)."""
    ref = R.load_reference()
    path = os.path.join(R.REFERENCE_ROOT, "src", "server", "f5_tts", "infer", "examples", "vocab.txt")
    got_map, got_size = T.get_tokenizer(path, "custom")
    want_map, want_size = ref.model_utils.get_tokenizer(path, "custom")
    assert got_size == want_size == 2545 and got_map == want_map
    texts = [T.synthetic_indic_text(50, s, script) + ", abc 12?" for s, script in ((1, "kannada"), (2, "devanagari"))]
    toks = T.convert_char_to_pinyin(texts)
    assert torch.equal(T.list_str_to_idx(toks, got_map), ref.model_utils.list_str_to_idx(toks, want_map))


@pytest.mark.skipif(not R.reference_available(), reason="reference tree only exists in the build container")
def test_boundary_signatures_match_the_reference_source():
    """Drop-in surface, checked against the reference's SOURCE (parsed, not imported: `core/managers.py` pulls in the LLM stack):
    `TTSManager` has the reference's methods with the reference's positional parameters, its instance attributes `device_type`,
    `model`, `repo_id` and the same error for an unloaded model; `infer_process`, `infer_batch_process`, `load_model`,
    `load_vocoder`, `preprocess_ref_audio_text`, `chunk_text` accept every parameter of the reference's functions under the
    same names, in the same order (extra keyword-only additions come after them), with the same defaults for the sampler constants."""
    import ast
    import inspect
    from tts_indic_server_f5_b200 import api
    root = os.path.join(R.REFERENCE_ROOT, "src", "server")
    mgr_src = ast.parse(open(os.path.join(root, "core", "managers.py"), encoding="utf-8").read())
    cls = next(n for n in ast.walk(mgr_src) if isinstance(n, ast.ClassDef) and n.name == "TTSManager")
    ref_methods = {f.name: [a.arg for a in f.args.args] for f in cls.body if isinstance(f, ast.FunctionDef)}
    assert set(ref_methods) == {"__init__", "load", "synthesize"}
    for name, params in ref_methods.items():
        ours = list(inspect.signature(getattr(api.TTSManager, name)).parameters)
        assert ours[: len(params)] == params, (name, ours, params)
    attrs = {t.attr for n in ast.walk(cls) if isinstance(n, ast.Assign) for t in n.targets if isinstance(t, ast.Attribute)}
    m = api.TTSManager(device_type="cuda")
    assert attrs <= set(vars(m)) and m.model is None and m.repo_id == "ai4bharat/IndicF5" and m.device_type == "cuda"
    with pytest.raises(ValueError, match="TTS model not loaded"):
        m.synthesize("x", ref_audio_path="p", ref_text="r")
    ui_src = ast.parse(open(os.path.join(root, "f5_tts", "infer", "utils_infer.py"), encoding="utf-8").read())
    ref_fns = {n.name: n for n in ui_src.body if isinstance(n, ast.FunctionDef)}
    consts = {t.id: ast.literal_eval(n.value) for n in ui_src.body if isinstance(n, ast.Assign) for t in n.targets
              if isinstance(t, ast.Name) and isinstance(n.value, (ast.Constant, ast.UnaryOp))}
    for k in ("target_sample_rate", "n_mel_channels", "hop_length", "win_length", "n_fft", "mel_spec_type", "target_rms",
              "cross_fade_duration", "ode_method", "nfe_step", "cfg_strength", "sway_sampling_coef", "speed", "fix_duration"):
        assert getattr(api, k) == consts[k], k                                # utils_infer.py:40-53
    for name in ("infer_process", "infer_batch_process", "load_model", "load_vocoder", "preprocess_ref_audio_text", "chunk_text"):
        want = [a.arg for a in ref_fns[name].args.args]
        fn = getattr(api, name, None) or getattr(T, name)
        got = list(inspect.signature(fn).parameters)
        assert got[: len(want)] == want, (name, got, want)
    cfm_src = ast.parse(open(os.path.join(root, "f5_tts", "model", "cfm.py"), encoding="utf-8").read())
    sample = next(n for n in ast.walk(cfm_src) if isinstance(n, ast.FunctionDef) and n.name == "sample")
    ours = inspect.signature(api.CFM.sample).parameters
    assert list(ours)[: len(sample.args.args)] == [a.arg for a in sample.args.args]          # self, cond, text, duration
    for a_, d_ in zip(sample.args.kwonlyargs, sample.args.kw_defaults):                         # lens ... edit_mask, same defaults
        assert a_.arg in ours and ours[a_.arg].kind is inspect.Parameter.KEYWORD_ONLY, a_.arg
        assert ours[a_.arg].default == ast.literal_eval(d_), a_.arg


@pytest.mark.skipif(not R.reference_available(), reason="reference tree only exists in the build container")
def test_oracle_fuzz_vs_the_real_reference(tiny):
    """The parity anchor itself, on 16 seeded random shapes: the reference's own `DiT.forward` (both CFG branches, random time) and
    `CFM.sample` (text shorter / longer than the prompt, durations at the lens + 1 floor, `lens` shorter than the prompt, sway on /
    off, 1 - 5 steps) against the restatement — bit for bit, trajectory included."""
    cfg, _, sd, _ = tiny
    cfm = R.build_reference_cfm(sd, cfg)
    g = torch.Generator().manual_seed(17)
    ri = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=g))  # noqa: E731
    with torch.inference_mode():
        for case in range(16):
            n, nt = ri(3, 160), ri(1, 180)
            x, cond = torch.randn(1, n, cfg.mel_dim, generator=g), torch.randn(1, n, cfg.mel_dim, generator=g)
            text = torch.randint(0, cfg.vocab_size, (1, nt), generator=g)
            text[0, ri(1, nt):] = -1
            t = torch.rand((), generator=g)
            for drop in (False, True):
                want = cfm.transformer(x=x, cond=cond, text=text, time=t, mask=None, drop_audio_cond=drop, drop_text=drop)
                assert torch.equal(O.dit_forward(sd, cfg, x, cond, text, t, drop, drop), want), (case, drop)
            F_ = ri(2, 60)
            prompt = torch.randn(1, F_, cfg.mel_dim, generator=g)
            dur, steps = ri(1, 140), ri(1, 5)
            lens = None if ri(0, 1) else torch.tensor([ri(1, F_)])
            sway = -1.0 if ri(0, 1) else None
            seed = ri(0, 99)
            want, want_traj = cfm.sample(cond=prompt, text=text.clone(), duration=dur, lens=None if lens is None else lens.clone(), steps=steps,
                                         cfg_strength=2.0, sway_sampling_coef=sway, seed=seed)
            got, traj = O.cfm_sample(sd, cfg, prompt, text, dur, steps=steps, sway_sampling_coef=sway, seed=seed, lens=lens,
                                     return_trajectory=True)
            assert torch.equal(got, want) and torch.equal(traj, want_traj), case
        modules = R.load_reference().modules                       # prompt mel (model/modules.py:75-101) at random lengths / levels
        for k in range(6):
            wave_ = torch.randn(ri(1, 2), ri(600, 60000), generator=g) * (10.0 ** -ri(0, 3))
            want = modules.get_vocos_mel_spectrogram(wave_)
            assert torch.equal(O.mel_spectrogram(wave_), want), k
