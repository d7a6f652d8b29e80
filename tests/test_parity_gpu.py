"""GPU parity tests proper: the CUDA path, called through the C ABI (ops -> libf5b200.so), against
  * the golden vectors minted from the REAL reference (tests/golden/*.npz, oracle/make_golden.py),
  * the oracle (CPU fp32 restatement) on seeded inputs at sizes it finishes in seconds,
  * size-independent properties at the full IndicF5 size (packing invariance, graph == eager, ISTFT perfect reconstruction).

Tolerances (bf16 tensor-core operands; fp32 accumulation, residual stream, LayerNorm, softmax, ODE state, ISTFT):
  mel   rel-L2 <= 1.0e-2, L-inf <= 6e-2 on a +-5 range    (measured 3.1e-3..3.9e-3 / 0.013..0.019)
  wave  SNR    >= 40 dB                                    (measured 44.7..50.1 dB)
For scale: the reference's OWN bf16 mode (whole-module .to(bfloat16)) sits at rel-L2 1.33e-2, L-inf 0.156 from its fp32
result on the tiny case (tests/golden/ref_bf16_error.json) — the CUDA path must be at least that close, and is ~4x closer.
Benchmark sizes (round 2; goldens minted from the real reference by `python -m oracle.make_golden --sizes`):
  full IndicF5 forward pair at n = 1384 (C2) and n = 3069 (C3)   rel-L2 <= 1.0e-2 per branch
  two complete NFE-32 utterances INSIDE the packed C2 batch of 64    same mel / wave tolerances as above."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

MEL_REL, MEL_LINF, WAVE_SNR = 1.0e-2, 6e-2, 40.0

if torch.cuda.is_available():
    from oracle import f5_oracle as O
    from tts_indic_server_f5_b200 import api, ops, synthetic as S, text as T, weights as W
    from tts_indic_server_f5_b200.engine import UtteranceInput


def ref_noise(specs):
    """The noise the golden vectors were minted with (oracle/make_golden.py): CPU generator, seed 1234 + noise_index."""
    return [S.initial_noise(4096, s.noise_index) for s in specs]


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def snr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(10 * np.log10((b ** 2).sum() / max(((a - b) ** 2).sum(), 1e-30)))


@pytest.fixture(scope="module")
def tiny_models():
    cfg, vcfg = W.tiny_dit_config(), W.tiny_vocos_config()
    sd, vsd = W.make_dit_state_dict(cfg, seed=1), W.make_vocos_state_dict(vcfg, seed=1)
    return cfg, vcfg, sd, vsd, api.load_model(state_dict=sd), api.load_vocoder(state_dict=vsd)


@pytest.fixture(scope="module")
def full_models():
    model = api.load_model(state_dict=W.make_dit_state_dict(W.INDICF5, seed=0))
    voc = api.load_vocoder(state_dict=W.make_vocos_state_dict(W.VOCOS_24K, seed=0))
    return model, voc


def test_native_library_is_the_path():
    from tts_indic_server_f5_b200 import _lib
    assert _lib.lib.f5_device_check() == 0 and os.path.basename(_lib.LIB) == "libf5b200.so"


def test_tiny_forward_pair_vs_reference_golden(tiny_models, golden_dir):
    *_, model, _ = tiny_models
    g = np.load(os.path.join(golden_dir, "tiny.npz"))
    u = UtteranceInput(cond=torch.from_numpy(g["fwd_condin"]), text_ids=torch.from_numpy(g["fwd_text"]), n=96, cond_len=96,
                       y0=torch.from_numpy(g["fwd_x"]))
    pc = model.engine.forward_flow([u], 0.37)[0].cpu().numpy()
    assert rel(pc[0], g["fwd_cond"]) < MEL_REL and rel(pc[1], g["fwd_null"]) < MEL_REL
    assert np.abs(pc[0] - g["fwd_cond"]).max() < MEL_LINF


@pytest.mark.parametrize("wl", ["tiny", "tiny3"])
def test_tiny_end_to_end_vs_reference_golden(tiny_models, golden_dir, wl):
    *_, model, voc = tiny_models
    g = np.load(os.path.join(golden_dir, "tiny.npz"))
    bf16_ref = json.load(open(os.path.join(golden_dir, "ref_bf16_error.json")))["tiny"]
    specs = S.workload(wl)
    waves, mels = api.Synthesizer(model, voc).generate(specs, return_mel=True, y0=ref_noise(specs))
    for i, (spec, wv, ml) in enumerate(zip(specs, waves, mels)):
        gm = g[f"{wl}_{i}_mel"][spec.meta["ref_len"]:]
        assert ml.shape == gm.T.shape and wv.shape == g[f"{wl}_{i}_wave"].shape
        r, li, s = rel(ml.T, gm), float(np.abs(ml.T - gm).max()), snr(wv, g[f"{wl}_{i}_wave"])
        print(f"{wl}[{i}] mel rel-L2 {r:.2e} L-inf {li:.3f} wave SNR {s:.1f} dB")
        assert r < MEL_REL and li < MEL_LINF and s > WAVE_SNR
        assert r < bf16_ref["mel_rel_l2"] and li < bf16_ref["mel_linf"]      # no worse than the reference's own bf16 mode


def test_full_size_c1_vs_reference_golden(full_models, golden_dir):
    model, voc = full_models
    g = np.load(os.path.join(golden_dir, "full_c1.npz"))
    spec = S.workload("c1")
    waves, mels = api.Synthesizer(model, voc).generate(spec, return_mel=True, y0=ref_noise(spec))
    gm = g["mel"][spec[0].meta["ref_len"]:]
    r, li, s = rel(mels[0].T, gm), float(np.abs(mels[0].T - gm).max()), snr(waves[0], g["wave"])
    print(f"full C1: mel rel-L2 {r:.2e} L-inf {li:.3f} wave SNR {s:.1f} dB")
    assert r < MEL_REL and li < MEL_LINF and s > WAVE_SNR
    wv = voc.decode(torch.from_numpy(gm.T[None].copy()))[0].cpu().numpy()      # vocoder alone on the reference's mel
    assert snr(wv, g["wave"]) > 43.0


def test_cfm_sample_api_vs_oracle(tiny_models):
    """`CFM.sample` surface (cfm.py:82-99): raw-wave cond, list-of-str text, seed, lens/duration rules, edit_mask."""
    cfg, _, sd, _, model, _ = tiny_models
    vocab = {t: i for i, t in enumerate(T.synthetic_indic_vocab())}
    audio = S.prompt_audio(0.5, 1)
    texts = [T.convert_char_to_pinyin([T.synthetic_indic_text(30, 4)])[0]]
    out, traj = model.sample(cond=audio, text=texts, duration=90, steps=8, cfg_strength=2.0, sway_sampling_coef=-1.0, seed=3)
    ids = O.list_str_to_idx(texts, vocab)
    with torch.inference_mode():
        ref = O.cfm_sample(sd, cfg, O.mel_spectrogram(audio).permute(0, 2, 1), ids, 90, steps=8, seed=3)
    assert out.shape == ref.shape == (1, 90, 100) and traj.shape == (1, 1, 90, 100)
    assert rel(out.cpu().numpy(), ref.numpy()) < MEL_REL
    # text longer than the mel: cond_mask extends, duration floor lens + 1 (cfm.py:123-137)
    long_text = [T.convert_char_to_pinyin([T.synthetic_indic_text(70, 5)])[0]]
    out2, _ = model.sample(cond=audio, text=long_text, duration=10, steps=4, cfg_strength=2.0, seed=1)
    with torch.inference_mode():
        ref2 = O.cfm_sample(sd, cfg, O.mel_spectrogram(audio).permute(0, 2, 1), O.list_str_to_idx(long_text, vocab), 10, steps=4,
                            seed=1, sway_sampling_coef=None)
    assert out2.shape == ref2.shape == (1, 71, 100) and rel(out2.cpu().numpy(), ref2.numpy()) < MEL_REL
    # batch of two ragged items == each item alone (batch-1 semantics), zero rows past an item's duration
    mel = O.mel_spectrogram(audio).permute(0, 2, 1)
    y0 = [S.initial_noise(120, 0), S.initial_noise(120, 1)]
    outb, _ = model.sample(cond=torch.cat([mel, mel]), text=[texts[0], long_text[0]], duration=torch.tensor([80, 110]), steps=4,
                           cfg_strength=2.0, sway_sampling_coef=-1.0, y0=y0)
    o0, _ = model.sample(cond=mel, text=[texts[0]], duration=80, steps=4, cfg_strength=2.0, sway_sampling_coef=-1.0, y0=y0[:1])
    assert outb.shape == (2, 110, 100) and float(outb[0, 80:].abs().max()) == 0.0
    assert torch.equal(outb[0, :80], o0[0])                                  # deterministic, row-local kernels: bit-exact
    for bad in (dict(duplicate_test=True), dict(cfg_strength=0.0)):
        with pytest.raises(NotImplementedError):
            model.sample(cond=mel, text=[texts[0]], duration=80, **bad)


def test_vocos_decode_vs_oracle(tiny_models):
    _, vcfg, _, vsd, _, voc = tiny_models
    g = torch.Generator().manual_seed(9)
    for B, T_ in ((1, 2), (3, 257), (2, 64)):
        mel = torch.randn(B, 100, T_, generator=g) * 2 - 4
        got = voc.decode(mel).cpu().numpy()
        with torch.inference_mode():
            ref = O.vocos_decode(vsd, vcfg, mel).numpy()
        assert got.shape == ref.shape == (B, 256 * (T_ - 1))
        assert snr(got, ref) > WAVE_SNR, (B, T_, snr(got, ref))


def test_prompt_mel_kernel_vs_oracle_and_golden(golden_dir):
    """f5_mel_frames (reflect pad, hann, 1024-pt FFT, HTK mel, log) against the oracle's torch.stft restatement of the
    reference's get_vocos_mel_spectrogram and against the prompt mel the REAL reference produced (tests/golden/tiny.npz);
    ragged batch in one launch; edge lengths (multiple of the hop, shorter than one frame + reflect both sides)."""
    from tts_indic_server_f5_b200 import melspec as M
    g = np.load(os.path.join(golden_dir, "tiny.npz"))
    w0 = S.prompt_audio(0.6, 0)
    got = M.mel_spectrogram(w0.cuda())[0].cpu().numpy()
    np.testing.assert_allclose(got, g["prompt_mel"], rtol=0, atol=2e-4)      # log-mel range is ~[-11.5, 3]
    waves = [S.prompt_audio(5.0, 1)[0], S.prompt_audio(1.0, 2)[0], S.prompt_audio(0.3, 3)[0][:7000], S.prompt_audio(0.2, 4)[0][:768]]
    rows = M.mel_rows([w.cuda() for w in waves])
    for w, r in zip(waves, rows):
        want = O.mel_spectrogram(w[None])[0].t().numpy()
        assert r.shape == want.shape == (1 + w.numel() // 256, 100)
        err = np.abs(r.cpu().numpy() - want)
        # bins at the 1e-5 clamp amplify fp32 FFT round-off; everything audible agrees to 1e-4
        assert err.max() < 5e-3 and np.median(err) < 2e-5, (err.max(), np.median(err))
    lin = np.exp(rows[0].cpu().numpy()) - np.exp(O.mel_spectrogram(waves[0][None])[0].t().numpy())
    assert np.abs(lin).max() < 1e-4 * np.exp(rows[0].cpu().numpy()).max()


def test_prompt_cache_reuses_the_voice(tiny_models):
    cfg, vcfg, sd, vsd, model, voc = tiny_models
    syn = api.Synthesizer(model, voc)
    specs = S.workload("tiny3")
    a = syn.generate(specs, noise_seed=11)
    first = (syn.prompt_cache.hits, syn.prompt_cache.misses)
    b = syn.generate(specs, noise_seed=11)
    assert syn.prompt_cache.hits > first[0] and syn.prompt_cache.misses == first[1]
    for x, y in zip(a, b):
        assert np.array_equal(x, y)                # cached mel == recomputed mel, bit for bit


def test_request_scheduler_equals_per_request_calls(tiny_models, tmp_path):
    """Two concurrent requests (one multi-chunk) through RequestScheduler == `infer_process` per request, bit for bit:
    chunks of different requests share packs, the noise index and the cross-fade stay per request."""
    cfg, vcfg, sd, vsd, model, voc = tiny_models
    audio = S.prompt_audio(5.0, 7)                                            # 5 s prompt: max_chars = ref_bytes / 5 * 20
    ref_text = T.finish_ref_text(T.synthetic_indic_text(30, 1, "kannada"))
    texts = [T.synthetic_indic_text(60, 2, "kannada"), " ".join(T.synthetic_indic_text(40, 10 + k, "devanagari") + "." for k in range(4))]
    sched = api.RequestScheduler(api.Synthesizer(model, voc), max_rows=6144, max_utts=3, nfe_step=8)
    ids = [sched.submit((audio, 24000), ref_text, t, seed=100 + k) for k, t in enumerate(texts)]
    assert len(sched.pending[1].chunks) >= 2                                 # the second text does not fit one chunk
    out = sched.run()
    assert sched.pending == [] and len(sched.last_packs) >= 2                # the row budget forced several packs
    for k, (rid, t) in enumerate(zip(ids, texts)):
        wave, sr, mel = api.infer_process((audio, 24000), ref_text, t, model, voc, nfe_step=8, seed=100 + k)
        assert sr == out[rid][1] == 24000
        np.testing.assert_array_equal(out[rid][0], wave)
        np.testing.assert_array_equal(out[rid][2], mel)


def test_step_graph_is_reused_across_lengths(tiny_models):
    """The captured 32-step graph depends on the packed row count and the padded attention item count only: a batch of other
    utterance lengths that packs into the same rows REPLAYS it (the tile table, positions, noise, prompt are buffer contents),
    and the replay is bit-identical to a fresh engine that captures its own graph for that batch."""
    cfg, vcfg, sd, vsd, model, voc = tiny_models
    syn = api.Synthesizer(model, voc)

    def specs_for(gens):
        out = []
        for i, g_ in enumerate(gens):
            sp = S.workload("tiny3")[i % 3]
            sp.duration = sp.meta["ref_len"] + g_
            out.append(sp)
        return out

    a = specs_for([150, 200, 170])
    b = specs_for([155, 195, 170])            # other lengths, same total rows after 128-row padding, same padded item count
    syn.generate(a, noise_seed=3)
    eng = model.engine
    c0, r0 = eng.graph_captures, eng.graph_replays
    wb = syn.generate(b, noise_seed=3)
    assert (eng.graph_captures, eng.graph_replays) == (c0, r0 + 1), "a new length signature must not re-capture the step graph"
    fresh = api.Synthesizer(api.load_model(state_dict=sd), voc)
    for x, y in zip(wb, fresh.generate(b, noise_seed=3)):
        np.testing.assert_array_equal(x, y)
    syn.generate(a, noise_seed=3)             # and back
    assert eng.graph_captures == c0


def test_istft_perfect_reconstruction_property():
    """Size-independent property of the ISTFT kernels: analysing a signal with the matching STFT and feeding
    (log|X|, angle X) back reconstructs the signal (hann, hop = n_fft/4 satisfies COLA)."""
    g = torch.Generator().manual_seed(2)
    x = (torch.randn(1, 256 * 300, generator=g) * 0.1).cuda()
    win = torch.hann_window(1024, device="cuda")
    X = torch.stft(x, 1024, 256, 1024, win, center=True, return_complex=True)[0].t()       # [T, 513]
    T_ = X.shape[0]
    spec = torch.zeros(T_, 1152, device="cuda")
    spec[:, :513] = X.abs().clamp_min(1e-30).log()
    spec[:, 513:1026] = X.angle()
    frames = torch.zeros(T_, 1024, device="cuda")
    wav = torch.zeros(256 * (T_ - 1), device="cuda")
    seg = torch.tensor([[0, T_, 0, 0]], dtype=torch.int32, device="cuda")
    ops.istft(spec, win, frames, seg, wav.numel(), wav)
    torch.cuda.synchronize()
    assert snr(wav.cpu().numpy(), x[0, : wav.numel()].cpu().numpy()) > 90.0


def test_packing_invariance_and_graph_equals_eager_full_size(full_models):
    """At the full IndicF5 size: an utterance sampled inside a ragged packed batch equals the same utterance sampled
    alone (per-utterance semantics: attention / conv halo / GRN never cross utterances), and CUDA-graph replay equals
    eager launches bit for bit."""
    model, voc = full_models
    syn = api.Synthesizer(model, voc)
    specs = S.workload("small8")
    waves = syn.generate(specs, nfe_step=4, noise_seed=9)
    alone = syn.generate([specs[3]], nfe_step=4, noise_seed=9)[0]   # same Philox key => same noise whatever the packing
    s_pack = snr(waves[3], alone)
    print(f"packed vs alone (through the batched prompt STFT) SNR {s_pack:.1f} dB")
    assert alone.shape == waves[3].shape and s_pack > 90.0   # identical up to the batched-vs-single prompt STFT plan
    assert all(np.isfinite(w).all() and np.abs(w).max() > 0 for w in waves)
    # bit-exact at the sampler level when the inputs are bit-identical (every kernel is deterministic and row-local)
    mels = [O.mel_spectrogram(sp.audio).permute(0, 2, 1)[0] for sp in specs[:4]]
    toks = [T.convert_char_to_pinyin([sp.ref_text + sp.gen_text])[0] for sp in specs[:4]]
    y0 = [S.initial_noise(4096, i) for i in range(4)]
    durs = torch.tensor([sp.duration for sp in specs[:4]])
    cond = torch.nn.utils.rnn.pad_sequence(mels, batch_first=True)
    lens = torch.tensor([m.shape[0] for m in mels])
    outb, _ = model.sample(cond=cond, text=toks, duration=durs, lens=lens, steps=3, cfg_strength=2.0, sway_sampling_coef=-1.0, y0=y0)
    out2, _ = model.sample(cond=mels[2][None], text=[toks[2]], duration=int(durs[2]), steps=3, cfg_strength=2.0,
                           sway_sampling_coef=-1.0, y0=[y0[2]])
    assert torch.equal(outb[2, : int(durs[2])], out2[0])
    model.engine.use_graphs = False
    try:
        eager = syn.generate(specs, nfe_step=4, noise_seed=9)
    finally:
        model.engine.use_graphs = True
    again = syn.generate(specs, nfe_step=4, noise_seed=9)
    assert all(np.array_equal(a, b) for a, b in zip(eager, again))


def test_infer_process_and_manager_surface(tiny_models, tmp_path):
    """Boundary #2 / #1 surface: infer_process (chunking + cross-fade) and the TTSManager contract."""
    import wave as wavmod
    *_, model, voc = tiny_models
    audio = S.prompt_audio(1.0, 2)
    p = str(tmp_path / "ref.wav")
    with wavmod.open(p, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(24000)
        w.writeframes((audio[0].numpy() * 32767).astype("<i2").tobytes())
    ref_text = T.finish_ref_text(T.synthetic_indic_text(20, 1))
    gen = ". ".join(T.synthetic_indic_text(60, 10 + i) for i in range(12)) + "."
    wave, sr, mel = api.infer_process(p, ref_text, gen, model, voc, nfe_step=4)
    assert sr == 24000 and wave.ndim == 1 and mel.shape[0] == 100 and np.isfinite(wave).all()
    nchunks = len(T.chunk_text(gen, int(len(ref_text.encode()) / 1.0 * 24)))
    assert nchunks >= 2 and len(wave) == 256 * (mel.shape[1] - nchunks) - (nchunks - 1) * 3600
    mgr = api.TTSManager()
    assert not mgr.model
    with pytest.raises(ValueError, match="TTS model not loaded"):
        mgr.synthesize("x", p, ref_text)


# ------------------------------------------------------------------------------------------------ round 2: benchmark sizes
@pytest.mark.parametrize("n", [1384, 3069])
def test_full_size_forward_pair_at_benchmark_lengths(full_models, golden_dir, n):
    """One CFG velocity pair of the full IndicF5 DiT at a C2 length (1384) and the C3 long-form length (3069: 24 key tiles per
    query tile, the attention-heavy case) against the real reference's `DiT.forward` (dit.py:130-163) on the same inputs."""
    model, _ = full_models
    g = np.load(os.path.join(golden_dir, "full_fwd.npz"))
    x, cond, text = S.forward_inputs(n, W.INDICF5.vocab_size)
    u = UtteranceInput(cond=cond[0], text_ids=text[0], n=n, cond_len=n, y0=x[0])
    pc = model.engine.forward_flow([u], 0.37)[0].cpu().numpy()
    for k, name in enumerate(("cond", "null")):
        want = g[f"fwd{n}_{name}"]
        r, li = rel(pc[k], want), float(np.abs(pc[k] - want).max())
        print(f"forward n={n} {name}: rel-L2 {r:.2e} L-inf {li:.3f} (|ref| max {np.abs(want).max():.2f})")
        assert r < MEL_REL
    # the same utterance inside a ragged pack (three other lengths around it) is bit-identical to the utterance alone
    others = [UtteranceInput(cond=cond[0, :m], text_ids=text[0, :50], n=m, cond_len=m, y0=x[0, :m]) for m in (200, 731)]
    packed = model.engine.forward_flow([others[0], u, others[1]], 0.37)[1]
    assert torch.equal(packed.cpu(), torch.from_numpy(pc))


def test_c2_batch_utterances_vs_reference_golden(full_models, golden_dir):
    """The benchmark workload itself: all 64 utterances of C2 through `Synthesizer.generate` (one packed batch, 32 Euler steps,
    CFG pair, Vocos), utterances 0 (n = 1384) and 37 (n = 1277) compared with the real reference's `CFM.sample` + Vocos run
    one utterance at a time in fp32 on the same noise (cfm.py:162-176, utils_infer.py:459-476)."""
    model, voc = full_models
    g = np.load(os.path.join(golden_dir, "full_c2_utts.npz"))
    specs = S.workload("c2")
    waves, mels = api.Synthesizer(model, voc).generate(specs, return_mel=True, y0=S.reference_noise(specs))
    for k in g["indices"]:
        gm, gw = g[f"utt{k}_mel"], g[f"utt{k}_wave"]
        r, li, s = rel(mels[k].T, gm), float(np.abs(mels[k].T - gm).max()), snr(waves[k], gw)
        print(f"C2 utterance {k} (n={specs[k].duration}) in the 64-pack: mel rel-L2 {r:.2e} L-inf {li:.3f} wave SNR {s:.1f} dB")
        assert mels[k].T.shape == gm.shape and waves[k].shape == gw.shape
        assert r < MEL_REL and li < MEL_LINF and s > WAVE_SNR


def test_tts_manager_load_and_synthesize_vs_oracle(tiny_models, tmp_path):
    """Boundary #1 end to end (managers.py:62-85): `TTSManager().load()` then `.synthesize(text, ref_audio_path, ref_text)` on a
    WAV file with silent edges -> int16 array, against the oracle chain on the same file: pydub-port prompt conditioning
    (utils_infer.py:285-320) -> '. ' rule -> byte-budget chunking -> per-chunk `infer_one` (fp32) -> cross-fade -> int16."""
    import wave as wavmod
    from oracle import pydub_port as PP
    cfg, vcfg, sd, vsd, _, _ = tiny_models
    rate = 24000
    body = S.prompt_audio(1.2, 3)[0].numpy()
    pcm = np.concatenate([np.zeros(int(0.25 * rate)), body, np.zeros(int(0.4 * rate))])
    path = str(tmp_path / "ref.wav")
    with wavmod.open(path, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(rate)
        w.writeframes(np.clip(np.round(pcm * 32767), -32768, 32767).astype("<i2").tobytes())
    ref_text = T.synthetic_indic_text(24, 1)
    gen = ". ".join(T.synthetic_indic_text(60, 20 + i) for i in range(14)) + "."
    mgr = api.TTSManager(state_dict=sd, vocoder_state_dict=vsd)
    mgr.load()
    assert mgr.model
    mgr.load()                                                               # idempotent
    mgr.model.noise_fn = lambda i, n: S.initial_noise(n, i)
    got = mgr.synthesize(gen, ref_audio_path=path, ref_text=ref_text)
    assert got.dtype == np.int16 and got.ndim == 1
    # ---- oracle chain
    seg = PP.clip_reference(PP.Seg.from_wav(path), True, lambda *_: None)
    audio = torch.from_numpy(np.frombuffer(seg._data, dtype="<i2").astype(np.float32) / 32768.0)[None]
    assert audio.shape[-1] < len(pcm) - int(0.5 * rate)                      # the silent edges were trimmed
    rt = T.finish_ref_text(ref_text)
    max_chars = int(len(rt.encode("utf-8")) / (audio.shape[-1] / rate) * (25 - audio.shape[-1] / rate))
    chunks = T.chunk_text(gen, max_chars=max_chars)
    assert len(chunks) >= 2
    vocab = {t: i for i, t in enumerate(T.synthetic_indic_vocab())}
    waves = []
    rt2 = rt + " " if len(rt[-1].encode("utf-8")) == 1 else rt              # utils_infer.py:438-439 (inside infer_batch_process)
    with torch.inference_mode():
        for i, ch in enumerate(chunks):
            a, _ = O.rms_normalise(audio)
            ids = O.list_str_to_idx(T.convert_char_to_pinyin([rt2 + ch]), vocab)
            dur = O.estimate_duration(a.shape[-1] // 256, rt2, ch)
            wv, _ = O.infer_one(sd, cfg, vsd, vcfg, audio, ids, dur, y0=S.initial_noise(4096, i))
            waves.append(wv.numpy())
    want = np.clip(O.cross_fade(waves) * 32768.0, -32768, 32767).astype(np.int16)
    assert got.shape == want.shape
    s = snr(got.astype(np.float64), want.astype(np.float64))
    print(f"TTSManager.synthesize vs oracle chain: {len(chunks)} chunks, {len(got)} samples, SNR {s:.1f} dB")
    assert s > 38.0                                                          # bf16-operand path + int16 quantisation
    # default path: a fresh device draw per call (the reference is stochastic unless seeded, cfm.py:181-186)
    mgr.model.noise_fn = None
    a1, a2 = mgr.synthesize(gen, path, ref_text), mgr.synthesize(gen, path, ref_text)
    assert a1.shape == a2.shape == got.shape and not np.array_equal(a1, a2)
    torch.manual_seed(5)
    b1 = mgr.synthesize(gen, path, ref_text)
    torch.manual_seed(5)
    assert np.array_equal(b1, mgr.synthesize(gen, path, ref_text))           # ... and repeatable under torch.manual_seed


def test_speech_route_on_the_engine_equals_the_manager_call(tiny_models, tmp_path):
    """SURVEY §8f row 4 end to end: `server.create_app` around a real `TTSManager` (tiny weights).  `POST /v1/audio/speech` goes
    prompt cache -> ContinuousScheduler (worker thread owns the GPU) -> packed engine batch -> cross-fade -> int16 -> 16-bit PCM
    WAV, and must carry exactly the samples `TTSManager.synthesize` (boundary #1, managers.py:82-85) returns for the same text
    under the same torch seed; concurrent requests come back correct and batched."""
    import io
    import threading
    import wave as wavmod
    from fastapi.testclient import TestClient
    from tts_indic_server_f5_b200 import server
    cfg, vcfg, sd, vsd, _, _ = tiny_models
    rate = 24000
    pcm = np.concatenate([np.zeros(int(0.2 * rate)), S.prompt_audio(1.2, 3)[0].numpy(), np.zeros(int(0.3 * rate))])
    path = str(tmp_path / "kan.wav")
    with wavmod.open(path, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(rate)
        w.writeframes(np.clip(np.round(pcm * 32767), -32768, 32767).astype("<i2").tobytes())
    ref_text = T.synthetic_indic_text(24, 1)
    texts = [". ".join(T.synthetic_indic_text(60, 40 + 7 * k + i) for i in range(6 + 4 * k)) + "." for k in range(3)]
    mgr = api.TTSManager(state_dict=sd, vocoder_state_dict=vsd)
    app = server.create_app(mgr, {"KAN_F (Happy)": server.Voice(path, ref_text)}, max_wait_ms=50.0)

    def samples(resp):
        assert resp.status_code == 200 and resp.headers["content-type"] == "audio/wav"
        with wavmod.open(io.BytesIO(resp.content)) as w:
            assert (w.getframerate(), w.getnchannels(), w.getsampwidth()) == (24000, 1, 2)
            return np.frombuffer(w.readframes(w.getnframes()), dtype="<i2")

    def as_wav_pcm(int16_audio):                                              # tts_utils.py:60-64: / 32768, then PCM_16
        return np.rint(int16_audio.astype(np.float32) / 32768.0 * 32767.0).astype(np.int16)

    with TestClient(app) as client:                                           # lifespan: mgr.load()
        assert mgr.model and client.get("/v1/health").json()["status"] == "healthy"
        want = []
        for k, t in enumerate(texts):
            torch.manual_seed(100 + k)
            want.append(mgr.synthesize(t, ref_audio_path=path, ref_text=ref_text))
        torch.manual_seed(100)
        got0 = samples(client.post("/v1/audio/speech", json={"text": texts[0]}))
        assert got0.shape == want[0].shape and np.array_equal(got0, as_wav_pcm(want[0]))
        res = {}

        def call(k):
            res[k] = client.post("/v1/audio/speech", json={"text": texts[k]})

        ths = [threading.Thread(target=call, args=(k,)) for k in range(3)]
        [t.start() for t in ths]
        [t.join() for t in ths]
        for k in range(3):                                                    # fresh noise per request: same length, sane audio
            g = samples(res[k])
            assert g.shape == want[k].shape and np.abs(g.astype(np.int32)).max() > 0
        batches = client.get("/v1/health").json()["batches"]
        print(f"speech route: batches {batches}")
        assert sum(batches) == 4                                              # (how they group is asserted on the CPU, test_host_logic.py)
        assert client.post("/v1/audio/speech", json={"text": " "}).status_code == 400


def test_checkpoint_file_loads_like_the_state_dict(tiny_models, tmp_path):
    """`load_model(ckpt_path=...)` on an EMA .safetensors / .pt file (utils_infer.py:195-213 key rules) builds the same engine
    as the in-memory state dict: bit-identical samples."""
    from safetensors.torch import save_file
    cfg, _, sd, _, model, _ = tiny_models
    ema = {"ema_model." + k: v.contiguous() for k, v in sd.items()}
    save_file({**ema, "initted": torch.tensor([1]), "step": torch.tensor([7])}, str(tmp_path / "m.safetensors"))
    torch.save({"ema_model_state_dict": {**ema, "initted": torch.tensor(True), "step": torch.tensor(7)}}, str(tmp_path / "m.pt"))
    mel = O.mel_spectrogram(S.prompt_audio(0.5, 1)).permute(0, 2, 1)
    texts = [T.convert_char_to_pinyin([T.synthetic_indic_text(30, 4)])[0]]
    y0 = [S.initial_noise(100, 0)]
    want, _ = model.sample(cond=mel, text=texts, duration=90, steps=4, cfg_strength=2.0, sway_sampling_coef=-1.0, y0=y0)
    for f in ("m.safetensors", "m.pt"):
        m2 = api.load_model(ckpt_path=str(tmp_path / f))
        assert m2.cfg == cfg
        got, _ = m2.sample(cond=mel, text=texts, duration=90, steps=4, cfg_strength=2.0, sway_sampling_coef=-1.0, y0=y0)
        assert torch.equal(got, want)


def test_engine_rejects_what_it_cannot_hold(tiny_models):
    """Loud failures instead of out-of-range device reads: too many Euler steps for the hoisted tables, a token id beyond the
    embedding (nn.Embedding raises IndexError at dit.py:56), a re-staged workspace."""
    cfg, _, sd, _, model, voc = tiny_models
    mel = O.mel_spectrogram(S.prompt_audio(0.5, 1)).permute(0, 2, 1)
    texts = [T.convert_char_to_pinyin([T.synthetic_indic_text(30, 4)])[0]]
    with pytest.raises(ValueError, match="128 Euler steps"):
        model.sample(cond=mel, text=texts, duration=90, steps=129, cfg_strength=2.0)
    bad = torch.full((1, 10), cfg.vocab_size + 5, dtype=torch.long)
    with pytest.raises(IndexError):
        model.sample(cond=mel, text=bad, duration=90, steps=2, cfg_strength=2.0)
    syn = api.Synthesizer(model, voc)
    specs = S.workload("tiny3")
    st1 = syn.stage(specs, noise_seed=1)
    syn.stage(specs, noise_seed=2)                                            # same packed size: takes over the workspace
    with pytest.raises(RuntimeError, match="overwritten"):
        syn.run(st1)


def test_programmatic_dependent_launch_is_bit_invisible(tiny_models):
    """PDL (f5_set_pdl) only moves each kernel's set-up ahead of its predecessor's tail: eager launches and a freshly captured
    step graph with it ON equal eager launches with it OFF bit for bit, over the whole path (sampler + vocoder)."""
    from tts_indic_server_f5_b200 import _lib
    cfg, vcfg, sd, vsd, model, voc = tiny_models
    specs = S.workload("tiny3")
    m2 = api.load_model(state_dict=sd)                       # own engine: own workspaces and graphs
    m2.engine.pdl_auto = False                               # this test drives the switch itself
    syn = api.Synthesizer(m2, voc)
    old = _lib.lib.f5_set_pdl(0)
    try:
        m2.engine.use_graphs = False
        off = syn.generate(specs, nfe_step=6, noise_seed=21)
        assert _lib.lib.f5_set_pdl(1) == 0
        on_eager = syn.generate(specs, nfe_step=6, noise_seed=21)
        m2.engine.use_graphs = True
        on_graph = syn.generate(specs, nfe_step=6, noise_seed=21)      # captured now, with programmatic edges
        on_replay = syn.generate(specs, nfe_step=6, noise_seed=21)
    finally:
        _lib.lib.f5_set_pdl(old)
    for a, b, c, d in zip(off, on_eager, on_graph, on_replay):
        assert np.array_equal(a, b) and np.array_equal(a, c) and np.array_equal(a, d)


# ------------------------------------------------------------------------------------------------ round 2: fp32 precision mode
FP32_MEL_REL, FP32_MEL_LINF, FP32_WAVE_SNR = 5e-5, 5e-4, 85.0      # measured 6.4e-6 .. 8.8e-6, < 1e-4, 98 .. 103 dB


@pytest.fixture(scope="module")
def tiny_models_fp32():
    cfg, vcfg = W.tiny_dit_config(), W.tiny_vocos_config()
    sd, vsd = W.make_dit_state_dict(cfg, seed=1), W.make_vocos_state_dict(vcfg, seed=1)
    return api.load_model(state_dict=sd, precision="fp32"), api.load_vocoder(state_dict=vsd, precision="fp32")


def test_fp32_mode_tiny_vs_reference_golden(tiny_models_fp32, golden_dir):
    """precision="fp32" (split-operand GEMMs + fp32 attention) against the real reference's fp32 output — its own tolerance,
    stated separately from the bf16 numbers: mel rel-L2 <= 5e-5, L-inf <= 5e-4, waveform SNR >= 85 dB."""
    model, voc = tiny_models_fp32
    g = np.load(os.path.join(golden_dir, "tiny.npz"))
    u = UtteranceInput(cond=torch.from_numpy(g["fwd_condin"]), text_ids=torch.from_numpy(g["fwd_text"]), n=96, cond_len=96,
                       y0=torch.from_numpy(g["fwd_x"]))
    pc = model.engine.forward_flow([u], 0.37)[0].cpu().numpy()
    print(f"fp32 forward pair: rel-L2 {rel(pc[0], g['fwd_cond']):.2e} / {rel(pc[1], g['fwd_null']):.2e}")
    assert rel(pc[0], g["fwd_cond"]) < 1e-4 and rel(pc[1], g["fwd_null"]) < 1e-4
    for wl in ("tiny", "tiny3"):
        specs = S.workload(wl)
        waves, mels = api.Synthesizer(model, voc).generate(specs, return_mel=True, y0=ref_noise(specs))
        for i, (spec, wv, ml) in enumerate(zip(specs, waves, mels)):
            gm = g[f"{wl}_{i}_mel"][spec.meta["ref_len"]:]
            r, li, s = rel(ml.T, gm), float(np.abs(ml.T - gm).max()), snr(wv, g[f"{wl}_{i}_wave"])
            print(f"fp32 {wl}[{i}] mel rel-L2 {r:.2e} L-inf {li:.4f} wave SNR {s:.1f} dB")
            assert r < FP32_MEL_REL and li < FP32_MEL_LINF and s > FP32_WAVE_SNR


def test_fp32_mode_full_size_c1_vs_reference_golden(golden_dir):
    model = api.load_model(state_dict=W.make_dit_state_dict(W.INDICF5, seed=0), precision="fp32")
    voc = api.load_vocoder(state_dict=W.make_vocos_state_dict(W.VOCOS_24K, seed=0), precision="fp32")
    g = np.load(os.path.join(golden_dir, "full_c1.npz"))
    spec = S.workload("c1")
    waves, mels = api.Synthesizer(model, voc).generate(spec, return_mel=True, y0=ref_noise(spec))
    gm = g["mel"][spec[0].meta["ref_len"]:]
    r, li, s = rel(mels[0].T, gm), float(np.abs(mels[0].T - gm).max()), snr(waves[0], g["wave"])
    print(f"fp32 full C1: mel rel-L2 {r:.2e} L-inf {li:.4f} wave SNR {s:.1f} dB")
    assert r < FP32_MEL_REL and li < FP32_MEL_LINF and s > FP32_WAVE_SNR
    del model, voc
    torch.cuda.empty_cache()


def test_trajectory_and_fp32_forward_at_c2_length(tiny_models, golden_dir):
    """(1) `return_trajectory=True` reproduces torchdiffeq's stacked states (cfm.py:200: steps + 1 of them, the first is y0, the
    last is the state BEFORE the prompt re-insert) against the oracle's odeint; (2) the fp32 mode at a benchmark length: full
    IndicF5 forward pair at n = 1384 against the real reference, its own tolerance."""
    cfg, _, sd, _, model, _ = tiny_models
    vocab = {t: i for i, t in enumerate(T.synthetic_indic_vocab())}
    mel = O.mel_spectrogram(S.prompt_audio(0.5, 1)).permute(0, 2, 1)
    texts = [T.convert_char_to_pinyin([T.synthetic_indic_text(30, 4)])[0]]
    y0 = [S.initial_noise(100, 0)]
    out, traj = model.sample(cond=mel, text=texts, duration=90, steps=6, cfg_strength=2.0, sway_sampling_coef=-1.0, y0=y0,
                             return_trajectory=True)
    with torch.inference_mode():
        ref_out, ref_traj = O.cfm_sample(sd, cfg, mel, O.list_str_to_idx(texts, vocab), 90, y0=y0[0], steps=6, return_trajectory=True)
    assert traj.shape == (7, 1, 90, 100) and tuple(ref_traj.shape) == (7, 1, 90, 100)
    assert torch.equal(traj[0, 0].cpu(), y0[0][:90])                            # state 0 is the injected noise, untouched
    for k in range(1, 7):
        assert rel(traj[k].cpu().numpy(), ref_traj[k].numpy()) < MEL_REL, k
    assert rel(out.cpu().numpy(), ref_out.numpy()) < MEL_REL
    plain, last = model.sample(cond=mel, text=texts, duration=90, steps=6, cfg_strength=2.0, sway_sampling_coef=-1.0, y0=y0)
    assert torch.equal(plain, out) and last.shape == (1, 1, 90, 100)            # graph replay == the eager trajectory run
    m32 = api.load_model(state_dict=W.make_dit_state_dict(W.INDICF5, seed=0), precision="fp32")
    g = np.load(os.path.join(golden_dir, "full_fwd.npz"))
    x, cond, text = S.forward_inputs(1384, W.INDICF5.vocab_size)
    u = UtteranceInput(cond=cond[0], text_ids=text[0], n=1384, cond_len=1384, y0=x[0])
    pc = m32.engine.forward_flow([u], 0.37)[0].cpu().numpy()
    r0, r1 = rel(pc[0], g["fwd1384_cond"]), rel(pc[1], g["fwd1384_null"])
    print(f"fp32 forward n=1384: rel-L2 {r0:.2e} / {r1:.2e}")
    assert r0 < FP32_MEL_REL and r1 < FP32_MEL_REL
    del m32
    torch.cuda.empty_cache()


def test_continuous_scheduler_on_the_engine(tiny_models):
    """The serving loop with the real engine: requests submitted from several threads come back through futures and equal
    `infer_process(seed=...)` per request bit for bit (the worker thread packs whatever is waiting into shared batches)."""
    import threading
    *_, model, voc = tiny_models
    audio = S.prompt_audio(2.0, 5)
    ref_text = T.finish_ref_text(T.synthetic_indic_text(30, 1, "kannada"))
    texts = [T.synthetic_indic_text(40 + 7 * k, 50 + k, "kannada" if k % 2 else "devanagari") for k in range(5)]
    cs = api.ContinuousScheduler(api.Synthesizer(model, voc), max_queue=16, max_batch_requests=8, max_wait_ms=50.0, nfe_step=6)
    futs = {}

    def client(k):
        futs[k] = cs.submit((audio, 24000), ref_text, texts[k], seed=300 + k)

    ths = [threading.Thread(target=client, args=(k,)) for k in range(5)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    got = {k: f.result(timeout=120) for k, f in futs.items()}
    cs.close()
    assert sum(cs.batches) == 5 and len(cs.batches) <= 3
    for k in range(5):
        wave, sr, mel = api.infer_process((audio, 24000), ref_text, texts[k], model, voc, nfe_step=6, seed=300 + k)
        np.testing.assert_array_equal(got[k][0], wave)
        np.testing.assert_array_equal(got[k][2], mel)
