"""ORACLE — test infrastructure only.  Mint the golden vectors under `tests/golden/` from the REAL reference.

Run in the build container (needs `/root/reference`):   python -m oracle.make_golden [--full]

Each fixture is produced by the reference's own `CFM.sample` / `DiT.forward` / `get_vocos_mel_spectrogram`
(imported in place through `oracle/ref_shims.py`) on the deterministic inputs of
`tts_indic_server_f5_b200/synthetic.py` and the seeded weights of `.../weights.py`; the vocoder leg uses the
Vocos restatement (third-party, parity unpinned).  The initial noise `y0` replaces the reference's
`torch.randn` draw (cfm.py:181-186) by patching `torch.randn` for the duration of the call.
Fixtures hold OUTPUTS (+ the few inputs that are cheap to store); inputs are regenerated from seeds.
"""
from __future__ import annotations

import argparse
import contextlib
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import f5_oracle as O  # noqa: E402
from oracle import ref_shims as R  # noqa: E402
from tts_indic_server_f5_b200 import synthetic as S  # noqa: E402
from tts_indic_server_f5_b200 import text as T  # noqa: E402
from tts_indic_server_f5_b200 import weights as W  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


@contextlib.contextmanager
def inject_noise(y0: torch.Tensor):
    real = torch.randn

    def fake(*size, **kw):
        n = size[0] if not isinstance(size[0], (tuple, list)) else size[0][0]
        return y0[: int(n)].to(kw.get("dtype") or torch.float32).clone()

    torch.randn = fake
    try:
        yield
    finally:
        torch.randn = real


def vocab_map():
    return {t: i for i, t in enumerate(T.synthetic_indic_vocab())}


def run_reference_utterance(cfm, vocoder, spec: S.UtteranceSpec, steps: int, dtype=torch.float32):
    """The tensor part of infer_batch_process (utils_infer.py:423-482) through the reference's own objects."""
    audio = spec.audio
    rms = torch.sqrt(torch.mean(torch.square(audio)))
    if rms < 0.1:
        audio = audio * 0.1 / rms
    ref_len = audio.shape[-1] // 256
    ref_text = spec.ref_text
    if len(ref_text[-1].encode("utf-8")) == 1:
        ref_text = ref_text + " "
    final_text_list = R.load_reference().model_utils.convert_char_to_pinyin([ref_text + spec.gen_text])
    assert final_text_list == T.convert_char_to_pinyin([ref_text + spec.gen_text])
    duration = spec.duration if spec.duration is not None else O.estimate_duration(ref_len, ref_text, spec.gen_text)
    y0 = S.initial_noise(4096, spec.noise_index)
    with torch.inference_mode(), inject_noise(y0):
        out, _ = cfm.sample(cond=audio, text=final_text_list, duration=duration, steps=steps,
                            cfg_strength=2.0, sway_sampling_coef=-1.0)
        out = out.to(torch.float32)
        gen = out[:, ref_len:, :].permute(0, 2, 1)
        wave = vocoder.decode(gen)
        if rms < 0.1:
            wave = wave * rms / 0.1
    return out[0].numpy(), wave.squeeze(0).numpy()


def make_tiny():
    cfg, vcfg = W.tiny_dit_config(), W.tiny_vocos_config()
    sd, vsd = W.make_dit_state_dict(cfg, seed=1), W.make_vocos_state_dict(vcfg, seed=1)
    cfm = R.build_reference_cfm(sd, cfg, vocab_map())
    voc = R.build_reference_vocos(vsd, vcfg)
    out = {}
    for wl in ("tiny", "tiny3"):
        for i, spec in enumerate(S.workload(wl)):
            mel, wave = run_reference_utterance(cfm, voc, spec, steps=32)
            out[f"{wl}_{i}_mel"] = mel
            out[f"{wl}_{i}_wave"] = wave
    # single DiT forward (both CFG branches) on fixed inputs
    g = torch.Generator("cpu").manual_seed(5)
    n = 96
    x = torch.randn(1, n, 100, generator=g)
    cond = torch.randn(1, n, 100, generator=g)
    cond[:, 40:] = 0
    text = torch.randint(0, cfg.vocab_size, (1, 50), generator=g)
    t = torch.tensor(0.37)
    with torch.inference_mode():
        out["fwd_cond"] = cfm.transformer(x=x, cond=cond, text=text, time=t, drop_audio_cond=False, drop_text=False)[0].numpy()
        out["fwd_null"] = cfm.transformer(x=x, cond=cond, text=text, time=t, drop_audio_cond=True, drop_text=True)[0].numpy()
    out["fwd_x"], out["fwd_condin"], out["fwd_text"] = x[0].numpy(), cond[0].numpy(), text[0].numpy()
    # prompt mel of the reference's own extractor
    out["prompt_mel"] = R.load_reference().modules.get_vocos_mel_spectrogram(S.prompt_audio(0.6, 0))[0].numpy()
    np.savez_compressed(os.path.join(GOLDEN, "tiny.npz"), **out)
    print("tiny.npz:", {k: v.shape for k, v in out.items()})


def make_full(steps=32):
    cfg, vcfg = W.INDICF5, W.VOCOS_24K
    sd, vsd = W.make_dit_state_dict(cfg, seed=0), W.make_vocos_state_dict(vcfg, seed=0)
    cfm = R.build_reference_cfm(sd, cfg, vocab_map())
    voc = R.build_reference_vocos(vsd, vcfg)
    spec = S.workload("c1")[0]
    t0 = time.time()
    mel, wave = run_reference_utterance(cfm, voc, spec, steps=steps)
    dt = time.time() - t0
    np.savez_compressed(os.path.join(GOLDEN, "full_c1.npz"), mel=mel, wave=wave,
                        cpu_seconds=np.float64(dt), threads=np.int64(torch.get_num_threads()))
    print(f"full_c1.npz: mel {mel.shape} wave {wave.shape} in {dt:.1f}s on {torch.get_num_threads()} threads")


def make_full_sizes(steps=32, c2_indices=(0, 37)):
    """Benchmark-size fixtures (round 2): full IndicF5 forward pairs at a C2 length (n = 1384) and the C3 length (n = 3069),
    and complete NFE-32 utterances of the C2 batch, all through the real reference's own modules in fp32."""
    cfg, vcfg = W.INDICF5, W.VOCOS_24K
    sd, vsd = W.make_dit_state_dict(cfg, seed=0), W.make_vocos_state_dict(vcfg, seed=0)
    cfm = R.build_reference_cfm(sd, cfg, vocab_map())
    voc = R.build_reference_vocos(vsd, vcfg)
    out = {}
    for n in (1384, 3069):
        x, cond, text = S.forward_inputs(n, cfg.vocab_size)
        t0 = time.time()
        with torch.inference_mode():
            out[f"fwd{n}_cond"] = cfm.transformer(x=x, cond=cond, text=text, time=torch.tensor(0.37), drop_audio_cond=False,
                                                  drop_text=False)[0].numpy()
            out[f"fwd{n}_null"] = cfm.transformer(x=x, cond=cond, text=text, time=torch.tensor(0.37), drop_audio_cond=True,
                                                  drop_text=True)[0].numpy()
        print(f"forward pair n={n}: {time.time() - t0:.1f}s", flush=True)
    np.savez_compressed(os.path.join(GOLDEN, "full_fwd.npz"), **out)
    specs = S.workload("c2")
    out = {"indices": np.asarray(c2_indices, dtype=np.int64)}
    for k in c2_indices:
        t0 = time.time()
        mel, wave = run_reference_utterance(cfm, voc, specs[k], steps=steps)
        out[f"utt{k}_mel"], out[f"utt{k}_wave"] = mel[specs[k].meta["ref_len"]:], wave
        print(f"c2 utterance {k}: n={specs[k].duration} in {time.time() - t0:.1f}s", flush=True)
    np.savez_compressed(os.path.join(GOLDEN, "full_c2_utts.npz"), **out)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true", help="also mint the full-size C1 fixture (~2 min CPU)")
    ap.add_argument("--sizes", action="store_true", help="only mint the benchmark-size fixtures (C2 / C3 lengths, ~6 min CPU)")
    args = ap.parse_args()
    os.makedirs(GOLDEN, exist_ok=True)
    if args.sizes:
        make_full_sizes()
        sys.exit(0)
    make_tiny()
    if args.full:
        make_full()
