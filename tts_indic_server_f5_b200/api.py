"""Drop-in boundary of the hot path: the model-call API surface behind `POST /v1/audio/speech`.

Mirrors, name for name, what the Dhwani server and the IndicF5 wrapper call:
  boundary #1  `TTSManager.load()` / `.synthesize(text, ref_audio_path, ref_text)`      src/server/core/managers.py:62-85
  boundary #2  `load_model`, `load_vocoder`, `preprocess_ref_audio_text`, `infer_process`,
               `infer_batch_process`                                                     f5_tts/infer/utils_infer.py:92-524
               `CFM.sample(...)` and `vocoder.decode(mel)`                               f5_tts/model/cfm.py:81-210
Same argument meaning, defaults and error behaviour; internals are the CUDA engines (engine.py, vocos.py).  Nothing
here computes on the CPU except what the reference also does on the host (tokenisation, RMS of the prompt, the
duration rule, cross-fade) and the initial noise draw (CPU generator, so results are reproducible across devices).
`generate()` is the batched entry point the reference lacks: independent utterances are packed into one pass.
"""
from __future__ import annotations

import logging
import os
import threading
from dataclasses import dataclass

import numpy as np
import torch

from . import text as T
from .engine import F5Engine, UtteranceInput
from .melspec import PromptCache, mel_rows, mel_spectrogram
from .prompt_audio import remove_silence_for_generated_wav  # noqa: F401
from .scheduler import ContinuousScheduler, RequestScheduler, cross_fade  # noqa: F401
from .synthetic import UtteranceSpec
from .vocos import VocosEngine
from .weights import (INDICF5, VOCOS_24K, DiTConfig, VocosConfig, infer_dit_config, make_dit_state_dict,
                      make_vocos_state_dict, strip_checkpoint)

logger = logging.getLogger("tts_indic_server_f5_b200")

# constants of f5_tts/infer/utils_infer.py:40-53
target_sample_rate = 24000
n_mel_channels = 100
hop_length = 256
win_length = 1024
n_fft = 1024
mel_spec_type = "vocos"
target_rms = 0.1
cross_fade_duration = 0.15
ode_method = "euler"
nfe_step = 32
cfg_strength = 2.0
sway_sampling_coef = -1.0
speed = 1.0
fix_duration = None


def _default_device() -> str:
    return "cuda" if torch.cuda.is_available() else "cpu"


def fresh_noise_seed() -> int:
    """A 64-bit key for the device noise draw, taken from torch's global CPU generator: like the reference's unseeded
    `torch.randn` (cfm.py:186) every call is a new draw, and like it `torch.manual_seed(s)` beforehand makes it repeatable."""
    return int(torch.empty((), dtype=torch.int64).random_())


def utterance_seed(base: int, index: int) -> int:
    """splitmix64 of (base, index): independent Philox keys for the utterances of one request batch."""
    z = (base + 0x9E3779B97F4A7C15 * (index + 1)) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return z ^ (z >> 31)


class CFM:
    """Engine-backed stand-in for the reference `CFM` (inference surface only: `.sample`, `.device`, `.eval`, `.to`)."""

    def __init__(self, state_dict: dict, cfg: DiTConfig, vocab_char_map: dict | None = None, device: str = "cuda",
                 precision: str = "bf16"):
        if not str(device).startswith("cuda"):
            raise RuntimeError("the B200 CFM engine runs on CUDA (sm_100a) only; there is no CPU fallback")
        self.cfg = cfg
        self.vocab_char_map = vocab_char_map
        self.num_channels = cfg.mel_dim
        self.engine = F5Engine(state_dict, cfg, device, precision=precision)
        self._device = self.engine.device

    @property
    def device(self):
        return self._device

    def eval(self):
        return self

    def to(self, device):
        if torch.device(device).type != "cuda":
            raise RuntimeError("the B200 CFM engine cannot be moved off CUDA")
        return self

    def __bool__(self):
        return True

    # -- CFM.sample prologue (cfm.py:100-149) on the host --------------------------------------------------------
    def _prepare(self, cond, text, duration, lens, seed, max_duration, edit_mask, y0) -> list[UtteranceInput]:
        if cond.ndim == 2:                                           # raw wave -> mel (cfm.py:103-106)
            cond = mel_spectrogram(cond.to(self._device).float()).permute(0, 2, 1)
            assert cond.shape[-1] == self.num_channels
        cond = cond.float()
        batch, cond_seq_len = cond.shape[:2]
        lens_t = torch.full((batch,), cond_seq_len, dtype=torch.long) if lens is None else torch.as_tensor(lens).long().cpu()
        if isinstance(text, list):
            if self.vocab_char_map is None:
                raise NotImplementedError("byte tokenizer (list_str_to_tensor) is not on the IndicF5 path")
            text = T.list_str_to_idx(text, self.vocab_char_map)
        text = text.cpu().long()
        assert text.shape[0] == batch
        text_lens = (text != -1).sum(dim=-1)
        lens_t = torch.maximum(text_lens, lens_t)                    # cfm.py:123-125
        if isinstance(duration, int):
            duration = torch.full((batch,), duration, dtype=torch.long)
        duration = torch.maximum(lens_t + 1, torch.as_tensor(duration).long().cpu()).clamp(max=max_duration)  # :136-137
        utts = []
        base = fresh_noise_seed() if (y0 is None and seed is None) else 0
        for i in range(batch):
            n = int(duration[i])
            noise = None
            if y0 is not None:
                noise = y0[i][:n].float().cpu()
            elif seed is not None:                                   # cfm.py:182-186: re-seed, then randn, per item — on the CPU
                torch.manual_seed(seed)                              # generator, so a seeded call reproduces the reference's CPU draw
                noise = torch.randn(n, self.num_channels, dtype=torch.float32)
            ids = text[i][text[i] != -1]
            em = None if edit_mask is None else edit_mask[i]
            utts.append(UtteranceInput(cond=cond[i], text_ids=ids, n=n, cond_len=int(lens_t[i]), y0=noise, edit_mask=em,
                                       noise_seed=utterance_seed(base, i)))       # unseeded: drawn on the device (f5_randn_rows)
        return utts

    @torch.inference_mode()
    def sample(self, cond, text, duration, *, lens=None, steps=32, cfg_strength=1.0, sway_sampling_coef=None, seed=None,
               max_duration=4096, vocoder=None, no_ref_audio=False, duplicate_test=False, t_inter=0.1, edit_mask=None,
               y0=None, return_trajectory: bool = False):
        """Signature and return convention of the reference `CFM.sample` (cfm.py:82-99, :210): returns
        `(out [b, n, 100], trajectory)`.  Each utterance is sampled with batch-1 semantics (what the server computes);
        rows past an utterance's own duration are zero.  By default `trajectory` holds only the final state ([1, b, n, 100]):
        the 33-state history the reference keeps (cfm.py:200) is never consumed on the served path; `return_trajectory=True`
        returns all `steps + 1` states ([steps + 1, b, n, 100], eager launches).  `y0` (list of [n_i, 100] or [b, n, 100])
        overrides the noise draw."""
        if duplicate_test:
            raise NotImplementedError("duplicate_test is a training-time probe, not part of the served path")
        utts = self._prepare(cond, text, duration, lens, seed, max_duration, edit_mask, y0)
        states: list = []

        def keep(k, ws_):                                            # odeint's stacked states (cfm.py:200), before the prompt re-insert
            states.append(ws_.x[:, : self.num_channels].clone())

        ws, layout = self.engine.sample_packed(utts, steps=steps, cfg_strength=cfg_strength,
                                               sway_sampling_coef=sway_sampling_coef, on_state=keep if return_trajectory else None)
        n_max = max(layout.lengths)
        out = torch.zeros(len(utts), n_max, self.num_channels, device=self._device, dtype=torch.float32)
        for i, (s, n) in enumerate(zip(layout.starts, layout.lengths)):
            out[i, :n] = ws.x[s:s + n, : self.num_channels]
            if no_ref_audio:                                         # cfm.py:157-158: prompt region re-inserted as zeros
                out[i, : utts[i].cond_len] = torch.where(
                    ws.cond_flag[s:s + utts[i].cond_len, None].bool(), torch.zeros_like(out[i, : utts[i].cond_len]),
                    out[i, : utts[i].cond_len])
        trajectory = out.unsqueeze(0)
        if return_trajectory:
            trajectory = torch.zeros(len(states), len(utts), n_max, self.num_channels, device=self._device, dtype=torch.float32)
            for k, xs in enumerate(states):
                for i, (s0, n) in enumerate(zip(layout.starts, layout.lengths)):
                    trajectory[k, i, :n] = xs[s0:s0 + n]
        if vocoder is not None:
            out = vocoder(out.permute(0, 2, 1))
        return out, trajectory


class Vocos:
    """Engine-backed stand-in for `vocos.Vocos` (only `.decode` is used on the path, utils_infer.py:472)."""

    def __init__(self, state_dict: dict, cfg: VocosConfig = VOCOS_24K, device: str = "cuda", precision: str = "bf16"):
        self.engine = VocosEngine(state_dict, cfg, device, precision=precision)
        self.cfg = cfg

    def decode(self, mel: torch.Tensor) -> torch.Tensor:
        return self.engine.decode(mel)

    def eval(self):
        return self

    def to(self, device):
        return self


def load_vocoder(vocoder_name="vocos", is_local=False, local_path="", device=None, hf_cache_dir=None, state_dict=None,
                 seed: int = 0, precision: str = "bf16") -> Vocos:
    """utils_infer.py:92-130.  `is_local`: reads `<local_path>/pytorch_model.bin` (vocos-mel-24khz layout).  Without
    local weights (no network here) a seeded random-init Vocos of the named architecture is built."""
    device = device or _default_device()
    if vocoder_name != "vocos":
        raise NotImplementedError("only the vocos branch is on the served path (bigvgan is not vendored by the reference)")
    if state_dict is None and is_local:
        print(f"Load vocos from local path {local_path}")
        state_dict = torch.load(f"{local_path}/pytorch_model.bin", map_location="cpu", weights_only=True)
    if state_dict is None:
        state_dict = make_vocos_state_dict(VOCOS_24K, seed=seed)
    cfg = VocosConfig(dim=state_dict["backbone.embed.weight"].shape[0],
                      intermediate_dim=state_dict["backbone.convnext.0.pwconv1.weight"].shape[0],
                      num_layers=1 + max(int(k.split(".")[2]) for k in state_dict if k.startswith("backbone.convnext.")))
    return Vocos(state_dict, cfg, device, precision=precision)


def load_model(model_cls=None, model_cfg=None, mel_spec_type=mel_spec_type, vocab_file="", ode_method=ode_method,
               use_ema=True, device=None, state_dict=None, ckpt_path: str | None = None, seed: int = 0,
               precision: str = "bf16") -> CFM:
    """utils_infer.py:224-260.  Like the reference fork, no checkpoint is read unless one is given explicitly
    (`state_dict=` or `ckpt_path=`, EMA key rules of :195-213); otherwise weights are seeded random-init."""
    device = device or _default_device()
    if ode_method != "euler":
        raise NotImplementedError("only the Euler solver is on the served path")
    if vocab_file == "":
        tokens = T.synthetic_indic_vocab()
        vocab_char_map, vocab_size = {t: i for i, t in enumerate(tokens)}, len(tokens)
    else:
        print("\nvocab : ", vocab_file)
        print("token : ", "custom")
        vocab_char_map, vocab_size = T.get_tokenizer(vocab_file, "custom")
    if ckpt_path is not None and state_dict is None:
        if ckpt_path.endswith(".safetensors"):
            from safetensors.torch import load_file
            state_dict = strip_checkpoint(load_file(ckpt_path), use_ema)
        else:
            state_dict = strip_checkpoint(torch.load(ckpt_path, map_location="cpu", weights_only=True), use_ema)
    if state_dict is not None:
        cfg = infer_dit_config(state_dict)
        if vocab_size > cfg.vocab_size:      # nn.Embedding would raise IndexError on the first out-of-range token (dit.py:56)
            raise ValueError(f"the vocabulary has {vocab_size} entries but the checkpoint's text embedding holds {cfg.vocab_size}")
    else:
        kw = dict(model_cfg or {})
        base = INDICF5
        cfg = DiTConfig(dim=kw.get("dim", base.dim), depth=kw.get("depth", base.depth), heads=kw.get("heads", base.heads),
                        ff_mult=kw.get("ff_mult", base.ff_mult), text_dim=kw.get("text_dim", base.text_dim),
                        conv_layers=kw.get("conv_layers", base.conv_layers), vocab_size=vocab_size)
        state_dict = make_dit_state_dict(cfg, seed=seed)
    return CFM(state_dict, cfg, vocab_char_map, device, precision=precision)


def preprocess_ref_audio_text(ref_audio_orig, ref_text, clip_short=True, show_info=print, device=None):
    """utils_infer.py:282-351.  Audio half (:285-320): clip to <= 15 s on silences, trim the silent edges at -42 dBFS, append
    50 ms of silence, re-export — `prompt_audio.py` restates the pydub semantics on integer PCM without pydub / ffmpeg.  Text
    half (:343-347): the sentence-final '. ' rule.  The Whisper fallback for an empty ref_text (:138-169, :327-338) is outside
    the served path (the server always passes a transcript, tts_utils.py:19) and is rejected loudly.  A prompt that is already a
    tensor pair `(audio, sr)` is passed through untouched (our `infer_process` accepts it directly)."""
    if not ref_text.strip():
        raise NotImplementedError("empty ref_text needs the ASR fallback (utils_infer.py:138-169), not on the served path")
    if isinstance(ref_audio_orig, (str, os.PathLike)):
        from .prompt_audio import preprocess_ref_audio
        ref_audio_orig = preprocess_ref_audio(ref_audio_orig, clip_short=clip_short, show_info=show_info)
    return ref_audio_orig, T.finish_ref_text(ref_text)


def _load_audio(path: str):
    try:
        import torchaudio
        audio, sr = torchaudio.load(path)
        return audio.float(), sr
    except Exception:
        import wave
        with wave.open(path, "rb") as w:
            sr, nch, sw, nfr = w.getframerate(), w.getnchannels(), w.getsampwidth(), w.getnframes()
            raw = w.readframes(nfr)
        if sw == 2:
            a = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
        elif sw == 4:
            a = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
        else:
            raise ValueError(f"unsupported WAV sample width {sw}")
        return torch.from_numpy(a.reshape(-1, nch).T.copy()), sr


@dataclass
class _Prepared:
    audio: torch.Tensor
    rms: float
    ref_len: int
    tokens: list
    duration: int
    noise_index: int
    noise_seed: int | None = None      # explicit Philox key (request scheduler: per-request seeds inside a shared pack)


@dataclass
class Staged:
    """One request batch resident on the device (output of `Synthesizer.stage`)."""
    ws: object
    generation: int
    layout: object
    preps: list
    frames: list
    offs: list
    total: int
    src_rows: torch.Tensor
    vpos: torch.Tensor
    gains: torch.Tensor
    seg: torch.Tensor
    nfe_step: int
    h2d_bytes: int


class Synthesizer:
    """Batched synthesis over (model, vocoder): the tensor part of `infer_batch_process` (utils_infer.py:423-482) for
    many independent utterances at once.  Inputs are host objects, outputs are host numpy arrays."""

    def __init__(self, model_obj: CFM, vocoder: Vocos):
        self.model, self.vocoder = model_obj, vocoder
        self.device = model_obj.device
        self.last_h2d_bytes = 0
        self.last_d2h_bytes = 0
        self.prompt_cache = PromptCache()
        self._host_wav: torch.Tensor | None = None
        self._lock = threading.RLock()     # stage -> run -> D2H of one call share workspaces and the landing buffer

    def _prep(self, spec: UtteranceSpec, speed_, fix_duration_) -> _Prepared:
        audio = spec.audio
        if audio.shape[0] > 1:
            audio = torch.mean(audio, dim=0, keepdim=True)                        # :424-425
        rms = float(torch.sqrt(torch.mean(torch.square(audio))))                  # :427
        if rms < target_rms:
            audio = audio * target_rms / rms
        sr = spec.meta.get("sr", target_sample_rate)
        if sr != target_sample_rate:                                              # :430-432
            import torchaudio
            audio = torchaudio.transforms.Resample(sr, target_sample_rate)(audio)
        ref_text = spec.ref_text
        if len(ref_text[-1].encode("utf-8")) == 1:                                # :438-439
            ref_text = ref_text + " "
        tokens = T.convert_char_to_pinyin([ref_text + spec.gen_text])[0]          # :443-444
        ref_len = audio.shape[-1] // hop_length                                   # :446
        if spec.duration is not None:
            duration = spec.duration
        else:
            duration = T.estimate_duration(ref_len, ref_text, spec.gen_text, speed_, fix_duration_)
        return _Prepared(audio, rms, ref_len, tokens, duration, spec.noise_index, spec.meta.get("noise_seed"))

    @torch.inference_mode()
    def stage(self, specs: list[UtteranceSpec], nfe_step=nfe_step, sway_sampling_coef=sway_sampling_coef, speed=speed,
              fix_duration=fix_duration, y0: list | None = None, noise_seed: int | None = None, slot: int = 0) -> "Staged":
        """Host side + H2D of one request batch: tokenise, duration rule, prompt RMS, pinned copies of prompt audio / tables,
        prompt mel and initial noise on the device.  After this the batch is resident in HBM.  `y0` (one [>= n_i, 100] tensor
        per utterance) injects the noise instead (parity tests); `noise_seed` fixes the device draw (default: a fresh one per
        call, cfm.py:181-186 — the reference is stochastic unless seeded)."""
        model, dev = self.model, self.device
        preps = [self._prep(s, speed, fix_duration) for s in specs]
        mels: list = [None] * len(preps)
        h2d = 0
        todo: dict[tuple, list[int]] = {}
        for i, p in enumerate(preps):                                             # one mel per distinct prompt (voice cache)
            k = PromptCache.key(p.audio)
            m = self.prompt_cache.get(k, dev)
            if m is not None:
                mels[i] = m
            else:
                todo.setdefault(k, []).append(i)
        if todo:
            firsts = [idx[0] for idx in todo.values()]
            host = torch.cat([preps[i].audio[0] for i in firsts]).pin_memory()    # all new prompts: one H2D copy, one launch
            h2d += host.numel() * 4
            flat = host.to(dev, non_blocking=True)
            waves_d, o = [], 0
            for i in firsts:
                n = preps[i].audio.shape[-1]
                waves_d.append(flat[o:o + n])
                o += n
            for (k, idx), m in zip(todo.items(), mel_rows(waves_d)):
                self.prompt_cache.put(k, dev, m)
                for i in idx:
                    mels[i] = m
        ids = T.list_str_to_idx([p.tokens for p in preps], model.vocab_char_map)
        base = fresh_noise_seed() if (y0 is None and noise_seed is None) else (noise_seed or 0)
        utts = []
        for i, (p, m, row) in enumerate(zip(preps, mels, ids)):                   # CFM.sample prologue, cfm.py:110-138
            tid = row[row != -1]
            lens_i = max(int(tid.numel()), m.shape[0])
            n = min(max(lens_i + 1, p.duration), 4096)
            utts.append(UtteranceInput(cond=m, text_ids=tid, n=n, cond_len=lens_i, y0=None if y0 is None else y0[i],
                                       noise_seed=p.noise_seed if p.noise_seed is not None else utterance_seed(base, p.noise_index)))
        ws, layout = model.engine.stage(utts, nfe_step, sway_sampling_coef, slot)
        h2d += ws.h2d_bytes
        veng = self.vocoder.engine                                                # vocoder rows = generated frames only
        frames = [n - p.ref_len for n, p in zip(layout.lengths, preps)]
        starts, Rv, pos, offs, tot = veng.plan(frames)
        src_rows = torch.full((Rv,), -1, dtype=torch.int32)
        for s, T_, ls, p in zip(starts, frames, layout.starts, preps):
            src_rows[s:s + T_] = torch.arange(ls + p.ref_len, ls + p.ref_len + T_, dtype=torch.int32)
        gains = torch.tensor([p.rms / target_rms if p.rms < target_rms else 1.0 for p in preps], dtype=torch.float32)
        seg = torch.tensor([[s, T_, o, 0] for s, T_, o in zip(starts, frames, offs)], dtype=torch.int32)
        h2d += (src_rows.numel() + pos.numel() + gains.numel() + seg.numel()) * 4
        st = Staged(ws, ws.generation, layout, preps, frames, offs, tot, src_rows.to(dev), pos.to(dev), gains.to(dev), seg.to(dev),
                    nfe_step, h2d)
        self.last_h2d_bytes = h2d
        return st

    @torch.inference_mode()
    def run(self, st: "Staged", cfg_strength=cfg_strength) -> torch.Tensor:
        """Device-resident hot path: sampler + vocoder on a staged batch -> flat fp32 waveform buffer on the device."""
        self.model.engine.compute(st.ws, st.nfe_step, cfg_strength, generation=st.generation)
        return self.vocoder.engine.decode_rows(st.ws.x, st.src_rows, st.vpos, st.seg, st.frames, st.total, st.gains)

    def generate_device(self, specs: list[UtteranceSpec], nfe_step=nfe_step, cfg_strength=cfg_strength,
                        sway_sampling_coef=sway_sampling_coef, speed=speed, fix_duration=fix_duration,
                        y0: list | None = None, noise_seed: int | None = None, slot: int = 0):
        st = self.stage(specs, nfe_step, sway_sampling_coef, speed, fix_duration, y0, noise_seed, slot)
        wav = self.run(st, cfg_strength)
        return wav, st.offs, st.frames, st.total, st.ws, st.layout, st.preps

    @torch.inference_mode()
    def generate(self, specs: list[UtteranceSpec], nfe_step=nfe_step, cfg_strength=cfg_strength,
                 sway_sampling_coef=sway_sampling_coef, speed=speed, fix_duration=fix_duration, y0: list | None = None,
                 return_mel: bool = False, noise_seed: int | None = None):
        """-> list of np.float32 waves (and optionally list of np mel [100, F_gen]); host in, host out.  Calls from several
        threads are serialised (the engine's workspaces and the pinned landing buffer belong to one batch at a time)."""
        with self._lock:
            return self._generate_locked(specs, nfe_step, cfg_strength, sway_sampling_coef, speed, fix_duration, y0, return_mel, noise_seed)

    def _generate_locked(self, specs, nfe_step, cfg_strength, sway_sampling_coef, speed, fix_duration, y0, return_mel, noise_seed):
        wav, offs, frames, tot, ws, layout, preps = self.generate_device(specs, nfe_step, cfg_strength, sway_sampling_coef,
                                                                         speed, fix_duration, y0, noise_seed)
        if self._host_wav is None or self._host_wav.numel() < tot:                # persistent pinned landing buffer (grown, never
            self._host_wav = torch.empty(max(tot, 1), dtype=torch.float32).pin_memory()   # re-pinned per request)
        host = self._host_wav[:tot]
        host.copy_(wav[:tot], non_blocking=True)                                  # D2H (:479)
        mel_out = None
        if return_mel:
            mel_out = [ws.x[ls + p.ref_len: ls + n, :n_mel_channels].t().cpu().numpy()
                       for ls, n, p in zip(layout.starts, layout.lengths, preps)]
        torch.cuda.current_stream().synchronize()
        self.last_d2h_bytes = tot * 4
        waves = [host[o:o + 256 * (T_ - 1)].numpy().copy() for o, T_ in zip(offs, frames)]   # caller-owned (the landing buffer is reused)
        return (waves, mel_out) if return_mel else waves


def _synthesizer_for(model_obj, vocoder) -> "Synthesizer":
    """One Synthesizer (prompt cache, pinned landing buffer, workspaces through the model's engine) per (model, vocoder) pair:
    `infer_process` is called once per request and must not rebuild them each time."""
    syn = getattr(model_obj, "_synthesizer", None)
    if syn is None or syn.vocoder is not vocoder:
        syn = Synthesizer(model_obj, vocoder)
        model_obj._synthesizer = syn
    return syn


def infer_batch_process(ref_audio, ref_text, gen_text_batches, model_obj, vocoder, mel_spec_type="vocos", progress=None,
                        target_rms=0.1, cross_fade_duration=0.15, nfe_step=32, cfg_strength=2.0, sway_sampling_coef=-1,
                        speed=1, fix_duration=None, device=None, y0: list | None = None, seed: int | None = None):
    """utils_infer.py:406-524.  The chunks of one text are independent until the cross-fade, so they are sampled as ONE
    packed batch instead of the reference's sequential loop; the returned triple is the reference's."""
    if mel_spec_type != "vocos":
        raise NotImplementedError("bigvgan is not on the served path")
    audio, sr = ref_audio
    specs = [UtteranceSpec(audio=audio, ref_text=ref_text, gen_text=g, duration=None, noise_index=i, meta={"sr": sr})
             for i, g in enumerate(gen_text_batches)]
    waves, mels = _synthesizer_for(model_obj, vocoder).generate(specs, nfe_step, cfg_strength, sway_sampling_coef, speed,
                                                                fix_duration, y0=y0, return_mel=True, noise_seed=seed)
    final_wave = cross_fade(waves, cross_fade_duration, target_sample_rate)                   # :485-519
    return final_wave, target_sample_rate, np.concatenate(mels, axis=1)


def infer_process(ref_audio, ref_text, gen_text, model_obj, vocoder, mel_spec_type=mel_spec_type, show_info=print,
                  progress=None, target_rms=target_rms, cross_fade_duration=cross_fade_duration, nfe_step=nfe_step,
                  cfg_strength=cfg_strength, sway_sampling_coef=sway_sampling_coef, speed=speed, fix_duration=fix_duration,
                  device=None, y0: list | None = None, seed: int | None = None):
    """utils_infer.py:357-400: chunk the text by the byte budget, then `infer_batch_process`.  The noise is a fresh device draw
    per call like the reference's (cfm.py:186); `seed` fixes it, `y0` (one tensor per text chunk) injects it (parity tests)."""
    audio, sr = _load_audio(ref_audio) if isinstance(ref_audio, (str, os.PathLike)) else ref_audio
    max_chars = int(len(ref_text.encode("utf-8")) / (audio.shape[-1] / sr) * (25 - audio.shape[-1] / sr))
    gen_text_batches = T.chunk_text(gen_text, max_chars=max_chars)
    return infer_batch_process((audio, sr), ref_text, gen_text_batches, model_obj, vocoder, mel_spec_type=mel_spec_type,
                               target_rms=target_rms, cross_fade_duration=cross_fade_duration, nfe_step=nfe_step,
                               cfg_strength=cfg_strength, sway_sampling_coef=sway_sampling_coef, speed=speed,
                               fix_duration=fix_duration, device=device, y0=y0, seed=seed)


class INF5Model:
    """Stand-in for the HF remote-code `ai4bharat/IndicF5` model object the server calls
    (`self.model(text, ref_audio_path=..., ref_text=...)`, managers.py:82-85): returns a 1-D 24 kHz numpy array.
    The conditioned prompt (utils_infer.py:285-320) is cached per voice, keyed by the md5 of the prompt file's bytes like the
    reference's own `_ref_audio_cache` (:322-325): the server re-downloads the same prompt for every request
    (tts_utils.py:31-36,54-58)."""

    def __init__(self, vocab_file: str = "", device: str | None = None, ckpt_path: str | None = None,
                 vocoder_path: str = "", seed: int = 0, output_int16: bool = True, model_cfg: dict | None = None,
                 state_dict: dict | None = None, vocoder_state_dict: dict | None = None, precision: str = "bf16"):
        device = device or _default_device()
        self.vocoder = load_vocoder("vocos", is_local=bool(vocoder_path), local_path=vocoder_path, device=device, seed=seed,
                                    state_dict=vocoder_state_dict, precision=precision)
        self.ema_model = load_model(None, model_cfg or dict(dim=1024, depth=22, heads=16, ff_mult=2, text_dim=512, conv_layers=4),
                                    vocab_file=vocab_file, device=device, ckpt_path=ckpt_path, seed=seed, state_dict=state_dict,
                                    precision=precision)
        self.output_int16 = output_int16
        self.noise_fn = None          # tests: callable (chunk index, frames) -> [frames, 100] noise; None: fresh device draw
        self._prompts: dict[str, tuple] = {}
        self._prompts_lock = threading.Lock()     # the batching route conditions prompts on worker threads (server.py)

    def to(self, device):
        return self

    def _prompt(self, ref_audio_path, ref_text):
        import hashlib
        with open(ref_audio_path, "rb") as f:
            key = hashlib.md5(f.read()).hexdigest() + "|" + ref_text
        with self._prompts_lock:
            hit = self._prompts.get(key)
        if hit is None:
            path, text = preprocess_ref_audio_text(ref_audio_path, ref_text, show_info=lambda *_: None)
            audio, sr = _load_audio(path)
            if path != os.fspath(ref_audio_path):
                os.unlink(path)                                   # the reference leaks its temp file (delete=False, :284)
            hit = ((audio, sr), text)
            with self._prompts_lock:
                if len(self._prompts) >= 64:
                    self._prompts.pop(next(iter(self._prompts)))
                self._prompts[key] = hit
        return hit

    def __call__(self, text: str, ref_audio_path: str, ref_text: str):
        ref_audio, ref_text = self._prompt(ref_audio_path, ref_text)
        y0 = None
        if self.noise_fn is not None:
            audio, sr = ref_audio
            max_chars = int(len(ref_text.encode("utf-8")) / (audio.shape[-1] / sr) * (25 - audio.shape[-1] / sr))
            y0 = [self.noise_fn(i, 4096) for i in range(len(T.chunk_text(text, max_chars=max_chars)))]
        wave, sr, _ = infer_process(ref_audio, ref_text, text, self.ema_model, self.vocoder, y0=y0)
        if self.output_int16:
            return np.clip(wave * 32768.0, -32768, 32767).astype(np.int16)
        return wave.astype(np.float32)


class TTSManager:
    """`src/server/core/managers.py:62-85`, byte-compatible surface: `.model` truthiness is the readiness probe
    (routes/speech.py:24), `load()` is idempotent and re-raises, `synthesize` raises ValueError when unloaded.
    One addition: a CUDA error is sticky for the process, so after one the manager drops the model (the readiness probe turns
    503 instead of every later request failing with a 500), keeps the kernel watchdog's record in `failed_reason`, and
    refuses to reload inside the dead context."""

    def __init__(self, device_type=None, **model_kwargs):
        self.device_type = device_type or _default_device()
        self.model = None
        self.repo_id = "ai4bharat/IndicF5"
        self._model_kwargs = model_kwargs
        self.failed_reason: str | None = None

    def load(self):
        if self.failed_reason is not None:
            raise RuntimeError(f"TTS engine lost its CUDA context ({self.failed_reason}); restart the process")
        if not self.model:
            logger.info("Loading TTS model IndicF5...")
            try:
                from . import _lib
                _lib.enable_diag()
                self.model = INF5Model(device=self.device_type, **self._model_kwargs)
                self.model = self.model.to(self.device_type)
                logger.info("TTS model IndicF5 loaded")
            except Exception as e:
                logger.error(f"Failed to load TTS model: {str(e)}")
                raise

    def synthesize(self, text, ref_audio_path, ref_text):
        if not self.model:
            raise ValueError("TTS model not loaded")
        try:
            return self.model(text, ref_audio_path=ref_audio_path, ref_text=ref_text)
        except RuntimeError as e:
            self.note_failure(e)
            raise

    def note_failure(self, e: BaseException) -> None:
        """A CUDA error is sticky for the process: drop the model and keep the watchdog's record (also called by the batching
        route, `server.py`, whose requests reach the engine through the scheduler instead of `synthesize`)."""
        if "CUDA error" in str(e) or "cudaError" in str(e):
            from . import _lib
            self.failed_reason = f"{str(e).splitlines()[0]}; watchdog record: {_lib.read_diag()}"
            logger.error(f"TTS engine failed: {self.failed_reason}")
            self.model = None


def wav_response_bytes(audio: np.ndarray, sample_rate: int = target_sample_rate) -> "io.BytesIO":
    """Response body of `synthesize_speech` (src/server/utils/tts_utils.py:60-65): an int16 array is rescaled by 1/32768,
    then the float wave is written as a RIFF/WAVE file, which libsndfile's default for format='WAV' stores as 16-bit PCM
    (sample = round-to-nearest-even of x * 32767).  Samples outside [-1, 1] are clipped here; libsndfile without
    SFC_SET_CLIPPING wraps them, which is audible garbage and not a behaviour worth matching.  Header layout is the
    canonical 44-byte PCM one (the same bytes `sf.write` emits for mono PCM_16)."""
    import io
    import struct
    a = np.asarray(audio)
    if a.ndim != 1:
        raise ValueError("mono 1-D audio expected")
    if a.dtype == np.int16:
        a = a.astype(np.float32) / 32768.0
    pcm = np.rint(np.clip(a.astype(np.float64), -1.0, 1.0) * 32767.0).astype("<i2")
    data = pcm.tobytes()
    buf = io.BytesIO()
    buf.write(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE")
    buf.write(b"fmt " + struct.pack("<IHHIIHH", 16, 1, 1, sample_rate, sample_rate * 2, 2, 16))
    buf.write(b"data" + struct.pack("<I", len(data)))
    buf.write(data)
    buf.seek(0)
    return buf
