"""ORACLE — test infrastructure only.  How far is the REFERENCE's own bf16 mode (whole-module `.to(bfloat16)`, the only
reduced-precision mode the reference has) from its fp32 result?  This is the yardstick the CUDA path's bf16-operand
error is graded against: ours_bf16-vs-ref_fp32 must not exceed ref_bf16-vs-ref_fp32.
Writes tests/golden/ref_bf16_error.json.   python -m oracle.ref_bf16_error [--full]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shims as R  # noqa: E402
from oracle.make_golden import GOLDEN, inject_noise, vocab_map  # noqa: E402
from tts_indic_server_f5_b200 import synthetic as S, text as T, weights as W  # noqa: E402


def run(cfg, sd, spec, dtype, steps=32):
    cfm = R.build_reference_cfm(sd, cfg, vocab_map()).to(dtype)
    audio = spec.audio
    rms = torch.sqrt(torch.mean(torch.square(audio)))
    if rms < 0.1:
        audio = audio * 0.1 / rms
    ref_text = spec.ref_text + (" " if len(spec.ref_text[-1].encode()) == 1 else "")
    toks = T.convert_char_to_pinyin([ref_text + spec.gen_text])
    with torch.inference_mode(), inject_noise(S.initial_noise(4096, spec.noise_index)):
        out, _ = cfm.sample(cond=audio, text=toks, duration=spec.duration, steps=steps, cfg_strength=2.0, sway_sampling_coef=-1.0)
    return out[0].float().numpy()


def main():
    res = {}
    cases = [("tiny", W.tiny_dit_config(), 1, "tiny")]
    if "--full" in sys.argv:
        cases.append(("full_c1", W.INDICF5, 0, "c1"))
    for name, cfg, seed, wl in cases:
        sd = W.make_dit_state_dict(cfg, seed=seed)
        spec = S.workload(wl)[0]
        a = run(cfg, sd, spec, torch.float32)
        b = run(cfg, sd, spec, torch.bfloat16)
        r0 = spec.meta["ref_len"]
        d = (b - a)[r0:]
        res[name] = dict(mel_rel_l2=float(np.linalg.norm(d) / np.linalg.norm(a[r0:])), mel_linf=float(np.abs(d).max()))
        print(name, res[name], flush=True)
    path = os.path.join(GOLDEN, "ref_bf16_error.json")
    old = json.load(open(path)) if os.path.exists(path) else {}
    old.update(res)
    json.dump(old, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
