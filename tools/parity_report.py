"""Print parity metrics of the CUDA path against the committed golden vectors (real reference outputs) and the oracle.
Runs on the GPU box:  python tools/parity_report.py [--full]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tts_indic_server_f5_b200 import api, synthetic as S, weights as W  # noqa: E402
from tts_indic_server_f5_b200.engine import UtteranceInput  # noqa: E402


def metrics(got, ref):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    d = got - ref
    return dict(linf=float(np.abs(d).max()), rel_l2=float(np.linalg.norm(d) / max(np.linalg.norm(ref), 1e-30)),
                snr_db=float(10 * np.log10((ref ** 2).sum() / max((d ** 2).sum(), 1e-30))), ref_max=float(np.abs(ref).max()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    args = ap.parse_args()
    out = {}
    gold = np.load(os.path.join(ROOT, "tests", "golden", "tiny.npz"))
    cfg, vcfg = W.tiny_dit_config(), W.tiny_vocos_config()
    model = api.load_model(state_dict=W.make_dit_state_dict(cfg, seed=1))
    voc = api.load_vocoder(state_dict=W.make_vocos_state_dict(vcfg, seed=1))
    # A. single CFG forward pair
    x = torch.from_numpy(gold["fwd_x"]); cond = torch.from_numpy(gold["fwd_condin"]); text = torch.from_numpy(gold["fwd_text"])
    u = UtteranceInput(cond=cond, text_ids=text, n=96, cond_len=96, y0=x)
    pc = model.engine.forward_flow([u], 0.37)[0].cpu().numpy()
    out["tiny_fwd_cond"] = metrics(pc[0], gold["fwd_cond"])
    out["tiny_fwd_null"] = metrics(pc[1], gold["fwd_null"])
    print(json.dumps({k: out[k] for k in ("tiny_fwd_cond", "tiny_fwd_null")}), flush=True)
    # B. end-to-end generate at tiny dims
    syn = api.Synthesizer(model, voc)
    for wl in ("tiny", "tiny3"):
        specs = S.workload(wl)
        waves, mels = syn.generate(specs, return_mel=True, y0=S.reference_noise(specs))
        for i, (wv, ml) in enumerate(zip(waves, mels)):
            gmel = gold[f"{wl}_{i}_mel"]
            ref_len = specs[i].meta["ref_len"]
            out[f"{wl}_{i}_mel"] = metrics(ml.T, gmel[ref_len:])
            out[f"{wl}_{i}_wave"] = metrics(wv, gold[f"{wl}_{i}_wave"])
            print(wl, i, json.dumps(out[f"{wl}_{i}_mel"]), json.dumps(out[f"{wl}_{i}_wave"]), flush=True)
    # graphs off must agree bit-for-bit with graphs on
    model.engine.use_graphs = False
    w2 = syn.generate(S.workload("tiny3"), noise_seed=1)
    w1 = api.Synthesizer(api.load_model(state_dict=W.make_dit_state_dict(cfg, seed=1)), voc).generate(S.workload("tiny3"), noise_seed=1)
    print("graph vs eager identical:", all(np.array_equal(a, b) for a, b in zip(w1, w2)), flush=True)
    if args.full:
        del model, voc, syn
        torch.cuda.empty_cache()
        goldf = np.load(os.path.join(ROOT, "tests", "golden", "full_c1.npz"))
        t0 = time.time()
        model = api.load_model(state_dict=W.make_dit_state_dict(W.INDICF5, seed=0))
        voc = api.load_vocoder(state_dict=W.make_vocos_state_dict(W.VOCOS_24K, seed=0))
        print(f"full weights built+loaded in {time.time()-t0:.1f}s", flush=True)
        syn = api.Synthesizer(model, voc)
        spec = S.workload("c1")
        for rep in range(2):
            torch.cuda.synchronize(); t0 = time.time()
            waves, mels = syn.generate(spec, return_mel=True, y0=S.reference_noise(spec))
            torch.cuda.synchronize(); dt = time.time() - t0
            print(f"c1 generate rep{rep}: {dt*1e3:.1f} ms -> {S.generated_audio_seconds(spec)/dt:.1f}x real-time", flush=True)
        ref_len = spec[0].meta["ref_len"]
        out["full_c1_mel"] = metrics(mels[0].T, goldf["mel"][ref_len:])
        out["full_c1_wave"] = metrics(waves[0], goldf["wave"])
        print("full_c1", json.dumps(out["full_c1_mel"]), json.dumps(out["full_c1_wave"]), flush=True)
        # vocoder alone on the golden mel (isolates the vocoder's own error)
        wv = voc.decode(torch.from_numpy(goldf["mel"][ref_len:].T[None].copy()))[0].cpu().numpy()
        out["full_vocos_only_wave"] = metrics(wv, goldf["wave"])
        print("vocos_only", json.dumps(out["full_vocos_only_wave"]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_report.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
