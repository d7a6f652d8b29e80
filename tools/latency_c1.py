"""Single-request latency (C1 and the 8 s variant) through Synthesizer.generate: median of 20 calls after warm-up."""
import os
import statistics
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tts_indic_server_f5_b200 import api, synthetic as S, weights as W  # noqa: E402

model = api.load_model(state_dict=W.make_dit_state_dict(W.INDICF5, seed=0))
voc = api.load_vocoder(state_dict=W.make_vocos_state_dict(W.VOCOS_24K, seed=0))
syn = api.Synthesizer(model, voc)
for wl in ("c1", "c1_8s"):
    specs = S.workload(wl)
    for _ in range(3):
        syn.generate(specs)
    ts = []
    for _ in range(20):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        syn.generate(specs)
        ts.append(time.perf_counter() - t0)
    print(f"{os.environ.get('F5_SMALL_M_RULE', 'default')} pdl={os.environ.get('F5_PDL', 'auto')} {wl}: median {statistics.median(ts) * 1e3:.2f} ms  min {min(ts) * 1e3:.2f} ms "
          f"-> {S.generated_audio_seconds(specs) / statistics.median(ts):.1f} x real time", flush=True)

# where a single request's time goes: host staging (tokenise, prompt key, layout, pinned copies + H2D, prompt mel, noise),
# device (hoist + 32-step graph + vocoder), D2H
specs = S.workload("c1")
stg, run, d2h = [], [], []
for _ in range(20):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st = syn.stage(specs)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    wav = syn.run(st)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    host = wav[: st.total].cpu()
    t3 = time.perf_counter()
    stg.append(t1 - t0); run.append(t2 - t1); d2h.append(t3 - t2)
med = lambda v: statistics.median(v) * 1e3  # noqa: E731
print(f"c1 breakdown (each part synchronised): stage {med(stg):.2f} ms | hoist + 32 Euler steps + vocoder {med(run):.2f} ms | D2H {med(d2h):.2f} ms", flush=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ws = st.ws
eng = model.engine
eng.compute(ws, 32, 2.0)
torch.cuda.synchronize()
e0.record()
eng.run_steps(ws, 32, 2.0)
e1.record()
torch.cuda.synchronize()
print(f"c1 32-step graph replay alone: {e0.elapsed_time(e1):.2f} ms (device time)", flush=True)
