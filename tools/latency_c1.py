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
    print(f"{os.environ.get('F5_SMALL_M_TILES', 'default')} pdl={os.environ.get('F5_PDL', 'auto')} {wl}: median {statistics.median(ts) * 1e3:.2f} ms  min {min(ts) * 1e3:.2f} ms "
          f"-> {S.generated_audio_seconds(specs) / statistics.median(ts):.1f} x real time", flush=True)
