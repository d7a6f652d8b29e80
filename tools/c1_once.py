"""One single-utterance request (C1) with a few eager Euler steps: the launch list of the server's B = 1 case for ncu.
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/c1_once.py [euler_steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tts_indic_server_f5_b200 import api, synthetic as S, weights as W  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
model = api.load_model(state_dict=W.make_dit_state_dict(W.INDICF5, seed=0))
voc = api.load_vocoder(state_dict=W.make_vocos_state_dict(W.VOCOS_24K, seed=0))
model.engine.use_graphs = False
syn = api.Synthesizer(model, voc)
w = syn.generate(S.workload("c1"), nfe_step=steps, noise_seed=1)
torch.cuda.synchronize()
print("c1 ok", w[0].shape)
