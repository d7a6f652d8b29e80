"""BASELINE.json configs[4] (C5): Vocos vocoder-only sweep — 100-band mel, T in {256..8192} frames x batch {1..256} ->
24 kHz waveform via the fused ISTFT head.  Device-resident timing (CUDA events, median of 5 after 2 warm-ups); random-init
vocos-mel-24khz weights; mel = randn*2-4 (log-mel-like).  Prints a table and writes JSON.
  python tools/vocos_sweep.py [out.json]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tts_indic_server_f5_b200 import api, weights as W  # noqa: E402

FLOP_PER_FRAME = 26.99e6                      # embed 0.717 + 8 x 3.153 + head 1.051 MFLOP
ISTFT_BYTES_PER_FRAME = 4104 + 4096 + 1024    # spectrum in, windowed frame out (+ re-read by the overlap-add), waveform out

dev = torch.device("cuda")
voc = api.load_vocoder(state_dict=W.make_vocos_state_dict(W.VOCOS_24K, seed=0), device=dev)
eng = voc.engine
rows = []
g = torch.Generator("cpu").manual_seed(0)
for T in (256, 512, 1024, 2048, 4096, 8192):
    for B in (1, 4, 16, 64, 256):
        frames = [T] * B
        try:
            src = (torch.randn(B * T, 100, generator=g) * 2 - 4).to(dev)
            starts, Rv, pos, offs, tot = eng.plan(frames)
            src_rows = torch.full((Rv,), -1, dtype=torch.int32)
            for i, s in enumerate(starts):
                src_rows[s:s + T] = torch.arange(i * T, (i + 1) * T, dtype=torch.int32)
            seg = torch.tensor([[s, T, o, 0] for s, o in zip(starts, offs)], dtype=torch.int32).to(dev)
            src_rows, pos = src_rows.to(dev), pos.to(dev)
            ts = []
            for it in range(7):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                wav = eng.decode_rows(src, src_rows, pos, seg, frames, tot)
                e1.record()
                torch.cuda.synchronize()
                if it >= 2:
                    ts.append(e0.elapsed_time(e1))
            assert torch.isfinite(wav[:tot]).all()
            ms = sorted(ts)[len(ts) // 2]
            audio_s = B * 256 * (T - 1) / 24000.0
            r = dict(T=T, B=B, ms=ms, frames_per_s=B * T / ms * 1e3, audio_s_per_s=audio_s / ms * 1e3,
                     tflops=FLOP_PER_FRAME * B * T / ms / 1e9, istft_algorithmic_gbs_if_alone=ISTFT_BYTES_PER_FRAME * B * T / ms / 1e6)
            rows.append(r)
            print(f"T={T:5d} B={B:4d}  {ms:9.3f} ms  {r['frames_per_s'] / 1e6:7.2f} Mframe/s  {r['audio_s_per_s']:10.0f} x real time  "
                  f"{r['tflops']:7.1f} TFLOP/s", flush=True)
            del src, wav
            eng._bufs = {}
            torch.cuda.empty_cache()
        except torch.OutOfMemoryError:
            print(f"T={T} B={B}: out of memory, skipped", flush=True)
            eng._bufs = {}
            torch.cuda.empty_cache()
if len(sys.argv) > 1:
    json.dump(dict(workload="c5 vocos-only sweep", flop_per_frame=FLOP_PER_FRAME, rows=rows), open(sys.argv[1], "w"), indent=1)
