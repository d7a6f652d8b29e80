#!/usr/bin/env python
"""bench.py — generated audio-seconds per wall-second (1/RTF) at NFE = 32, CFG 2.0, sway -1, Euler, bf16 operands.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                     (the reference algorithm's CPU path on this box's host cores)

A "step" is one pass of the hot path (CFM sampler + DiT + Vocos) over one request batch:
  N = 1 : BASELINE.json configs[1] — 64 mixed-length Indic utterances (C2), one packed batch.
  N > 1 : ONE request batch of 64*N utterances of the C4 distribution (N = 8: the 512 utterances of BASELINE.json configs[3]),
          sharded through the product path: `dist.shard_plan` (LPT partition by algorithmic cost, then row-budget packs per
          rank) -> `Synthesizer` per pack -> NCCL gather of the waveforms on rank 0.  Per-GPU work is fixed: weak scaling.
          `--scaling strong` keeps the batch at 512 utterances for every N (the strong-scaling table of DESIGN.md).
  value : whole-job audio-sec / wall-sec with the batch already resident in HBM (CUDA events, max over ranks)
  e2e   : the same batch through the public API from HOST inputs (prompt audio, text) to HOST waveforms every step: H2D of the
          inputs, prompt mel, device noise draw, sampler, vocoder, gather, D2H — nothing is pre-staged or pre-drawn.
Weights are seeded random-init of the IndicF5 / vocos-mel-24khz architectures, text / prompt are synthetic.
A launch failure is fatal: there is no retry (the JSON line carries "restarts": 0 by construction) and the kernels' watchdog
record, if any, is printed to stderr.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "generated audio-sec per wall-sec (1/RTF) at NFE=32"
UNIT = "audio-s/s"
PACK_ROWS = 180224          # row budget of one packed batch incl. the CFG duplicate (~5 GB of workspace): a 64-utterance C2 share fits


def flops_per_utterance(n: int, f_gen: int, nfe: int = 32) -> float:
    """64 DiT forwards (NFE x CFG pair) + Vocos."""
    return 2 * nfe * (387.305e6 * n + 90112.0 * n * n + 0.284e9) + 26.99e6 * f_gen


class ClockSampler:
    """nvidia-smi clocks + throttle reasons every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self) -> dict:
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def describe(workload: str, specs) -> str:
    gen = [s.meta["gen_frames"] * 256 / 24000 for s in specs]
    prompt = specs[0].audio.shape[-1] / 24000
    span = f"{min(gen):.1f}-{max(gen):.1f} s" if max(gen) - min(gen) > 0.05 else f"{gen[0]:.1f} s"
    return (f"{workload}: {len(specs)} Indic utterance{'s' if len(specs) != 1 else ''} ({prompt:.0f} s prompt, {span} generated), "
            "NFE 32, CFG 2.0 (pair batched), sway -1, Euler, Vocos 24 kHz")


# ---------------------------------------------------------------------------------------------- CPU legs (oracle port)
def _oracle_inputs(workload: str, index: int = 0):
    from oracle import f5_oracle as O
    from tts_indic_server_f5_b200 import synthetic as S, text as T, weights as W
    cfg, vcfg = W.INDICF5, W.VOCOS_24K
    sd, vsd = W.make_dit_state_dict(cfg, 0), W.make_vocos_state_dict(vcfg, 0)
    spec = S.workload(workload)[index]
    vocab = {t: i for i, t in enumerate(T.synthetic_indic_vocab())}
    audio, _ = O.rms_normalise(spec.audio)
    ref_text = spec.ref_text + (" " if len(spec.ref_text[-1].encode()) == 1 else "")
    ids = O.list_str_to_idx(T.convert_char_to_pinyin([ref_text + spec.gen_text]), vocab)
    return O, S, cfg, vcfg, sd, vsd, spec, audio, ids, S.initial_noise(4096, spec.noise_index)


def cpu_sample(workload: str, nfe_sample: int, reps: int, warmup: int, full_run: bool):
    """The reference algorithm on the host cores: the oracle port of the reference's `CFM.sample` + DiT + Vocos (bit-identical
    to the real reference modules on CPU, tests/test_oracle_golden.py::test_oracle_vs_real_reference_modules), fp32, all host
    threads.  One bounded step = ONE utterance of the workload, `nfe_sample` of its 32 Euler steps (2*nfe_sample DiT forwards)
    + mel + the full Vocos decode; the step is credited with nfe_sample/32 of the utterance's audio (every Euler step is the
    same work), so steps x ms_per_step is real wall time.  `full_run`: one un-extrapolated NFE-32 run of the same utterance
    first (outside the timed steps) to show that the credit rule holds."""
    O, S, cfg, vcfg, sd, vsd, spec, audio, ids, y0 = _oracle_inputs(workload)
    torch.set_num_threads(os.cpu_count() or 1)
    ref_len = audio.shape[-1] // 256
    audio_sec = S.generated_audio_seconds([spec])
    full = None
    with torch.inference_mode():
        if full_run:
            t0 = time.perf_counter()
            out = O.cfm_sample(sd, cfg, O.mel_spectrogram(audio).permute(0, 2, 1), ids, spec.duration, y0=y0, steps=32)
            O.vocos_decode(vsd, vcfg, out[:, ref_len:, :].permute(0, 2, 1))
            dt = time.perf_counter() - t0
            full = {"seconds": dt, "value": audio_sec / dt, "what": f"complete NFE-32 utterance (n={spec.duration} frames), un-extrapolated"}
        times = []
        for r in range(warmup + reps):
            t0 = time.perf_counter()
            cond = O.mel_spectrogram(audio).permute(0, 2, 1)
            out = O.cfm_sample(sd, cfg, cond, ids, spec.duration, y0=y0, steps=nfe_sample)
            O.vocos_decode(vsd, vcfg, out[:, ref_len:, :].permute(0, 2, 1))
            if r >= warmup:
                times.append(time.perf_counter() - t0)
    step_s = statistics.mean(times)
    credited = audio_sec * nfe_sample / 32.0
    cb = {"value": credited / step_s, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
          "sample": (f"1 of the {workload} utterances (n={spec.duration} frames): {nfe_sample} of 32 Euler steps ({2 * nfe_sample} DiT "
                     f"forwards) + prompt mel + full Vocos decode per step, fp32, credited {nfe_sample}/32 of its {audio_sec:.2f} s of audio; "
                     f"{step_s:.2f} s measured per step")}
    if full is not None:
        cb["full_run"] = full
    return cb, step_s


def gpu_eager_baseline(workload: str):
    """SURVEY.md §8d "same-box bar": the reference graph as torch library calls (cuBLAS / SDPA / cuDNN through the oracle port,
    which restates the reference's modules op for op) on THIS B200 under torch eager, one utterance, batch 1, NFE 32, CFG as two
    passes (cfm.py:162-176): fp32 with TF32 off (what the server deploys, managers.py:76) and bf16 autocast."""
    import contextlib
    O, S, cfg, vcfg, sd, vsd, spec, audio, ids, y0 = _oracle_inputs(workload)
    audio_sec = S.generated_audio_seconds([spec])
    out = {"utterance": f"1 of the {workload} utterances (n={spec.duration} frames), batch 1, NFE 32, two CFG passes per step"}
    dev = torch.device("cuda", torch.cuda.current_device())
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    sdd = {k: v.to(dev) for k, v in sd.items()}
    vsdd = {k: v.to(dev) for k, v in vsd.items()}
    a, i_, y = audio.to(dev), ids.to(dev), y0.to(dev)
    ref_len = audio.shape[-1] // 256
    for name in ("fp32", "bf16_autocast"):
        cast = torch.autocast("cuda", dtype=torch.bfloat16) if name != "fp32" else contextlib.nullcontext()
        try:
            with torch.inference_mode(), torch.device(dev), cast:      # factory calls inside the port (arange, hann_window, ...) land on the GPU
                for rep in range(2):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    cond = O.mel_spectrogram(a).permute(0, 2, 1)
                    o = O.cfm_sample(sdd, cfg, cond, i_, spec.duration, y0=y, steps=32)
                    O.vocos_decode(vsdd, vcfg, o[:, ref_len:, :].permute(0, 2, 1).float())
                    torch.cuda.synchronize()
                    dt_s = time.perf_counter() - t0
            out[name] = {"value": audio_sec / dt_s, "unit": UNIT, "ms": dt_s * 1e3}
        except Exception as e:  # noqa: BLE001
            out[name] = {"error": str(e).splitlines()[0][:200]}
    del sdd, vsdd
    torch.cuda.empty_cache()
    return out


def memory_kernel_rates(voc, dev, hbm_gbs: float) -> dict:
    """The HBM-bound kernels of the path timed ALONE on benchmark-sized inputs (CUDA events, median of 7 launches after 2
    warm-ups; inputs of 0.5 - 2 GB, far above the 126 MB L2): algorithmic bytes per launch / time, against the measured copy
    bandwidth of MEASURED_PEAKS.json.  LayerNorm at the C2 step's shape; the two Vocos-side kernels at 64 x 4096 frames (C5)."""
    import torch
    from tts_indic_server_f5_b200 import ops

    def med(fn):
        ts = []
        for it in range(9):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            if it >= 2:
                ts.append(a.elapsed_time(b))
        return sorted(ts)[len(ts) // 2]

    out = {}

    def rec(name, ms, nbytes, what):
        gbs = nbytes / ms / 1e6
        out[name] = {"us": ms * 1e3, "algorithmic_bytes": nbytes, "gbs": gbs, "frac_of_hbm_peak": gbs / hbm_gbs, "bytes": what}

    g = torch.Generator(device=dev).manual_seed(11)
    M, D = 158976, 1024
    x = torch.randn(M, D, generator=g, device=dev)
    y = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    sc, sh = torch.randn(D, generator=g, device=dev), torch.randn(D, generator=g, device=dev)
    rec("layernorm_mod", med(lambda: ops.layernorm_mod(x, y, sc, sh, 1.0)), M * D * 6, f"{M} rows x {D}: fp32 in + bf16 out")
    del x, y
    eng = voc.engine
    _, Rv, pos, _, _ = eng.plan([4096] * 64)
    pos = pos.to(dev)
    C = eng.cfg.dim
    x = torch.randn(Rv, C, generator=g, device=dev)
    y = torch.empty(Rv, C, device=dev, dtype=torch.bfloat16)
    blk = eng.blocks[0]
    rec("dwconv7_ln", med(lambda: ops.dwconv7_ln(x, y, pos, blk["dw_w"], blk["dw_b"], blk["ln_w"], blk["ln_b"])), Rv * C * 6,
        f"{Rv} rows x {C}: fp32 in + bf16 out")
    del x, y
    spec = torch.randn(Rv, 1152, generator=g, device=dev)
    frames = torch.empty(Rv, 1024, device=dev)
    rec("istft_frames", med(lambda: ops.call("f5_istft_frames", ops.ptr(spec), spec.stride(0), Rv, ops.ptr(eng.window), ops.ptr(frames),
                                             ops.stream_ptr())), Rv * (4104 + 4096), f"{Rv} frames: 1026 fp32 in + 1024 fp32 out")
    del spec, frames
    torch.cuda.empty_cache()
    return out


def emit(line: dict) -> None:
    """The ONE JSON line goes to the real stdout; everything else the process prints (NCCL banners, library chatter)
    was re-routed to stderr at start-up."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"],
                    help="fp32: split-operand GEMMs + fp32 attention (reported beside the headline, never instead of it)")
    ap.add_argument("--cpu-nfe", type=int, default=1, help="Euler steps per bounded CPU step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary blocks (C1 latency, C3, C5, torch-eager baseline)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        cb, step_s = cpu_sample(args.workload, max(args.cpu_nfe, 1), max(args.steps, 1), min(args.warmup, 1), full_run=True)
        line = {"metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
                "config": {"workload": f"{args.workload}: bounded CPU sample of the same workload (see cpu_baseline.sample)", "nfe": 32,
                           "cfg": 2.0, "sway": -1.0},
                "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return

    import torch.distributed as dist
    from tts_indic_server_f5_b200 import _lib, api, ops, synthetic as S, weights as W
    from tts_indic_server_f5_b200.dist import WaveGatherer, shard_plan

    dev = torch.device("cuda", local_rank)
    for attempt in range(20):          # a device still being released by the previous process (back-to-back runs) is retried, not fatal
        try:
            torch.cuda.set_device(local_rank)
            torch.zeros(1, device=dev)
            torch.cuda.synchronize()
            break
        except RuntimeError as e:
            if attempt == 19:
                raise
            sys.stderr.write(f"bench: CUDA device not ready ({str(e).splitlines()[0]}); retrying\n")
            time.sleep(2.0)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.enable_diag()

    model = api.load_model(state_dict=W.make_dit_state_dict(W.INDICF5, seed=0), device=dev, precision=args.precision)
    voc = api.load_vocoder(state_dict=W.make_vocos_state_dict(W.VOCOS_24K, seed=0), device=dev, precision=args.precision)
    model.engine.max_workspaces = 8
    syn = api.Synthesizer(model, voc)
    syn.prompt_cache.capacity = 0      # every step is a fresh request: prompt audio H2D + prompt mel are redone each time

    # ------------------------------------------------------------------ the request batch and its sharding (product path)
    if world == 1:
        wl_name, all_specs = args.workload, S.workload(args.workload)
    else:
        wl_name = "c4"
        all_specs = S.workload("c4")
        if args.scaling == "weak":
            all_specs = all_specs[: 64 * world]
    plan = shard_plan([s.duration for s in all_specs], world, max_rows=PACK_ROWS)
    mine = [all_specs[i] for i in plan["parts"][rank]]
    packs = [[mine[j] for j in p] for p in plan["packs"][rank]]
    audio_sec_total = S.generated_audio_seconds(all_specs)
    samples_per_rank = [sum(256 * (all_specs[i].meta["gen_frames"] - 1) for i in part) for part in plan["parts"]]
    gatherer = WaveGatherer(samples_per_rank, dev) if world > 1 else None
    flat = torch.zeros(max(samples_per_rank[rank], 1), device=dev)      # this rank's waveforms, packs back to back

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def collect(wavs):
        """Pack outputs -> this rank's flat buffer -> (N > 1) NCCL gather on rank 0."""
        o = 0
        for w, st in wavs:
            flat[o:o + st.total].copy_(w[: st.total])
            o += st.total
        if gatherer is not None:
            gatherer.gather(flat)

    # ------------------------------------------------------------------ device-resident timing (value)
    staged = [syn.stage(p, noise_seed=1000 * rank + k, slot=k) for k, p in enumerate(packs)]
    for _ in range(args.warmup):
        collect([(syn.run(st), st) for st in staged])
    barrier()
    n_launch0 = _lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        e0.record()
        for _ in range(args.steps):
            collect([(syn.run(st), st) for st in staged])
        e1.record()
        barrier()
    launches = _lib.launch_count - n_launch0
    dt = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dt = float(dt)
    value = audio_sec_total * args.steps / dt

    # ------------------------------------------------------------------ end to end through the public API (e2e)
    host_land = torch.zeros(flat.numel(), dtype=torch.float32).pin_memory() if world == 1 else None
    h2d_step = [0]

    def e2e_step():
        wavs, h2d = [], 0
        for k, p in enumerate(packs):                                   # host specs -> stage (H2D, mel, noise) -> run, per pack
            st = syn.stage(p, slot=k)
            h2d += st.h2d_bytes
            wavs.append((syn.run(st), st))
        collect(wavs)
        h2d_step[0] = h2d
        if world == 1:
            host_land.copy_(flat, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return host_land.numpy()
        return gatherer.to_host()                                        # rank 0: D2H of every rank's waveforms; others: None

    e2e_steps = args.steps if (world == 1 or args.scaling == "weak") else max(2, args.steps // 4)
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    dt_e2e = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt_e2e, op=dist.ReduceOp.MAX)
    e2e_value = audio_sec_total * e2e_steps / float(dt_e2e)
    d2h = sum(samples_per_rank) * 4 if rank == 0 else 0

    # ------------------------------------------------------------------ roofline of the dominant kernel (tcgen05 GEMM, BLOCK_N = 256)
    # One instrumented eager Euler step on the first pack: CUDA events around every layer GEMM launch (QKV / out / FF1 / FF2).
    st = staged[0]
    recs, arecs, orig_gemm, orig_attn = [], [], ops.gemm, ops.attention

    def timed_gemm(A, B, **kw):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        orig_gemm(A, B, **kw)
        b.record()
        recs.append((A.shape[0], kw.get("N") or B.shape[0], B.shape[1], kw.get("num_taps", 1), a, b))

    def timed_attn(*a_, **kw):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        orig_attn(*a_, **kw)
        b.record()
        arecs.append((a, b))

    eng = model.engine
    eng.use_graphs, ops.gemm, ops.attention = False, timed_gemm, timed_attn
    try:
        eng.step(st.ws, 0, 2.0)
        torch.cuda.synchronize()
    finally:
        eng.use_graphs, ops.gemm, ops.attention = True, orig_gemm, orig_attn
    real_rows = 2 * st.layout.real_tokens                      # CFG pair; gap rows are not algorithmic work
    if args.precision == "fp32":             # split-operand launches: B holds three row-stacked planes, three products per k-block
        recs = [(M, N // 3 if taps == 1 else N, K, taps, a, b) for (M, N, K, taps, a, b) in recs]
    layer = [(N, K, a.elapsed_time(b) * 1e-3) for (M, N, K, taps, a, b) in recs if taps == 1 and N % 256 == 0 and K >= 1024]
    g_flops = sum(2.0 * real_rows * N * K for N, K, _ in layer)
    g_time = sum(t for _, _, t in layer)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    # the GEMMs are timed inside a live step of a run that has been loading the board for many seconds: the sustained peak
    # is the denominator (the burst figure is kept alongside)
    peak_tf = peaks.get("bf16_tflops_sustained", 1361.0)
    burst_tf = peaks.get("bf16_tflops", 1590.0)
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "gemm_traffic.json")))
        traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source", "profiles/gemm_traffic.json (ncu --set full capture, not measured in this run)")
    except Exception:
        pass
    achieved = g_flops / g_time / 1e12 if g_time > 0 else 0.0
    a_time = sum(a.elapsed_time(b) * 1e-3 for a, b in arecs)
    a_flops = len(arecs) * 2.0 * sum(4.0 * 1024 * n * n for n in st.layout.lengths)    # QK^T + PV, both CFG halves, per layer
    roofline = {"bound": "tensor", "kernel": f"gemm_tcgen05_kernel<256> (QKV/out/FF1/FF2, {len(layer)} launches of one Euler step)",
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                "peak_source": ("MEASURED_PEAKS.json bf16_tflops_sustained (kernels timed with CUDA events inside one live Euler step "
                                "right after the timed region)") if peaks else "fallback 1361 TFLOP/s sustained",
                "frac_of_burst_peak": achieved / burst_tf, "burst_peak": burst_tf,
                "avg_launch_ms": g_time / max(len(layer), 1) * 1e3, "traffic": traffic, "traffic_source": traffic_src,
                "secondary": {"kernel": f"attn_d64_kernel ({len(arecs)} launches of the same step)", "bound": "mufu+tensor",
                              "achieved": a_flops / a_time / 1e12 if a_time > 0 else 0.0, "unit": "TFLOP/s",
                              "avg_launch_ms": a_time / max(len(arecs), 1) * 1e3}}
    total_flops = sum(flops_per_utterance(s.duration, s.meta["gen_frames"]) for s in all_specs)
    job_tflops = total_flops * args.steps / dt / 1e12 / world
    sustained = peaks.get("bf16_tflops_sustained", 1400.0)

    # ------------------------------------------------------------------ secondary blocks (rank 0, N = 1): reported, not the headline
    extras = {}
    if rank == 0 and world == 1 and not args.no_extras:
        del staged
        c1 = S.workload("c1")
        for _ in range(3):
            syn.generate(c1)
        lat = []
        for _ in range(10):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            syn.generate(c1)
            lat.append(time.perf_counter() - t0)
        c1_audio = S.generated_audio_seconds(c1)
        extras["latency_c1"] = {"workload": describe("c1", c1), "ms_median": statistics.median(lat) * 1e3, "ms_min": min(lat) * 1e3,
                                "x_realtime": c1_audio / statistics.median(lat),
                                "what": "host prompt + text -> host waveform through Synthesizer.generate (the server's B = 1 request)"}
        c3 = S.workload("c3")
        st3 = syn.stage(c3, noise_seed=3)
        syn.run(st3)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(2):
            syn.run(st3)
        b.record()
        torch.cuda.synchronize()
        extras["c3"] = {"workload": describe("c3", c3), "value": S.generated_audio_seconds(c3) * 2 / (a.elapsed_time(b) * 1e-3),
                        "unit": UNIT, "ms_per_step": a.elapsed_time(b) / 2}
        del st3
        T5, B5 = 2048, 64
        mel = (torch.randn(B5, 100, T5, generator=torch.Generator().manual_seed(5)) * 2 - 4).to(dev)
        voc.decode(mel)
        torch.cuda.synchronize()
        a.record()
        for _ in range(3):
            voc.decode(mel)
        b.record()
        torch.cuda.synchronize()
        extras["c5"] = {"workload": f"Vocos only: {B5} x {T5} frames, 100-band mel -> 24 kHz", "value": 3 * B5 * T5 / (a.elapsed_time(b) * 1e-3) / 1e6,
                        "unit": "Mframe/s"}
        del mel
        torch.cuda.empty_cache()
        try:
            extras["memory_kernels"] = memory_kernel_rates(voc, dev, peaks.get("hbm_gbs", 6551.0))
        except Exception as e:                      # a secondary block never takes the headline line down with it
            extras["memory_kernels"] = {"error": f"{type(e).__name__}: {e}"}
        extras["gpu_eager_baseline"] = gpu_eager_baseline(args.workload)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline, _ = cpu_sample(args.workload, args.cpu_nfe, 1, 0, full_run=False)

    if rank == 0:
        npk = [len(p) for p in plan["packs"]]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": args.scaling if world > 1 else "weak",
            "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "fp32x3 (hi+lo bf16 operand planes, fp32 attention)",
            "data": "synthetic", "restarts": 0,
            "config": {"workload": describe(wl_name, all_specs), "utterances": len(all_specs),
                       "utterances_per_gpu": [len(p) for p in plan["parts"]], "packs_per_gpu": npk,
                       "lpt_imbalance": plan["imbalance"], "tokens_rank0": sum(s.duration for s in mine),
                       "weights": "random-init IndicF5 DiT (dim 1024, depth 22, 16 heads) + vocos-mel-24khz",
                       "noise": "drawn on the device per request (Philox, f5_randn_rows), inside the e2e region",
                       "l2": "activations per pass (GBs) exceed the 126 MB L2; no flush needed",
                       "parallelism": f"utterance-sharded x{world} (LPT + row-budget packs), NCCL gather of waveforms on rank 0"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_step[0], "d2h_bytes_per_step": d2h, "steps": e2e_steps},
            "gpu_launches": launches,
            "roofline": roofline,
            "model_tflops_per_gpu": job_tflops, "model_frac_of_sustained_peak": job_tflops / sustained,
            "clocks": clocks.summary(),
        }
        line.update(extras)
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    try:
        main()
    except Exception:
        import traceback
        traceback.print_exc()
        try:
            from tts_indic_server_f5_b200 import _lib
            sys.stderr.write(f"bench: FAILED, no retry.  kernel watchdog record: {_lib.read_diag()}\n")
        except Exception:
            pass
        sys.stderr.flush()
        os._exit(1)
