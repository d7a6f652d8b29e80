#!/usr/bin/env python
"""bench.py — generated audio-seconds per wall-second (1/RTF) at NFE = 32, CFG 2.0, sway -1, Euler, bf16 operands.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                     (the reference algorithm's CPU path on this box's host cores)

A "step" is one pass of the hot path (CFM sampler + DiT + Vocos) over one request batch of 64 mixed-length Indic
utterances (BASELINE.json configs[1]; at N ranks every rank has its own 64 => 512 utterances at N = 8 = configs[3],
weak scaling, with the NCCL gather of waveforms on rank 0 inside the timed region).
  value : whole-job audio-sec / wall-sec with the batch already resident in HBM (CUDA events, max over ranks)
  e2e   : same metric through `Synthesizer.generate` — host buffers in, host numpy waveforms out, every step
Weights are seeded random-init of the IndicF5 / vocos-mel-24khz architectures, text/prompt/noise are synthetic.
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


def cpu_reference_leg(workload: str, nfe_sample: int, reps: int, warmup: int):
    """The reference algorithm on the host cores (oracle port of the reference's CFM.sample + DiT + Vocos; bit-identical
    to the real reference modules on CPU, tests/test_oracle_vs_reference.py).  Bounded sample: ONE utterance of the
    workload, `nfe_sample` of the 32 Euler steps (2*nfe_sample DiT forwards) + the full Vocos decode, extrapolated
    linearly to NFE 32."""
    from oracle import f5_oracle as O
    from tts_indic_server_f5_b200 import synthetic as S, text as T, weights as W
    torch.set_num_threads(os.cpu_count() or 1)
    cfg, vcfg = W.INDICF5, W.VOCOS_24K
    sd, vsd = W.make_dit_state_dict(cfg, 0), W.make_vocos_state_dict(vcfg, 0)
    spec = S.workload(workload)[0]
    vocab = {t: i for i, t in enumerate(T.synthetic_indic_vocab())}
    audio, _ = O.rms_normalise(spec.audio)
    ref_len = audio.shape[-1] // 256
    ref_text = spec.ref_text + (" " if len(spec.ref_text[-1].encode()) == 1 else "")
    ids = O.list_str_to_idx(T.convert_char_to_pinyin([ref_text + spec.gen_text]), vocab)
    y0 = S.initial_noise(4096, spec.noise_index)
    times = []
    with torch.inference_mode():
        for r in range(warmup + reps):
            t0 = time.perf_counter()
            cond = O.mel_spectrogram(audio).permute(0, 2, 1)
            t1 = time.perf_counter()
            out = O.cfm_sample(sd, cfg, cond, ids, spec.duration, y0=y0, steps=nfe_sample)
            t2 = time.perf_counter()
            O.vocos_decode(vsd, vcfg, out[:, ref_len:, :].permute(0, 2, 1))
            t3 = time.perf_counter()
            if r >= warmup:
                times.append(((t1 - t0) + (t2 - t1) * 32.0 / nfe_sample + (t3 - t2), t3 - t0))
    audio_sec = S.generated_audio_seconds([spec])
    est = statistics.mean(t[0] for t in times)
    return {"value": audio_sec / est, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"1 of the {workload} utterances (n={spec.duration} frames), {nfe_sample} of 32 Euler steps "
                      f"({2 * nfe_sample} DiT forwards) + full Vocos decode, fp32, extrapolated linearly to NFE 32; "
                      f"{statistics.mean(t[1] for t in times):.1f} s measured per sample"}, est


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
    ap.add_argument("--cpu-nfe", type=int, default=1, help="Euler steps in the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        cb, est = cpu_reference_leg(args.workload, args.cpu_nfe, max(args.steps, 1), min(args.warmup, 1))
        line = {"metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": est * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
                "config": {"workload": f"{args.workload}: bounded CPU sample of the same workload", "nfe": 32, "cfg": 2.0,
                           "sway": -1.0},
                "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return

    import torch.distributed as dist
    from tts_indic_server_f5_b200 import _lib, api, ops, synthetic as S, weights as W
    from tts_indic_server_f5_b200.dist import gather_waveforms

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

    model = api.load_model(state_dict=W.make_dit_state_dict(W.INDICF5, seed=0), device=dev)
    voc = api.load_vocoder(state_dict=W.make_vocos_state_dict(W.VOCOS_24K, seed=0), device=dev)
    syn = api.Synthesizer(model, voc)
    syn.prompt_cache.capacity = 0      # every step is a fresh request: prompt audio H2D + prompt mel are redone each time
    specs = S.workload(args.workload, seed=rank)
    noise = [S.initial_noise(4096, s.noise_index + 1000 * rank) for s in specs]   # inputs: prepared before any timing
    audio_sec_rank = S.generated_audio_seconds(specs)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather(wav, st):
        if world > 1:
            return gather_waveforms(wav[: st.total], [256 * (f - 1) for f in st.frames])
        return None

    # ------------------------------------------------------------------ device-resident timing (value)
    st = syn.stage(specs, y0=noise)
    for _ in range(args.warmup):
        gather(syn.run(st), st)
    barrier()
    n_launch0 = _lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        e0.record()
        for _ in range(args.steps):
            gather(syn.run(st), st)
        e1.record()
        barrier()
    launches = _lib.launch_count - n_launch0
    dt = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dt = float(dt)
    value = audio_sec_rank * world * args.steps / dt

    # ------------------------------------------------------------------ end to end through the public API (e2e)
    def e2e_step():
        if world == 1:
            return syn.generate(specs, y0=noise)
        wav, offs, frames, tot, *_ = syn.generate_device(specs, y0=noise)
        got = gather_waveforms(wav[:tot], [256 * (f - 1) for f in frames])
        if got is not None:
            flat = torch.cat([g[0] for g in got])
            host = torch.empty(flat.numel(), dtype=torch.float32).pin_memory()
            host.copy_(flat, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return host
        return None

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    dt_e2e = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt_e2e, op=dist.ReduceOp.MAX)
    e2e_value = audio_sec_rank * world * args.steps / float(dt_e2e)
    d2h = st.total * 4 * (world if rank == 0 else 1)

    # ------------------------------------------------------------------ roofline of the dominant kernel (tcgen05 GEMM, BLOCK_N = 256)
    # One instrumented eager Euler step: CUDA events around every layer GEMM launch (QKV / out / FF1 / FF2, 22 layers).
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
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "gemm_traffic.json"))).get("dram_bytes_per_launch")
    except Exception:
        pass
    achieved = g_flops / g_time / 1e12 if g_time > 0 else 0.0
    a_time = sum(a.elapsed_time(b) * 1e-3 for a, b in arecs)
    a_flops = len(arecs) * 2.0 * sum(4.0 * 1024 * n * n for n in st.layout.lengths)    # QK^T + PV, both CFG halves, per layer
    roofline = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel<256> (QKV/out/FF1/FF2, 88 launches of one Euler step)",
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                "peak_source": ("MEASURED_PEAKS.json bf16_tflops_sustained (kernels timed with CUDA events inside one live Euler step "
                                "right after the timed region)") if peaks else "fallback 1361 TFLOP/s sustained",
                "frac_of_burst_peak": achieved / burst_tf, "burst_peak": burst_tf,
                "avg_launch_ms": g_time / max(len(layer), 1) * 1e3, "traffic": traffic,
                "secondary": {"kernel": "attn_d64_kernel (22 launches of the same step)", "bound": "mufu+tensor",
                              "achieved": a_flops / a_time / 1e12 if a_time > 0 else 0.0, "unit": "TFLOP/s",
                              "avg_launch_ms": a_time / max(len(arecs), 1) * 1e3}}
    total_flops = sum(flops_per_utterance(n, n - p.ref_len) for n, p in zip(st.layout.lengths, st.preps)) * world
    job_tflops = total_flops * args.steps / dt / 1e12 / world
    sustained = peaks.get("bf16_tflops_sustained", 1400.0)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline, _ = cpu_reference_leg(args.workload, args.cpu_nfe, 1, 0)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {len(specs)} mixed-length Indic utterances per GPU (5 s prompt, 6-10 s "
                                   "generated), NFE 32, CFG 2.0 (pair batched), sway -1, Euler, Vocos 24 kHz",
                       "utterances_per_gpu": len(specs), "tokens_per_gpu": st.layout.real_tokens,
                       "weights": "random-init IndicF5 DiT (dim 1024, depth 22, 16 heads) + vocos-mel-24khz",
                       "l2": "activations per pass (GBs) exceed the 126 MB L2; no flush needed",
                       "parallelism": f"utterance-sharded x{world}, NCCL gather of waveforms"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": st.h2d_bytes, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "roofline": roofline,
            "model_tflops_per_gpu": job_tflops, "model_frac_of_sustained_peak": job_tflops / sustained,
            "clocks": clocks.summary(),
        }
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def _fake_worker() -> None:
    """Stand-in for the measurement (tests/test_bench_supervisor.py, CPU only): F5_BENCH_FAKE = "fail=<rank>" makes that rank's
    worker die on the first attempt; every other worker blocks (as a rank stuck in a collective would) until it is killed or,
    on a healthy attempt, prints the JSON line on rank 0."""
    spec = os.environ["F5_BENCH_FAKE"]
    rank, attempt = int(os.environ.get("RANK", "0")), int(os.environ.get("F5_BENCH_ATTEMPT", "0"))
    failing = int(spec.split("=")[1]) if spec.startswith("fail=") else -1
    if attempt == 0 and failing >= 0:
        if rank == failing:
            time.sleep(1.0)
            sys.stderr.write(f"fake worker rank {rank}: simulated launch failure\n")
            sys.exit(3)
        time.sleep(600.0)                       # stuck in the collective the dead rank never joins
    time.sleep(0.5)
    if spec == "pg":                            # a real process group on the child's own rendezvous port (gloo; the product uses nccl)
        import torch.distributed as dist
        dist.init_process_group("gloo")
        t = torch.tensor([float(rank + 1)])
        dist.all_reduce(t)
        assert float(t) == sum(range(1, dist.get_world_size() + 1))
        dist.destroy_process_group()
    if rank == 0:
        emit({"metric": METRIC, "value": 1.0, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "fake": True,
              "attempt": attempt, "master_port": os.environ.get("MASTER_PORT")})


def supervise_ranks() -> int:
    """Multi-rank runs (launched by torchrun): this process supervises the real measurement, which runs in a child process
    with its own rendezvous port.  A rank that dies (a CUDA launch failure is sticky for its process) cannot be restarted
    alone — its peers are inside NCCL collectives — so the supervisors agree through a small TCPStore: as soon as one child
    fails, every supervisor kills its child, and all ranks start ONE fresh attempt together.  Rank 0 forwards its child's
    JSON line.  Single-GPU runs use the simpler in-place restart below."""
    import datetime
    import subprocess
    from torch.distributed import TCPStore
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    addr = os.environ.get("MASTER_ADDR", "127.0.0.1")
    base_port = int(os.environ.get("MASTER_PORT", "29500"))
    for attempt in range(2):
        env = dict(os.environ)
        env.update(F5_BENCH_WORKER="1", F5_BENCH_ATTEMPT=str(attempt), MASTER_PORT=str(base_port + 20 + attempt))
        env.pop("TORCHELASTIC_USE_AGENT_STORE", None)   # the children rendezvous on their OWN port: rank 0's child hosts that store
        store = TCPStore(addr, base_port + 10 + attempt, world, is_master=(rank == 0), timeout=datetime.timedelta(seconds=900),
                         wait_for_workers=False)
        child = subprocess.Popen([sys.executable, os.path.abspath(__file__)] + sys.argv[1:], env=env, stdout=subprocess.PIPE)
        deadline = time.time() + 1500.0
        while child.poll() is None:
            if store.check(["abort"]) or time.time() > deadline:
                child.kill()
                break
            time.sleep(0.5)
        out = child.stdout.read()
        rc = child.wait()
        ok = rc == 0 and (rank != 0 or out.strip().startswith(b"{"))
        if not ok:
            store.set("abort", "1")
        store.set(f"done{rank}", "1" if ok else "0")
        store.wait([f"done{r}" for r in range(world)])
        all_ok = all(store.get(f"done{r}") == b"1" for r in range(world))
        store.set(f"seen{rank}", "1")           # nobody tears the store down while a peer still reads it
        store.wait([f"seen{r}" for r in range(world)])
        if all_ok:
            if rank == 0:
                os.write(_REAL_STDOUT, out if out.endswith(b"\n") else out + b"\n")
            return 0
        sys.stderr.write(f"bench: attempt {attempt} failed on some rank (rank {rank}: rc={rc}); "
                         f"{'restarting all ranks' if attempt == 0 else 'giving up'}\n")
        del store
        time.sleep(3.0)
    return 1


if __name__ == "__main__":
    _multi = int(os.environ.get("WORLD_SIZE", "1")) > 1
    _worker = os.environ.get("F5_BENCH_WORKER") == "1"
    if _multi and not _worker and "reference" not in sys.argv and os.environ.get("F5_BENCH_NO_SUPERVISOR") is None:
        sys.exit(supervise_ranks())
    if os.environ.get("F5_BENCH_FAKE") is not None:
        _fake_worker()
        sys.exit(0)
    try:
        main()
    except Exception:
        import traceback
        traceback.print_exc()
        # one clean re-start of a single-GPU run (a fresh process and CUDA context); multi-rank runs restart through supervise_ranks
        if not _multi and os.environ.get("F5_BENCH_RETRIED") is None:
            sys.stderr.write("bench: measurement failed, restarting once in a fresh process\n")
            sys.stderr.flush()
            os.environ["F5_BENCH_RETRIED"] = "1"
            os.dup2(_REAL_STDOUT, 1)
            time.sleep(5.0)
            os.execv(sys.executable, [sys.executable] + sys.argv)
        raise
