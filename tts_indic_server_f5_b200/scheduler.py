"""Request scheduler: length-bucketed batching of concurrent synthesis requests.

The reference serves one request at a time and, inside a request, one text chunk at a time
(`f5_tts/infer/utils_infer.py:441-466`, B = 1 per `sample` call).  Chunks of one text — and chunks of different
requests — are independent until the final cross-fade (`:485-519`), so this scheduler flattens every pending request
into utterances, packs them into row-budgeted batches for the packed var-len engine and runs each pack through
`Synthesizer.generate`; results are re-assembled per request in submission order with the reference's cross-fade.

Packing (`plan_packs`): utterances sorted by length (longest first — neighbours in a pack have similar lengths, which
keeps the persistent attention grid's tail short), first-fit into packs bounded by
  * `max_rows`  — rows of the packed activation matrix INCLUDING the CFG duplicate and the 16-row gaps
                  (~30 KB of workspace per row, so 131 072 rows ~ 4 GB), and
  * `max_utts`  — utterances per pack.
A pack is one CUDA-graph replay per distinct length signature; nothing here touches the device.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch

from . import text as T
from .layout import GAP, ROW_ALIGN
from .synthetic import UtteranceSpec

TARGET_SR = 24000


def pack_rows(lengths: list[int]) -> int:
    """Rows of the packed matrix for these utterance lengths, both CFG halves (layout.build_layout)."""
    r = GAP + sum(n + GAP for n in lengths)
    return 2 * ((r + ROW_ALIGN - 1) // ROW_ALIGN * ROW_ALIGN)


def plan_packs(lengths: list[int], max_rows: int = 131072, max_utts: int = 256) -> list[list[int]]:
    """Indices of `lengths` grouped into packs; every index appears exactly once; an utterance that alone exceeds
    `max_rows` still gets its own pack (the engine's 4096-frame clamp bounds it)."""
    order = sorted(range(len(lengths)), key=lambda i: (-lengths[i], i))
    packs: list[list[int]] = []
    for i in order:
        for p in packs:                                   # first fit
            if len(p) < max_utts and pack_rows([lengths[j] for j in p] + [lengths[i]]) <= max_rows:
                p.append(i)
                break
        else:
            packs.append([i])
    return packs


def cross_fade(waves: list[np.ndarray], cross_fade_duration: float = 0.15, sample_rate: int = TARGET_SR) -> np.ndarray:
    """utils_infer.py:485-519: linear cross-fade of `cross_fade_duration` seconds between consecutive chunk waves."""
    if not waves:
        return np.zeros(0, dtype=np.float32)
    if cross_fade_duration <= 0:
        return np.concatenate(waves)
    final = waves[0]
    for nxt in waves[1:]:
        n = min(int(cross_fade_duration * sample_rate), len(final), len(nxt))
        if n <= 0:
            final = np.concatenate([final, nxt])
            continue
        mixed = final[-n:] * np.linspace(1, 0, n) + nxt[:n] * np.linspace(0, 1, n)
        final = np.concatenate([final[:-n], mixed, nxt[n:]])
    return final


@dataclass
class Request:
    """One `infer_process` call waiting to run (utils_infer.py:357-400 arguments)."""
    rid: int
    audio: torch.Tensor            # [ch, nw] fp32
    sr: int
    ref_text: str
    gen_text: str
    speed: float = 1.0
    fix_duration: float | None = None
    cross_fade_duration: float = 0.15
    chunks: list[str] = field(default_factory=list)
    seed: int = 0                  # base key of this request's noise (chunk i draws with utterance_seed(seed, i))


class RequestScheduler:
    """submit() any number of requests, then run() — every chunk of every request goes through the engine in packed
    batches.  Same results as calling `infer_process` per request (packing is bit-invariant, tests/test_parity_gpu.py)."""

    def __init__(self, synthesizer, max_rows: int = 131072, max_utts: int = 256, nfe_step: int = 32, cfg_strength: float = 2.0,
                 sway_sampling_coef: float = -1.0):
        self.syn = synthesizer
        self.max_rows, self.max_utts = max_rows, max_utts
        self.nfe_step, self.cfg_strength, self.sway = nfe_step, cfg_strength, sway_sampling_coef
        self.pending: list[Request] = []
        self._next = 0
        self.last_packs: list[list[int]] = []

    def submit(self, ref_audio, ref_text: str, gen_text: str, speed: float = 1.0, fix_duration: float | None = None,
               cross_fade_duration: float = 0.15, seed: int | None = None) -> int:
        """Queue one `infer_process`-shaped request.  `seed` fixes its noise (same value => same audio as
        `infer_process(..., seed=seed)`); by default every request is a fresh draw like the reference's (cfm.py:186)."""
        from .api import fresh_noise_seed
        audio, sr = ref_audio
        max_chars = int(len(ref_text.encode("utf-8")) / (audio.shape[-1] / sr) * (25 - audio.shape[-1] / sr))   # utils_infer.py:377
        req = Request(self._next, audio, sr, ref_text, gen_text, speed, fix_duration, cross_fade_duration,
                      T.chunk_text(gen_text, max_chars=max_chars), fresh_noise_seed() if seed is None else seed)
        self._next += 1
        self.pending.append(req)
        return req.rid

    def run(self) -> dict[int, tuple[np.ndarray, int, np.ndarray]]:
        """-> {request id: (wave fp32, 24000, mel [100, F])}, the reference's `infer_process` triple per request."""
        from .api import utterance_seed
        reqs, self.pending = self.pending, []
        specs, owner = [], []
        for r in reqs:
            for ci, chunk in enumerate(r.chunks):
                specs.append(UtteranceSpec(audio=r.audio, ref_text=r.ref_text, gen_text=chunk, duration=None, noise_index=ci,
                                           meta={"sr": r.sr, "speed": r.speed, "fix_duration": r.fix_duration,
                                                 "noise_seed": utterance_seed(r.seed, ci)}))
                owner.append((r.rid, ci))
        if not specs:
            return {}
        lengths = [min(self.syn._prep(s, s.meta["speed"], s.meta["fix_duration"]).duration, 4096) for s in specs]
        self.last_packs = plan_packs(lengths, self.max_rows, self.max_utts)
        waves: list = [None] * len(specs)
        mels: list = [None] * len(specs)
        for pack in self.last_packs:
            # speed / fix_duration are per request: one generate() call per distinct setting inside the pack
            groups: dict[tuple, list[int]] = {}
            for i in pack:
                groups.setdefault((specs[i].meta["speed"], specs[i].meta["fix_duration"]), []).append(i)
            for (speed, fixd), idx in groups.items():
                w, m = self.syn.generate([specs[i] for i in idx], self.nfe_step, self.cfg_strength, self.sway, speed, fixd,
                                         return_mel=True)
                for i, wi, mi in zip(idx, w, m):
                    waves[i], mels[i] = wi, mi
        out = {}
        for r in reqs:
            idx = [k for k, (rid, _) in enumerate(owner) if rid == r.rid]
            out[r.rid] = (cross_fade([waves[k] for k in idx], r.cross_fade_duration), TARGET_SR,
                          np.concatenate([mels[k] for k in idx], axis=1))
        return out


class ContinuousScheduler:
    """The serving loop the reference lacks (its route runs the model on the event-loop thread, one request at a time:
    src/server/routes/speech.py:19-41).  Requests arrive from any thread through `submit()` (a bounded queue: back-pressure
    instead of unbounded memory), ONE worker thread owns the GPU: it drains whatever is waiting — after the first request it
    lingers `max_wait_ms` for company, up to `max_batch_requests` — and runs the drained set through `RequestScheduler` (chunks
    of all requests length-bucketed into packed batches, per-request cross-fade).  Callers get a `concurrent.futures.Future`
    resolving to the reference's `infer_process` triple (wave fp32, 24000, mel [100, F]).  An `async` route awaits
    `asyncio.wrap_future(scheduler.submit(...))` and the event loop stays free."""

    def __init__(self, synthesizer, max_queue: int = 256, max_batch_requests: int = 64, max_wait_ms: float = 4.0, **sched_kw):
        import queue
        import threading
        self._sched = RequestScheduler(synthesizer, **sched_kw)
        self._q: "queue.Queue" = queue.Queue(maxsize=max_queue)
        self.max_batch_requests, self.max_wait_ms = max_batch_requests, max_wait_ms
        self.batches: list[int] = []                 # requests per executed batch (observability / tests)
        self._stop = threading.Event()
        self._worker = threading.Thread(target=self._loop, name="f5-scheduler", daemon=True)
        self._worker.start()

    def submit(self, ref_audio, ref_text: str, gen_text: str, *, timeout: float | None = None, **kw):
        """Queue a request; raises `queue.Full` after `timeout` seconds if the queue stays full (HTTP 503 material)."""
        from concurrent.futures import Future
        if self._stop.is_set():
            raise RuntimeError("scheduler is closed")
        fut: Future = Future()
        self._q.put((fut, ref_audio, ref_text, gen_text, kw), timeout=timeout)
        return fut

    def _loop(self) -> None:
        import queue
        import time
        dev = getattr(self._sched.syn, "device", None)
        if dev is not None and getattr(dev, "type", None) == "cuda":
            torch.cuda.set_device(dev)               # the current device is per thread: this worker owns the engine's GPU
        while not self._stop.is_set():
            try:
                first = self._q.get(timeout=0.05)
            except queue.Empty:
                continue
            items = [first]
            deadline = time.monotonic() + self.max_wait_ms * 1e-3
            while len(items) < self.max_batch_requests:
                left = deadline - time.monotonic()
                try:
                    items.append(self._q.get(timeout=left) if left > 0 else self._q.get_nowait())
                except queue.Empty:
                    break
            rids = []
            try:
                for fut, ref_audio, ref_text, gen_text, kw in items:
                    rids.append(self._sched.submit(ref_audio, ref_text, gen_text, **kw))
                out = self._sched.run()
                self.batches.append(len(items))
                for (fut, *_), rid in zip(items, rids):
                    fut.set_result(out[rid])
            except BaseException as e:  # noqa: BLE001 — every waiter must hear about it (a CUDA error is sticky)
                self._sched.pending = []
                for fut, *_ in items:
                    if not fut.done():
                        fut.set_exception(e)

    def close(self, timeout: float = 5.0) -> None:
        self._stop.set()
        self._worker.join(timeout)
        import queue
        while True:                                   # nobody is left waiting on a dead worker
            try:
                fut, *_ = self._q.get_nowait()
            except queue.Empty:
                break
            fut.set_exception(RuntimeError("scheduler closed"))
