"""Multi-GPU: independent utterances shard across ranks (one process per GPU); no collective runs inside the sampling
loop.  NCCL is used once per request batch to gather waveforms on rank 0.  The reference has no
multi-GPU inference at all (single process, `cuda:0`, `src/server/utils/device_utils.py:7-8`)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def utterance_cost(n: int) -> float:
    """Algorithmic FLOPs of one DiT forward for an n-frame utterance."""
    return 387.305e6 * n + 90112.0 * n * n


def lpt_partition(lengths: list[int], world_size: int) -> list[list[int]]:
    """Longest-processing-time-first assignment of utterance indices to ranks."""
    order = sorted(range(len(lengths)), key=lambda i: -utterance_cost(lengths[i]))
    loads = [0.0] * world_size
    parts: list[list[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: loads[k])
        parts[r].append(i)
        loads[r] += utterance_cost(lengths[i])
    for p in parts:
        p.sort()
    return parts


def gather_waveforms(wav: torch.Tensor, lengths: list[int], dst: int = 0):
    """Gather each rank's flat waveform buffer (+ per-utterance sample counts) on `dst`.
    Returns on dst: list over ranks of (flat tensor, lengths list); elsewhere None.  Works with nccl (CUDA tensors) and
    gloo (CPU tensors, used by the world_size-2 CPU tests)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = wav.device
    meta = torch.tensor([wav.numel(), len(lengths)], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta)
    max_n, max_u = int(max(m[0] for m in metas)), int(max(m[1] for m in metas))
    lens_t = torch.zeros(max_u, dtype=torch.int64, device=dev)
    lens_t[: len(lengths)] = torch.tensor(lengths, dtype=torch.int64, device=dev)
    padded = torch.zeros(max_n, dtype=wav.dtype, device=dev)
    padded[: wav.numel()] = wav
    if rank == dst:
        bufs = [torch.zeros_like(padded) for _ in range(world)]
        lbufs = [torch.zeros_like(lens_t) for _ in range(world)]
    else:
        bufs = lbufs = None
    dist.gather(padded, bufs, dst=dst)
    dist.gather(lens_t, lbufs, dst=dst)
    if rank != dst:
        return None
    return [(b[: int(m[0])], lb[: int(m[1])].tolist()) for b, lb, m in zip(bufs, lbufs, metas)]


def shard_plan(lengths: list[int], world_size: int, max_rows: int = 180224, max_utts: int = 256) -> dict:
    """The whole sharding decision for ONE request batch, computed identically on every rank from the full list of utterance
    lengths (no communication): LPT partition over ranks by `utterance_cost`, then first-fit packs under a row budget per
    rank (scheduler.plan_packs).  Returns parts (utterance indices per rank), packs (per rank: lists of LOCAL positions into
    parts[r]) and the LPT imbalance max(load) / mean(load)."""
    from .scheduler import plan_packs
    parts = lpt_partition(lengths, world_size)
    loads = [sum(utterance_cost(lengths[i]) for i in p) for p in parts]
    packs = [plan_packs([lengths[i] for i in p], max_rows, max_utts) if p else [] for p in parts]
    mean = sum(loads) / max(len(loads), 1)
    return {"parts": parts, "packs": packs, "loads": loads, "imbalance": (max(loads) / mean) if mean > 0 else 1.0}


class WaveGatherer:
    """End-of-batch gather of every rank's flat waveform buffer on `dst` with PERSISTENT buffers: one padded send buffer per
    rank, one [world, max_samples] receive buffer and one pinned host landing buffer on `dst`.  The sample counts of every
    rank are known everywhere from the shard plan, so no counts are exchanged.  Round 1 allocated (and pinned) ~400 MB afresh
    on rank 0 every step, which was the serial tail of the 8-GPU end-to-end number."""

    def __init__(self, samples_per_rank: list[int], device, dtype=torch.float32, dst: int = 0):
        self.world, self.rank, self.dst = dist.get_world_size(), dist.get_rank(), dst
        assert len(samples_per_rank) == self.world
        self.samples = list(samples_per_rank)
        self.max = max(max(self.samples), 1)
        self.send = torch.zeros(self.max, dtype=dtype, device=device)
        self.recv = self.host = None
        if self.rank == dst:
            self.recv = torch.zeros(self.world, self.max, dtype=dtype, device=device)
            self.host = torch.zeros(self.world, self.max, dtype=dtype)
            if torch.device(device).type == "cuda":
                self.host = self.host.pin_memory()

    def gather(self, wav: torch.Tensor):
        """wav: this rank's flat buffer (>= samples[rank] elements).  Returns the [world, max] device buffer on dst."""
        n = self.samples[self.rank]
        self.send[:n].copy_(wav[:n])
        dist.gather(self.send, [self.recv[r] for r in range(self.world)] if self.rank == self.dst else None, dst=self.dst)
        return self.recv

    def to_host(self):
        """dst only: D2H of the gathered rows (used prefixes only) into the pinned landing buffer -> list of numpy views."""
        if self.rank != self.dst:
            return None
        for r, n in enumerate(self.samples):
            self.host[r, :n].copy_(self.recv[r, :n], non_blocking=True)
        if self.recv.is_cuda:
            torch.cuda.current_stream().synchronize()
        return [self.host[r, :n].numpy() for r, n in enumerate(self.samples)]
