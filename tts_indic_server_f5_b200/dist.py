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
