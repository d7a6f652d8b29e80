"""CPU, world_size 2, gloo: the N > 1 path — utterance sharding + the end-of-batch gather of waveforms."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tts_indic_server_f5_b200.dist import gather_waveforms, lpt_partition


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, lens, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = lpt_partition(lens, world)[rank]
    samples = [256 * (lens[i] - 1) for i in mine]
    wav = torch.cat([torch.full((n,), float(i)) for i, n in zip(mine, samples)]) if mine else torch.zeros(0)
    got = gather_waveforms(wav, samples, dst=0)
    if rank == 0:
        out = {}
        for r, (flat, ls) in enumerate(got):
            idx, off = lpt_partition(lens, world)[r], 0
            for i, n in zip(idx, ls):
                out[i] = (n, float(flat[off]), float(flat[off + n - 1]))
                off += n
            assert off == flat.numel()
        q.put(out)
    else:
        assert got is None
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    lens = [50, 7, 33, 90, 12]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lens, q)) for r in range(2)]
    [p.start() for p in procs]
    out = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert sorted(out) == list(range(len(lens)))
    for i, n in enumerate(lens):
        assert out[i] == (256 * (n - 1), float(i), float(i))
