"""bench.py's multi-rank supervisor on CPU (fake workers, world size 2 under torchrun): one JSON line on stdout, one attempt
when every rank is healthy (with a real gloo process group formed by the children on their own rendezvous port), a coordinated restart of ALL ranks when one rank's worker dies (its peers would otherwise sit in
a collective for ever), the worker's rendezvous port distinct from the launcher's."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(fake: str, port: int):
    env = dict(os.environ, F5_BENCH_FAKE=fake)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "1", "--warmup", "1"]
    r = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=240)
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    return r, lines


@pytest.mark.parametrize("fake,port,attempt", [("pg", 29711, 0), ("fail=1", 29741, 1), ("fail=0", 29771, 1)])
def test_supervisor_restarts_all_ranks_together(fake, port, attempt):
    r, lines = _run(fake, port)
    assert r.returncode == 0, r.stderr[-2000:]
    assert len(lines) == 1, lines                       # exactly ONE JSON line, from rank 0
    d = json.loads(lines[0])
    assert d["n_gpus"] == 2 and d["attempt"] == attempt
    assert int(d["master_port"]) == port + 20 + attempt  # workers rendezvous on their own port
    if attempt:
        assert "restarting all ranks" in r.stderr
