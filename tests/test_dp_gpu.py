"""Data-parallel correctness on real GPUs (runs when >= 2 are visible): scripts/dp_check.py under torchrun -- the gradients
left in param.grad by the overlapped schedule (graph 1, tail, chunked prop_fc weight gradient, all-reduces in between) must
equal the mean over ranks of the single-rank gradients of each rank's shard (SURVEY.md 8e parity oracle), for the explicit
wrapper, for the round-1 order, and for the WORLD_SIZE hook that `torchrun main.py` relies on."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("mode", ["tail_first", "split", "overlap", "r01", "auto"])
def test_dp_gradients_equal_mean_of_single_rank_gradients(mode):
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    if mode != "auto":
        env["DRN_DP_ORDER"] = mode
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_port()), os.path.join(REPO, "scripts", "dp_check.py")] + (["--auto"] if mode == "auto" else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=REPO)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "dp_check" in r.stdout and "max rel-L2" in r.stdout, r.stdout[-2000:]
