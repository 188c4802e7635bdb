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
@pytest.mark.parametrize("mode", ["tail_first", "split", "overlap", "r01", "auto", "nccl"])
def test_dp_gradients_equal_mean_of_single_rank_gradients(mode):
    """Every schedule on the library's peer-memory all-reduce (the transport is asserted), and once on NCCL (DRN_DP_P2P=0)."""
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    if mode == "nccl":
        env["DRN_DP_P2P"], env["DRN_EXPECT_TRANSPORT"] = "0", "nccl"
    else:
        env["DRN_DP_P2P"], env["DRN_EXPECT_TRANSPORT"] = "1", "p2p"
    if mode not in ("auto", "nccl"):
        env["DRN_DP_ORDER"] = mode
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_port()), os.path.join(REPO, "scripts", "dp_check.py")] + (["--auto"] if mode == "auto" else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=REPO)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "dp_check" in r.stdout and "max rel-L2" in r.stdout, r.stdout[-2000:]


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_peer_memory_allreduce_matches_nccl():
    """drn_p2p_allreduce_avg on ragged regions, 200 back-to-back calls, bit-identical results on every rank (scripts/p2p_check.py)."""
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_port()), os.path.join(REPO, "scripts", "p2p_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=REPO)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "p2p_check" in r.stdout, r.stdout[-2000:]
