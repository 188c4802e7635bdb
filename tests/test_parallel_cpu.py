"""Host-side logic of the data-parallel wrapper on CPU: 2 ranks, gloo, 127.0.0.1.  Checks the rank-0 broadcast of
parameters/buffers, the averaged all-reduce of the dense path's flat gradient buffer and of gradients produced outside that buffer."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _Toy(nn.Module):
    def __init__(self, rank):
        super().__init__()
        self.query_encoder = nn.Linear(4, 3)
        self.prop_fc = nn.Linear(3, 2)
        self.register_buffer("running", torch.full((2,), float(rank)))
        with torch.no_grad():
            for p in self.parameters():
                p.fill_(float(rank) + 1.0)
        self._dp = None
        self._trainable_names = ["prop_fc.weight", "prop_fc.bias"]  # produced in the flat buffer of the path


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from drn_b200.parallel import DataParallelDRN
    toy = _Toy(rank)
    dp = DataParallelDRN(toy)
    ok = all(bool((p == 1.0).all()) for p in toy.parameters()) and bool((toy.running == 0.0).all())  # replica 0 wins
    flat = torch.arange(6, dtype=torch.float32) * (rank + 1)
    ok = ok and toy._dp is dp.reducer and "_dp" not in toy._modules  # a plain object: no cycle in the module tree
    # with the wrapper registered as a submodule of the model it wraps these recursed for ever
    from drn_b200.checkpoint import reference_checkpoint
    dp.eval()
    dp.train()
    toy.eval()
    toy.train()
    ok = ok and sorted(dp.state_dict()) == sorted("module." + k for k in toy.state_dict())
    ok = ok and sorted(reference_checkpoint(dp)["state_dict"]) == sorted(dp.state_dict())
    dp.to("cpu")
    work = dp.reduce_regions([flat[:4]], wait=False)  # what _run_backward does: first region async, tail regions, then wait
    dp.reduce_regions([flat[4:], flat[:0]])
    dp.wait(work)
    ok = ok and torch.allclose(flat, torch.arange(6, dtype=torch.float32) * 1.5)
    # the peer-memory transport needs CUDA buffers and NCCL for the handle exchange: on CPU / gloo registration declines on every
    # rank and the regions keep going through the process group
    ok = ok and dp.reducer.register(flat) is False and dp.reducer.transport_used == {"p2p": 0, "nccl": 2}
    for p in toy.query_encoder.parameters():
        p.grad = torch.full_like(p, float(rank))
    toy.prop_fc.weight.grad = torch.full_like(toy.prop_fc.weight, 7.0)  # dense grads are NOT touched by finish_gradient_sync
    dp.finish_gradient_sync()
    ok = ok and all(torch.allclose(p.grad, torch.full_like(p, 0.5)) for p in toy.query_encoder.parameters())
    ok = ok and bool((toy.prop_fc.weight.grad == 7.0).all())
    q.put((rank, ok))
    dist.destroy_process_group()


def test_data_parallel_wrapper_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def _auto_worker(rank, world, port, q):
    """The `torchrun main.py` route: WORLD_SIZE in the environment, no process group yet -> auto_reducer creates it (gloo on a
    CPU device), broadcasts rank 0's tensors and returns the reducer; with WORLD_SIZE=1 or DRN_AUTO_DP=0 it returns None."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from drn_b200.parallel import auto_reducer
    toy = _Toy(rank)
    os.environ["DRN_AUTO_DP"] = "0"
    ok = auto_reducer(toy, torch.device("cpu")) is None and not dist.is_initialized()
    os.environ["DRN_AUTO_DP"] = "1"
    red = auto_reducer(toy, torch.device("cpu"))
    ok = ok and red is not None and red.world == world and dist.is_initialized()
    ok = ok and all(bool((p == 1.0).all()) for p in toy.parameters()) and bool((toy.running == 0.0).all())
    flat = torch.full((5,), float(rank + 1))
    red.reduce_regions([flat])
    ok = ok and torch.allclose(flat, torch.full((5,), 1.5))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_auto_reducer_from_environment_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_auto_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
