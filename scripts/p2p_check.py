"""Peer-memory all-reduce (drn_p2p_allreduce_avg, csrc/p2p.cu) against NCCL on real GPUs -- run under torchrun, >= 2 ranks:
ragged regions (slice sizes not divisible by the world size, a few floats, tens of MB), 200 back-to-back calls (epoch / flag
reuse), results bit-identical on every rank, and the transfer rate of a gradient-sized region.  Prints one line on rank 0."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from drn_b200.parallel import GradReducer  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
red = GradReducer(None)
N = 40 * 1024 * 1024  # 160 MB of fp32: the size of the path's flat gradient buffer
flat = torch.zeros(N, device=dev)
assert red.register(flat), "peer-memory mapping failed"
g = torch.Generator(device=dev)
g.manual_seed(1234 + rank)
worst, nbit = 0.0, 0
cases = [(0, 4), (4, 8), (8, 4 * 1000 + 4), (4096, 4 * 999983), (0, N), (N - 12, 12), (1 << 20, 21 * 1024 * 1024)]
for off, n in cases:
    flat.normal_(generator=g)
    ref = flat[off:off + n].clone()
    dist.all_reduce(ref, op=dist.ReduceOp.SUM)
    ref /= world
    before = flat.clone()
    red.reduce_regions([flat[off:off + n]])
    torch.cuda.synchronize()
    err = float((flat[off:off + n] - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
    worst = max(worst, err)
    # outside the region nothing changed
    assert torch.equal(flat[:off], before[:off]) and torch.equal(flat[off + n:], before[off + n:]), "wrote outside the region"
    # every rank holds the same bits
    mine = flat[off:off + n].clone()
    other = mine.clone()
    dist.broadcast(other, 0)
    nbit += int(not torch.equal(mine, other))
assert red.transport_used["p2p"] == len(cases) and red.transport_used["nccl"] == 0, red.transport_used
# many calls back to back, alternating regions, no host synchronisation in between
flat.fill_(float(rank + 1))
for i in range(200):
    red.reduce_regions([flat[(i % 7) * 4096:(i % 7) * 4096 + 4096 * 3]], wait=True)
torch.cuda.synchronize()
mean = (world + 1) / 2.0
assert float((flat[:4096 * 9] - mean).abs().max()) < 1e-5, "repeated calls diverged"
# transfer rate of the whole buffer (algorithmic bytes = 4 N per rank; NVLink moves 2 * (world-1)/world of that per GPU)
sweep = {}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for ctas in (10, 40, 80, 0):  # 0 = the library default (the last one: `ms` below)
    red.p2p_ctas = ctas
    for n_mb in (16, 160):
        reg = flat[:n_mb * 262144]
        red.reduce_regions([reg])
        dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            red.reduce_regions([reg])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        sweep["%d ctas, %d MB" % (ctas, n_mb)] = round(ms, 4)
ref = flat.clone()
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    dist.all_reduce(ref, op=dist.ReduceOp.AVG)
e1.record()
torch.cuda.synchronize()
ms_nccl = e0.elapsed_time(e1) / 10
t = torch.tensor([worst, float(nbit), ms, ms_nccl], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("p2p_check world=%d max rel err vs NCCL %.2e, ranks with different bits %d, %d MB all-reduce: peer-memory kernel %.3f ms "
          "(%.0f GB/s algorithmic), NCCL %.3f ms" % (world, float(t[0]), int(t[1]), 4 * N >> 20, float(t[2]), 4 * N / float(t[2]) / 1e6, float(t[3])))
    print("p2p_check sweep (ms per all-reduce, rank 0):", sweep)
assert float(t[0]) < 1e-6 and int(t[1]) == 0
dist.destroy_process_group()
