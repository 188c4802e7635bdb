"""Data parallelism for the DRN dense-regression path: one process per GPU, full replica per rank, the batch sharded
across ranks, ONE exchange step per iteration -- an NCCL all-reduce (sum, then 1/world) of the fp32 gradients over
NVLink/NVSwitch.  BatchNorm statistics and the loss normalisers stay per replica, which is what the
reference's own (nominal) multi-GPU mode, nn.DataParallel at main.py:99, computes.

The path (query encoder included) produces all of its gradients in one flat buffer (model/main_model.py:
_run_backward); the backward runs in parts, each followed by the all-reduce of the contiguous region it completed, on
NCCL's stream, so that the collective of one part runs under the kernels of the next.

Two ways to switch it on:
  * `DataParallelDRN(model)` -- explicit wrapper (bench.py, scripts/dp_check.py) over an initialised process group;
  * with the reference's UNCHANGED main.py: start one process per GPU (`torchrun --nproc-per-node 8 main.py ... --gpu
    $LOCAL_RANK`; main.py:52-53 sets CUDA_VISIBLE_DEVICES itself).  `mainModel` reads WORLD_SIZE / RANK at its first CUDA
    forward (`GradReducer.from_env`), creates the NCCL group if the driver did not, broadcasts rank 0's parameters and
    buffers once and installs the same overlapped all-reduce (DRN_AUTO_DP=0 disables the hook).

The reducer is a plain object, NOT an nn.Module: storing a Module on the model it wraps would register it as a submodule
and make the module tree cyclic (`.eval()`, `.state_dict()`, `.to()` would recurse for ever).
"""
import os

import torch
import torch.distributed as dist
from torch import nn


def nccl_env_defaults():
    """Call BEFORE the NCCL communicator is created.  DRN_NCCL_MAX_CTAS=n caps NCCL's CTAs (A/B knob).  Measured on 8 x B200
    (profiles/r02_dp_timeline8_*.json): with 16 CTAs the collectives no longer stretch the kernels they run beside (the
    cooperative LSTM launches of the tail, the 64-cluster chunks of the prop_fc weight gradient) but lose a third of their
    bandwidth, with 8 more than half -- a net loss, so NCCL's default stays."""
    if os.environ.get("DRN_NCCL_MAX_CTAS"):
        os.environ.setdefault("NCCL_MAX_CTAS", os.environ["DRN_NCCL_MAX_CTAS"])


class GradReducer:
    """Averaging all-reduce of regions of the flat gradient buffer over the ranks of a process group."""

    def __init__(self, process_group=None):
        self.group = process_group
        self.world = dist.get_world_size(process_group)
        self.rank = dist.get_rank(process_group)

    @classmethod
    def from_env(cls, module, device):
        """The hook behind `torchrun main.py`: WORLD_SIZE > 1 in the environment.  Creates the default process group when the
        driver has none (NCCL on a CUDA device, env:// rendezvous: MASTER_ADDR / MASTER_PORT / RANK / WORLD_SIZE as torchrun
        exports them) and makes every rank start from rank 0's parameters and buffers."""
        if not dist.is_initialized():
            if device.type == "cuda":
                nccl_env_defaults()
                dist.init_process_group("nccl", device_id=device)
            else:
                dist.init_process_group("gloo")
        r = cls(None)
        r.broadcast_module(module)
        return r

    def broadcast_module(self, module):
        """Replica 0's parameters and buffers win, as with nn.DataParallel (which re-replicates them every step)."""
        with torch.no_grad():
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t, 0, group=self.group)

    def reduce_regions(self, regions, wait=True):
        """Average the given slices of the flat gradient buffer over the ranks (NCCL all-reduce, op = AVG; gloo: SUM then
        scale).  wait=False returns the pending work handles: the collective runs on the backend's stream, ordered after
        what is already enqueued on the current stream, while the caller enqueues more work (the next part of the backward)."""
        if self.world == 1:
            return []
        nccl = dist.get_backend(self.group) == "nccl"
        work = []
        for t in regions:
            if t.numel() == 0:
                continue
            if nccl:
                work.append((dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group, async_op=True), None))
            else:
                work.append((dist.all_reduce(t, group=self.group, async_op=True), t))
        if wait:
            self.wait(work)
            return []
        return work

    def wait(self, work):
        for w, t in work:
            w.wait()
            if t is not None:
                t.mul_(1.0 / self.world)


class DataParallelDRN(nn.Module):
    def __init__(self, module, process_group=None):
        super().__init__()
        self.module = module
        self.reducer = GradReducer(process_group)
        self.group, self.world = process_group, self.reducer.world
        self.reducer.broadcast_module(module)
        module._dp = self.reducer  # plain object: nothing is registered on the wrapped module

    def reduce_regions(self, regions, wait=True):
        return self.reducer.reduce_regions(regions, wait)

    def wait(self, work):
        self.reducer.wait(work)

    def forward(self, *a, **k):
        return self.module(*a, **k)

    def finish_gradient_sync(self):
        """All-reduce the gradients autograd produced outside the flat buffer of the path."""
        if self.world == 1:
            return
        inside = set(getattr(self.module, "_trainable_names", ()))
        grads = [p.grad for n, p in self.module.named_parameters() if n not in inside and p.grad is not None]
        if not grads:
            return
        flat = torch.cat([g.reshape(-1) for g in grads])
        self.reducer.reduce_regions([flat])
        o = 0
        for g in grads:
            g.copy_(flat[o:o + g.numel()].view_as(g))
            o += g.numel()


def auto_reducer(module, device):
    """mainModel's construction-time contract of SURVEY.md section 8e, evaluated at the first forward (the device is known
    then): WORLD_SIZE > 1 and DRN_AUTO_DP != 0 -> a GradReducer over the default group; otherwise None."""
    if int(os.environ.get("WORLD_SIZE", "1")) <= 1 or os.environ.get("DRN_AUTO_DP", "1") != "1":
        return None
    return GradReducer.from_env(module, device)
