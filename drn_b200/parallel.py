"""Data parallelism for the DRN dense-regression path: one process per GPU, full replica per rank, the batch sharded
across ranks, ONE exchange step per iteration -- an NCCL all-reduce (sum, then 1/world) of the fp32 gradients over
NVLink/NVSwitch.  BatchNorm statistics and the loss normalisers stay per replica, which is what the
reference's own (nominal) multi-GPU mode, nn.DataParallel at main.py:99, computes.

The path (query encoder included) produces all of its gradients in one flat buffer (model/main_model.py:
_run_backward).  The backward runs in three parts, each followed by the all-reduce of the contiguous region it completed, on
NCCL's stream: head / FPN / backbone gradients are reduced WHILE the prop_fc weight gradient (0.53 ms) runs, prop_fc.weight's
WHILE the tail (gates, query encoder: ~0.35 ms) runs; the tail's region follows.  Gradients
that autograd produced outside that buffer (none for the reference model; kept for wrapped modules that add their own
parameters) are all-reduced in `finish_gradient_sync()` after `loss.backward()`.
"""
import torch
import torch.distributed as dist
from torch import nn


class DataParallelDRN(nn.Module):
    def __init__(self, module, process_group=None):
        super().__init__()
        self.module = module
        self.group = process_group
        self.world = dist.get_world_size(process_group)
        # replica 0's parameters and buffers win, as with nn.DataParallel
        with torch.no_grad():
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t, 0, group=process_group)
        module._dp = self

    def reduce_regions(self, regions, wait=True):
        """Average the given slices of the flat gradient buffer over the ranks (NCCL all-reduce, op = AVG; gloo: SUM then
        scale).  wait=False returns the pending work handles: the collective runs on the backend's stream, ordered after
        what is already enqueued on the current stream, while the caller enqueues more work (the backward tail)."""
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
        self.reduce_regions([flat])
        o = 0
        for g in grads:
            g.copy_(flat[o:o + g.numel()].view_as(g))
            o += g.numel()
