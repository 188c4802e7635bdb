"""Data parallelism for the DRN dense-regression path: one process per GPU, full replica per rank, the batch sharded
across ranks, ONE exchange step per iteration -- an averaging all-reduce of the fp32 gradients over NVLink/NVSwitch
(SURVEY.md section 8e).  BatchNorm statistics and the loss normalisers stay per replica, which is what the reference's own
(nominal) multi-GPU mode, nn.DataParallel at main.py:99, computes.

The path (query encoder included) produces all of its gradients in one flat buffer (model/main_model.py:
_run_backward); the backward runs in parts, each followed by the all-reduce of the contiguous region it completed, on a
side stream, so that the exchange of one part runs under the kernels of the next.

The exchange itself is the library's own kernel over peer memory (`drn_p2p_allreduce_avg`, csrc/p2p.cu): the flat buffers of
all ranks are mapped into every process through CUDA IPC (`PeerMemory`) and each rank reduces one slice of the region straight
out of its peers' buffers and writes the mean back into all of them -- one NVLink round instead of a ring's 2 (world - 1)
dependent steps.  NCCL carries the plumbing (rendezvous, handle exchange, parameter broadcast) and remains the transport when
the buffers cannot be mapped (one visible GPU per process, expandable allocator segments, gloo on CPU, DRN_DP_P2P=0).

Two ways to switch it on:
  * `DataParallelDRN(model)` -- explicit wrapper (bench.py, scripts/dp_check.py) over an initialised process group;
  * with the reference's UNCHANGED main.py: start one process per GPU (`torchrun --nproc-per-node 8 main.py ... --gpu
    $LOCAL_RANK`; main.py:52-53 sets CUDA_VISIBLE_DEVICES itself).  `mainModel` reads WORLD_SIZE / RANK at its first CUDA
    forward (`GradReducer.from_env`), creates the NCCL group if the driver did not, broadcasts rank 0's parameters and
    buffers once and installs the same overlapped all-reduce (DRN_AUTO_DP=0 disables the hook).

The reducer is a plain object, NOT an nn.Module: storing a Module on the model it wraps would register it as a submodule
and make the module tree cyclic (`.eval()`, `.state_dict()`, `.to()` would recurse for ever).
"""
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist
from torch import nn

P2P_MAX_RANKS, P2P_FLAG_WORDS = 8, 64


class P2PComm(C.Structure):
    """drn_p2p_t (include/drn_b200.h)."""
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("buf", C.c_void_p * P2P_MAX_RANKS), ("flags", C.c_void_p * P2P_MAX_RANKS)]


class PeerMemory:
    """CUDA-IPC view of the other ranks' device buffers.  `map(t)` is COLLECTIVE (every rank calls it with its own tensor of the
    same role at the same point): it returns the address of that tensor in every rank's memory, as seen from this process,
    or None -- on EVERY rank -- if any rank could not export or open a handle."""

    def __init__(self, group, rank, world):
        from . import lib as L
        self.L, self.lib = L, L.load()
        self.group, self.rank, self.world = group, rank, world
        self.bases = {}  # (rank, handle bytes) -> base address of the peer allocation in this process (opened once)

    def _agree(self, ok):
        flags = [None] * self.world
        dist.all_gather_object(flags, bool(ok), group=self.group)
        return all(flags)

    def map(self, t):
        handle, off = (C.c_ubyte * 64)(), C.c_int64(0)
        rc = self.lib.drn_ipc_export(C.c_void_p(t.data_ptr()), handle, C.byref(off))
        mine = (bytes(handle), int(off.value)) if rc == 0 else None
        err = None if rc == 0 else self.lib.drn_last_error().decode()
        infos = [None] * self.world
        dist.all_gather_object(infos, mine, group=self.group)
        ptrs, ok = [0] * self.world, all(i is not None for i in infos)
        if ok:
            for r, (h, o) in enumerate(infos):
                if r == self.rank:
                    ptrs[r] = t.data_ptr()
                    continue
                base = self.bases.get((r, h))
                if base is None:
                    out = C.c_void_p(0)
                    if self.lib.drn_ipc_open((C.c_ubyte * 64).from_buffer_copy(h), C.c_int64(0), C.byref(out)) != 0:
                        err = self.lib.drn_last_error().decode()
                        ok = False
                        break
                    base = self.bases[(r, h)] = out.value
                ptrs[r] = base + o
        ok = self._agree(ok)
        if not ok and err and self.rank == 0:
            print("drn_b200.parallel: peer-memory mapping unavailable (%s): gradients go through NCCL" % err, file=sys.stderr)
        return ptrs if ok else None


class _P2PWork:
    """Handle of an exchange enqueued on the reducer's side stream: wait() orders the CURRENT stream after it."""

    def __init__(self, event):
        self.event = event

    def wait(self):
        torch.cuda.current_stream().wait_event(self.event)


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
        # peer-memory transport (csrc/p2p.cu): buffers registered by `register`; None = not tried yet, False = unavailable
        self.peers = None
        self.flags = None
        self.flag_ptrs = None
        self.registered = {}  # storage address -> (peer addresses of the storage base, keep-alive tensor)
        self.side = None
        self.transport_used = {"p2p": 0, "nccl": 0}
        self.p2p_ctas = int(os.environ.get("DRN_P2P_CTAS", "0"))

    def _p2p_setup(self, device):
        if self.peers is None:
            self.peers = False
            if (self.world <= P2P_MAX_RANKS and device.type == "cuda" and os.environ.get("DRN_DP_P2P", "1") == "1"
                    and dist.get_backend(self.group) == "nccl"):
                pm = PeerMemory(self.group, self.rank, self.world)
                self.flags = torch.zeros(P2P_FLAG_WORDS, dtype=torch.int32, device=device)
                torch.cuda.synchronize(device)  # the zero fill has landed before any peer may write a flag
                self.flag_ptrs = pm.map(self.flags)
                if self.flag_ptrs is not None:
                    self.peers = pm
                    self.side = torch.cuda.Stream(device=device, priority=-1)
                    dist.barrier(group=self.group)  # every rank's flag block is mapped (and zero) before the first exchange
        return self.peers

    def register(self, flat):
        """COLLECTIVE: map the flat gradient buffer `flat` (same role on every rank) into all ranks, so that `reduce_regions` of
        its slices runs on the peer-memory kernel.  Without it (or when the mapping fails) the slices go through NCCL."""
        if self.world == 1 or flat.device.type != "cuda" or flat.dtype != torch.float32:
            return False
        pm = self._p2p_setup(flat.device)
        if not pm:
            return False
        st = flat.untyped_storage()
        if st.data_ptr() in self.registered:
            return True
        base = torch.empty(0, dtype=torch.float32, device=flat.device).set_(st, 0, (st.nbytes() // 4,), (1,))
        ptrs = pm.map(base)
        if ptrs is None:
            return False
        self.registered[st.data_ptr()] = (ptrs, base)
        return True

    def _p2p_region(self, t):
        """(peer addresses, offset, n) if the slice `t` can go through the peer-memory kernel, else None.  The decision depends
        on shapes and layout only, which are the same on every rank."""
        ent = self.registered.get(t.untyped_storage().data_ptr()) if self.registered else None
        if ent is None or not t.is_contiguous() or t.dtype != torch.float32:
            return None
        off, n = t.storage_offset(), t.numel()
        if (off & 3) or (n & 3):
            return None
        return ent[0], off, n

    def _p2p_allreduce(self, regions):
        """Enqueue the exchange of `regions` (all peer-mapped) on the side stream, ordered after the current stream."""
        from . import lib as L
        lib = L.load()
        cur = torch.cuda.current_stream()
        self.side.wait_stream(cur)
        with torch.cuda.stream(self.side):
            for ptrs, off, n in regions:
                c = P2PComm()
                c.world, c.rank = self.world, self.rank
                for r in range(self.world):
                    c.buf[r], c.flags[r] = ptrs[r], self.flag_ptrs[r]
                L.check(lib.drn_p2p_allreduce_avg(C.byref(c), C.c_int64(off), C.c_int64(n), self.p2p_ctas,
                                                  C.c_void_p(self.side.cuda_stream)), "p2p_allreduce_avg")
            ev = torch.cuda.Event()
            ev.record(self.side)
        self.transport_used["p2p"] += len(regions)
        return _P2PWork(ev)

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
        """Average the given slices of the flat gradient buffer over the ranks: the peer-memory kernel for slices of a
        registered buffer, else an NCCL all-reduce (op = AVG; gloo: SUM then scale).  wait=False returns the pending work
        handles: the exchange runs on a side stream, ordered after what is already enqueued on the current stream, while the
        caller enqueues more work (the next part of the backward)."""
        if self.world == 1:
            return []
        nccl = dist.get_backend(self.group) == "nccl"
        work = []
        regions = [t for t in regions if t.numel() > 0]
        mapped = [self._p2p_region(t) for t in regions] if self.registered else [None] * len(regions)
        if regions and all(m is not None for m in mapped):
            work.append((self._p2p_allreduce(mapped), None))
            regions = []
        for t in regions:
            self.transport_used["nccl"] += 1
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
