"""`mainModel`: the drop-in boundary of the DRN dense-regression path (reference model/main_model.py:13-81).

Same constructor signature, attribute tree, state_dict keys/shapes and `forward(query_tokens, query_length,
props_features, props_start_end, gt_start_end, props_num, num_frames) -> (box_lists | None, loss_dict)` as the reference,
so the reference's main.py (19, 89-99, 124-140, 218-243) runs against it unchanged.  Everything, query encoder included,
is executed by hand-written sm_100a kernels (drn_b200.dense.DensePath -> libdrn_sm100.so); gradients arrive in
`param.grad` through one autograd.Function whose backward is the hand-derived backward schedule.  There is no CPU or
torch-op fallback: calling forward without a B200 raises.
"""
import os

import torch
import torch.nn as nn

from drn_b200.dense import DensePath
from model.inference import assemble
from model.language_module import QueryEncoder
from model.modules import FPN, Backbone, FCOSModule


# Optional per-phase CUDA-event trace of the data-parallel backward (scripts/dp_timeline.py sets it to a list): (name, event)
# pairs recorded on the compute stream; `wait:` entries are recorded right after the stream was made to wait for a collective,
# i.e. they carry the time at which that all-reduce had finished.
DP_TRACE = None


def _mark(name):
    if DP_TRACE is not None:
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        DP_TRACE.append((name, e))


# NVTX ranges around the phases of a step (tracing row): DRN_NVTX=1.  Under CUDA-graph replay a range
# brackets the host-side enqueue of the phase; with DRN_NO_GRAPHS=1 every kernel launch of the phase falls inside it.
_NVTX = os.environ.get("DRN_NVTX", "0") == "1"


class _nvtx:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _NVTX:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *a):
        if _NVTX:
            torch.cuda.nvtx.range_pop()


def _sig(p):
    return hash(tuple(t.data_ptr() for t in p.values()))


def _replay(path, key, fn, use_graphs):
    """First call for a (part, mode, parameter-storage) signature runs `fn` eagerly and records a CUDA graph of the same launch
    sequence; later calls replay it (~50 kernel launches -> one graph launch)."""
    if not use_graphs:
        fn()
        return
    g = path.graphs.get(key)
    if g is None:
        fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):  # a DataLoader's pin-memory thread may call CUDA meanwhile
            fn()
        path.graphs[key] = g
    else:
        g.replay()


def _run_forward(path, p, training, use_graphs, tokens, lengths, feats, pse, gt):
    """Forward schedule: the eager staging of the caller's tensors, then ONE replayable graph (query encoder + gates with the
    weight packing as a parallel branch under the recurrence, then prop_fc ... losses)."""
    with _nvtx("drn.stage_inputs"):
        path.stage_inputs(p, tokens, lengths, feats, pse, gt)

    def core():
        path.forward_pre(p)
        path.forward_main(p, training)
    with _nvtx("drn.forward"):
        _replay(path, ("fwd", training, _sig(p)), core, use_graphs)


def _run_backward(path, p, names, upstream, use_graphs, dp=None, flat_cache=None):
    """Backward of the path into ONE flat gradient buffer (one per model and gradient layout, shared by all (B, T, L) shapes --
    `flat_cache` -- so that in a data-parallel run rank r's gradients always sit in the buffer its peers have mapped, whichever
    shape each rank happens to run).  Layout (1 = head + FPN, 2 = backbone):
        [C1: stored][B1: accumulated][B2: accumulated][A: accumulated, produced by the tail (gates, query encoder)][C2: stored]
        [D: prop_fc.weight]
    B1+B2+A are zero-filled every backward (atomics / += land there); C (conv weights) and D are fully overwritten by their
    kernels.  Every all-reduce region of every schedule below is a contiguous slice of this order.
    Every slot starts on a 32-byte boundary (full-sector vector stores in the contraction epilogues).
    Data parallel (dp = drn_b200.parallel.GradReducer): the prop_fc weight gradient -- 44 % of the gradient bytes and, in the
    single-GPU order, the last-but-one thing produced -- moves to the END and is cut into DRN_DP_CHUNKS (4) row chunks, so that
    every all-reduce but the last runs under a contraction that leaves SMs free:
        head, FPN, backbone (graph 1)       -> B+C complete
        tail: gates, query encoder (graph 2)-> A complete.  Nothing is reduced beside the tail: its LSTM kernels are
                                               cooperative launches that need their whole grid resident, so a collective
                                               occupying SMs would serialise with them instead of overlapping
        all-reduce A+B+C (86 MB, one call)     runs WHILE chunks 0.. of the prop_fc weight gradient are computed: a chunk is 64
        prop_fc wgrad chunk i (graphs 3..)     tiles = one wave on 64 of the 74 SM pairs, 20 SMs stay free for NCCL's CTAs
        all-reduce chunk i (17 MB)             queued behind it on NCCL's stream, under chunk i+1; only the last one is exposed.
    DRN_DP_ORDER=r01 keeps the round-1 order (first part -> all-reduce B+C || whole prop_fc wgrad -> all-reduce D || tail ->
    all-reduce A) for A/B runs."""
    order = os.environ.get("DRN_DP_ORDER", "tail_first") if dp is not None else "single"
    nchunk = max(1, int(os.environ.get("DRN_DP_CHUNKS", "4"))) if order in ("tail_first", "overlap", "split") else 1
    key = ("bwd", _sig(p), tuple(names), order, nchunk)
    ent = path.graphs.get(key)
    if ent is None:
        stored, tailn = path.stored_grad_names(names), path.part2_grad_names(names)
        last = {"prop_fc.weight"} & set(names)
        early = path.early_grad_name
        # [C1 stored, head+FPN][B1 accumulated, head+FPN][B2 accumulated, backbone][A accumulated, tail][C2 stored, backbone][D]
        groups = [[n for n in names if n in stored and n not in tailn and n not in last and early(n)],
                  [n for n in names if n not in stored and n not in tailn and early(n)],
                  [n for n in names if n not in stored and n not in tailn and not early(n)],
                  [n for n in names if n in tailn],
                  [n for n in names if n in stored and n not in tailn and n not in last and not early(n)],
                  [n for n in names if n in last]]
        pad = lambda k: (k + 7) // 8 * 8  # noqa: E731
        bounds = [0]
        for g_ in groups:
            bounds.append(bounds[-1] + sum(pad(p[n].numel()) for n in g_))
        fkey = (key[1], key[2], tuple(bounds))
        flat = flat_cache.get(fkey) if flat_cache is not None else None
        if flat is None:
            flat = torch.zeros(bounds[-1], device=upstream.device, dtype=torch.float32)
            if flat_cache is not None:
                flat_cache[fkey] = flat
            if dp is not None and hasattr(dp, "register"):
                dp.register(flat)  # collective: peers map this buffer for the peer-memory all-reduce (drn_b200/parallel.py)
        if not hasattr(path, "flat_storages"):
            path.flat_storages = set()
        path.flat_storages.add(flat.untyped_storage().data_ptr())
        grads, o = {}, 0
        for g_ in groups:
            for n in g_:
                grads[n] = flat[o:o + p[n].numel()].view_as(p[n])
                o += pad(p[n].numel())
        b = bounds
        regions = {"zero": flat[b[1]:b[4]], "early": flat[:b[2]], "late": flat[b[2]:b[5]], "tail": flat[b[3]:b[4]],
                   "first": [flat[:b[3]], flat[b[4]:b[5]]], "propfc": flat[b[5]:], "all_but_propfc": flat[:b[5]]}
        if last and nchunk > 1 and p["prop_fc.weight"].shape[0] % (8 * nchunk) == 0:
            rows = p["prop_fc.weight"].shape[0] // nchunk
            per = rows * p["prop_fc.weight"].shape[1]
            regions["propfc_chunks"] = [flat[b[5] + i * per:b[5] + (i + 1) * per] for i in range(nchunk)]
        else:
            nchunk = 1
            regions["propfc_chunks"] = [regions["propfc"]]
        pc = int(os.environ.get("DRN_DP_PAIR_CLUSTERS", "0"))
        path.upstream.copy_(upstream)
        path.backward(p, grads, path.upstream)
        cap = lambda fn: _capture(fn) if use_graphs else None  # noqa: E731
        split = dp is not None

        def first_part():
            regions["zero"].zero_()
            if order == "split":
                path.backward(p, grads, path.upstream, part="early")
            else:
                path.backward(p, grads, path.upstream, tail=not split, propfc=not split)

        def late_part():
            path.backward(p, grads, path.upstream, tail=False, propfc=False, part="late")
        g1 = cap(first_part)
        g_late = cap(late_part) if order == "split" else None
        g_tail = cap(lambda: path.backward_tail(p, grads)) if split else None
        g_chunks = []
        if split:
            for i in range(nchunk):
                ch = (i, nchunk) if nchunk > 1 and last else None
                g_chunks.append((cap(lambda ch=ch: path.backward_propfc(grads, pair_clusters=pc, chunk=ch)), ch))
        if dp is not None:  # the eager run above produced a complete local gradient: reduce it in one go
            dp.reduce_regions([flat])
        ent = (g1, g_tail, g_chunks, flat, grads, regions, pc, first_part, g_late, late_part)
        path.graphs[key] = ent
        return flat, grads
    g1, g_tail, g_chunks, flat, grads, regions, pc, first_part, g_late, late_part = ent
    path.upstream.copy_(upstream)
    _mark("backward start")
    if g1 is not None:
        g1.replay()
    else:
        first_part()
    _mark("head+FPN done" if order == "split" else "head+FPN+backbone done")
    if dp is None:
        return flat, grads

    def run_tail():
        if g_tail is not None:
            g_tail.replay()
        else:
            path.backward_tail(p, grads)

    def run_chunk(i):
        g, ch = g_chunks[i]
        if g is not None:
            g.replay()
        else:
            path.backward_propfc(grads, pair_clusters=pc, chunk=ch)
    if order == "r01":
        work = dp.reduce_regions(regions["first"], wait=False)  # overlaps the prop_fc weight gradient below
        run_chunk(0)
        _mark("prop_fc wgrad done")
        work += dp.reduce_regions([regions["propfc"]], wait=False)  # overlaps the tail below
        run_tail()
        _mark("tail done")
        work += dp.reduce_regions([regions["tail"]], wait=False)
        for i, w in enumerate(work):
            dp.wait([w])
            _mark("wait: all-reduce %d done" % i)
        return flat, grads
    if order == "split":
        # head + FPN gradients (C1+B1, ~20 MB) are reduced WHILE the backbone is differentiated; everything else as tail_first
        work = dp.reduce_regions([regions["early"]], wait=False)
        if g_late is not None:
            g_late.replay()
        else:
            late_part()
        _mark("backbone done")
        run_tail()
        _mark("tail done")
        work += dp.reduce_regions([regions["late"]], wait=False)
    elif order == "overlap":
        # B+C is reduced WHILE the tail runs: its cooperative LSTM kernels (128 CTAs) stay co-resident with the collective only
        # if NCCL keeps to <= 20 SMs (NCCL_MAX_CTAS=16), which halves its bandwidth: measured slower (profiles/r02_dp_timeline8_11)
        work = dp.reduce_regions(regions["first"], wait=False)
        run_tail()
        _mark("tail done")
        work += dp.reduce_regions([regions["tail"]], wait=False)
    else:
        run_tail()
        _mark("tail done")
        work = dp.reduce_regions([regions["all_but_propfc"]], wait=False)
    for i in range(len(g_chunks)):
        run_chunk(i)
        _mark("prop_fc wgrad chunk %d done" % i)
        work += dp.reduce_regions([regions["propfc_chunks"][i]], wait=False)
    for i, w in enumerate(work):
        dp.wait([w])
        _mark("wait: all-reduce %d done" % i)
    return flat, grads


def _capture(fn):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, capture_error_mode="thread_local"):
        fn()
    return g


class _DenseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, path, training, tokens, lengths, feats, pse, gt, *params):
        p = model._tensor_dict()
        _run_forward(path, p, training, model.use_graphs, tokens, lengths, feats, pse, gt)
        ctx.model, ctx.path = model, path
        ctx.nextra = len(params)  # the trainable parameters, or the single gradient proxy (mainModel.forward)
        ctx.names = tuple(model._trainable_names)
        # the backward reads the activations from the path's static buffers: stamp which forward filled them
        path.generation = getattr(path, "generation", 0) + 1
        ctx.generation = path.generation
        return path.losses[:3] + path.poison  # poison = NaN when drn_qe_stage rejected a token id / length of this batch

    @staticmethod
    def backward(ctx, g):
        model, path = ctx.model, ctx.path
        names = list(ctx.names)
        if path.generation != ctx.generation:
            raise RuntimeError("mainModel: another forward of the same (batch, T, query-length) shape ran between this forward and "
                               "its backward; the activations live in static per-shape buffers (CUDA-graph replay), so the "
                               "backward would differentiate the wrong batch.  Call backward() before the next forward of that "
                               "shape (the reference driver does: main.py:217-236)")
        p = model._tensor_dict()
        # Gradients are produced in ONE flat static buffer (graph-replayable, and the unit of the data-parallel all-reduce).
        # They are handed to the parameters as VIEWS of that buffer (param.grad = view) instead of being returned to autograd,
        # whose AccumulateGrad would clone every one of them: ~60 copy kernels, 153 MB read + written, 0.29 ms of a 3.67 ms step
        # (in-graph timestamps, profiles/r01_insitu_v17.json).  The views stay valid until the next backward of this path; a
        # gradient that is still attached to its parameter then (no zero_grad in between: gradient accumulation, which the
        # reference driver never does, main.py:236-244) is detached from the buffer first and accumulated into afterwards.
        # DRN_GRAD_VIEWS=0: return the gradients to autograd (copies).  With views (the default) autograd sees no gradient for
        # the parameters, so per-parameter hooks (register_hook / register_post_accumulate_grad_hook: DDP, gradient scalers,
        # loggers) do NOT fire; use DRN_GRAD_VIEWS=0 with such tools, and drn_b200.parallel for data parallelism.
        # data parallel: model._dp (drn_b200/parallel.py) all-reduces the flat gradient buffer, overlapped with the tail
        views = os.environ.get("DRN_GRAD_VIEWS", "1") == "1"
        if views:
            owned = set(getattr(path, "flat_storages", ()))
            owned.update(f.untyped_storage().data_ptr() for f in model.__dict__.get("_flat_cache", {}).values())
            for n in names:
                old = p[n].grad
                if old is not None and old.untyped_storage().data_ptr() in owned:
                    p[n].grad = old.clone()
        with _nvtx("drn.backward"):
            flat, grads = _run_backward(path, p, names, g.contiguous().float(), model.use_graphs, dp=model._dp,
                                        flat_cache=model.__dict__.setdefault("_flat_cache", {}))
        if not views:
            return (None,) * 8 + tuple(grads[n] for n in names)
        for n in names:
            prm = p[n]
            if prm.grad is None:
                prm.grad = grads[n]
            else:
                prm.grad.add_(grads[n])
        return (None,) * (8 + ctx.nextra)


class mainModel(nn.Module):
    def __init__(self, vocab_size, dataset_configs, hidden_dim=512, embed_dim=300, bidirection=True,
                 graph_node_features=1024):
        super().__init__()
        cfg = dict(vars(dataset_configs))
        self.cfg = cfg
        self.first_output_dim = cfg["first_output_dim"]
        self.fpn_feature_dim = cfg["fpn_feature_dim"]
        self.feature_dim = cfg[cfg["feature_type"]]["feature_dim"]
        c1 = self.first_output_dim
        if (c1, self.fpn_feature_dim) != (256, 512):
            raise NotImplementedError("channel plan 256/512/1024 -> 512 of the reference config is assumed by the FPN "
                                      "(model/main_model.py:29 hard-codes [256, 512, 1024])")
        self.query_encoder = QueryEncoder(vocab_size, hidden_dim, embed_dim, cfg["lstm_layers"], bidirection)
        channels_list = [(self.feature_dim + 256, c1, 3, 1), (c1, c1 * 2, 3, 2), (c1 * 2, c1 * 4, 3, 2)]
        self.backbone_net = Backbone(channels_list)
        self.fpn = FPN([256, 512, 1024], 512)
        self.fcos = FCOSModule(cfg, self.fpn_feature_dim)
        self.prop_fc = nn.Linear(self.feature_dim, self.feature_dim)
        self.position_transform = nn.Linear(3, 256)
        for t in range(len(channels_list)):
            setattr(self, "qInput%d" % t, nn.Linear(1024, channels_list[t - 1][1] if t > 0 else self.feature_dim))
        self._paths = {}
        self._trainable_names = []
        # gradient reducer (drn_b200.parallel.GradReducer, a plain object): set by DataParallelDRN, or -- the reference's main.py
        # has no process-group code (main.py:52-53,99) -- by the first CUDA forward when WORLD_SIZE > 1 (`torchrun main.py`)
        self._dp = None
        self._dp_checked = False
        # CUDA graphs over the dense path (static shapes, library-owned buffers); DRN_NO_GRAPHS=1 launches kernel by kernel
        self.use_graphs = os.environ.get("DRN_NO_GRAPHS", "0") != "1"

    # ---- parameter plumbing ---------------------------------------------------------------------------------------
    # Walking named_parameters() / named_buffers() of the module tree costs ~0.3 ms of host time per call and a step needs the
    # name -> tensor map three times; the map is cached (the Parameter / buffer OBJECTS are stable: load_state_dict, optimizers
    # and .data assignments all work in place) and dropped whenever nn.Module._apply (.to / .cuda / .float) may have replaced them.
    def _apply(self, fn, *a, **k):
        self.__dict__["_pcache"] = None
        return super()._apply(fn, *a, **k)

    def _tensor_dict(self):
        d = self.__dict__.get("_pcache")
        if d is None:
            d = {k: v for k, v in self.named_parameters()}
            self.__dict__["_pnames"] = list(d)
            d.update({k: v for k, v in self.named_buffers()})
            self.__dict__["_pcache"] = d
        return d

    def _dense_trainable(self):
        d = self._tensor_dict()
        names, tensors = [], []
        for k in self.__dict__["_pnames"]:
            if k.startswith("query_encoder.textualAttention") or k.startswith("fcos.head.centerness"):
                continue  # built by the reference, never called: no gradient (language_module.py:17, fcos.py:53-56,97)
            v = d[k]
            if v.requires_grad:
                names.append(k)
                tensors.append(v)
        return names, tensors

    def _path(self, B, T, L, device):
        """Buffers + CUDA graphs of one (B, T, L) shape: ~2.3 GB at B=32, T=256.  Kept in a small LRU (DRN_MAX_PATHS, default 6:
        a training run sees the full batch, the last ragged batch and a few query-length buckets, in train and eval mode)."""
        key = (B, T, L, device.index, bool(self.cfg["is_first_stage"]))
        path = self._paths.pop(key, None)
        if path is None:
            cap = max(1, int(os.environ.get("DRN_MAX_PATHS", "6")))
            while len(self._paths) >= cap:
                old = self._paths.pop(next(iter(self._paths)))
                old.graphs.clear()
                del old
            qe = self.query_encoder
            path = DensePath(self.cfg, B, T, device, L=L, qe_hidden=qe.hidden_dim, qe_embed=qe.embed_dim)
        self._paths[key] = path  # most recently used last
        return path

    def _small_to_device(self, t, dtype, dev):
        """Host -> device copy of a SMALL tensor (tokens, lengths, ground truth, proposal boundaries: <= 1 MB) that never stalls the host: a `.to(device,
        non_blocking=True)` from PAGEABLE memory is staged synchronously by the runtime, which synchronises the stream and costs
        the whole launch-ahead of the step (~0.4 ms measured in bench.py's device-resident leg, where the query lengths used to
        be a pageable host tensor).  Pageable sources go through a ring of pinned staging buffers (an event per slot guards
        its reuse); pinned and device sources are passed straight on."""
        if t.device.type != "cpu" or t.is_pinned() or t.numel() * t.element_size() > (1 << 20):
            return t.to(dev, dtype=dtype, non_blocking=True).contiguous()
        ring = self.__dict__.setdefault("_pin_ring", {"slots": [None] * 16, "events": [None] * 16, "next": 0})
        i = ring["next"]
        ring["next"] = (i + 1) % 16
        if ring["events"][i] is not None:
            ring["events"][i].synchronize()
        nbytes = max(t.numel(), 1) * torch.empty((), dtype=dtype).element_size()
        buf = ring["slots"][i]
        if buf is None or buf.numel() < nbytes:
            buf = ring["slots"][i] = torch.empty(max(nbytes, 4096), dtype=torch.uint8, pin_memory=True)
        host = buf[:nbytes].view(dtype)[:t.numel()].view(t.shape)
        host.copy_(t)
        out = host.to(dev, non_blocking=True)
        ev = ring["events"][i] = ring["events"][i] or torch.cuda.Event()
        ev.record()
        return out

    def input_error(self):
        """Synchronising check of the sticky input-validation flags of every path (drn_qe_stage): None, or a message naming
        what was rejected since the last call.  A rejected batch has NaN losses (no synchronisation needed to notice)."""
        bits = 0
        for path in self._paths.values():
            bits |= int(path.input_err.item())
            path.input_err.zero_()
        if not bits:
            return None
        what = [w for b, w in ((1, "token id outside the vocabulary (staged as padding)"), (2, "query length outside [1, L] (clamped)")) if bits & b]
        return "mainModel input validation: " + "; ".join(what)

    # ---- forward -------------------------------------------------------------------------------------------------
    def forward(self, query_tokens, query_length, props_features, props_start_end, gt_start_end, props_num, num_frames):
        dev = self.prop_fc.weight.device
        if getattr(self, "_is_replica", False):
            raise RuntimeError("mainModel was replicated by nn.DataParallel over several devices: replicas have no parameters of "
                               "their own, so no gradient would reach the model.  Run one process per GPU instead (torchrun "
                               "main.py ... --gpu $LOCAL_RANK): mainModel then all-reduces its gradients over NCCL itself "
                               "(INTEGRATION.md section 3)")
        if dev.type != "cuda":
            raise RuntimeError("mainModel runs on a B200 through libdrn_sm100.so only (no CPU / torch fallback): call .cuda()")
        if self._dp is None and not self._dp_checked:
            self._dp_checked = True
            from drn_b200.parallel import auto_reducer
            self._dp = auto_reducer(self, dev)
        # non_blocking: a synchronous host->device copy would drain the stream every step (the host could then never run ahead
        # of the GPU); pageable sources are staged by the runtime, pinned ones are the caller's to keep unchanged until used
        query_tokens, query_length = torch.as_tensor(query_tokens), torch.as_tensor(query_length)
        if query_tokens.device.type == "cpu" and query_length.device.type == "cpu" and query_tokens.numel():
            # host tensors: validate for free, with the reference's exceptions (nn.Embedding / pack_padded_sequence,
            # language_module.py:41-42).  Device tensors are validated by drn_qe_stage (NaN losses + input_error()).
            vocab = self.query_encoder.embedding.weight.shape[0]
            if int(query_tokens.min()) < 0 or int(query_tokens.max()) >= vocab:
                raise IndexError("index out of range in self (token id outside [0, %d))" % vocab)
            if int(query_length.min()) < 1 or int(query_length.max()) > query_tokens.shape[1]:
                raise RuntimeError("query lengths must lie in [1, %d] (got %d .. %d)"
                                   % (query_tokens.shape[1], int(query_length.min()), int(query_length.max())))
        tokens = self._small_to_device(query_tokens, torch.int64, dev)
        lengths = self._small_to_device(query_length, torch.int64, dev)
        feats = props_features.to(dev, dtype=torch.float32, non_blocking=True).contiguous()
        pse = self._small_to_device(props_start_end, torch.float64, dev)
        gt = self._small_to_device(gt_start_end, gt_start_end.dtype, dev).float().contiguous()  # `.float()` as at main_model.py:74
        B, T = feats.shape[0], feats.shape[1]
        # the reference collate pads queries to the longest of the batch (dataset.py:186,198), so the token width varies from
        # batch to batch; every (B, T, L) owns ~2 GB of buffers and two CUDA graphs, so L is bucketed (zero = padding tokens,
        # masked by the lengths: results are unchanged)
        Lq = tokens.shape[1]
        Lb = max(10, (Lq + 3) // 4 * 4) if Lq > 10 else 10  # the staging kernel pads the token columns Lq .. Lb with zeros
        path = self._path(B, T, Lb, dev)
        names, tensors = self._dense_trainable()
        self._trainable_names = names
        training = self.training
        if torch.is_grad_enabled() and tensors:
            if os.environ.get("DRN_GRAD_VIEWS", "1") == "1":
                # the gradients are attached to the parameters as views by _DenseFn.backward itself, so autograd only needs ONE
                # differentiable input to call it: a per-model proxy scalar instead of 79 parameters (~0.3 ms of host time a step)
                proxy = self.__dict__.get("_grad_proxy")
                if proxy is None or proxy.device != dev:
                    proxy = self.__dict__["_grad_proxy"] = torch.zeros((), device=dev, requires_grad=True)
                losses = _DenseFn.apply(self, path, training, tokens, lengths, feats, pse, gt, proxy)
            else:
                losses = _DenseFn.apply(self, path, training, tokens, lengths, feats, pse, gt, *tensors)
        else:
            with torch.no_grad():
                p = self._tensor_dict()
                _run_forward(path, p, training, self.use_graphs, tokens, lengths, feats, pse, gt)
                path.generation = getattr(path, "generation", 0) + 1  # a pending backward of this shape is now stale
                losses = path.losses[:3] + path.poison
        loss_dict = {"loss_cls": losses[0], "loss_reg": losses[1]}
        if self.cfg["is_first_stage"]:
            loss_dict["loss_iou"] = torch.zeros(1, device=dev)  # torch.FloatTensor([0]).cuda(), loss.py:239
        elif int(path.losses[4].item()) == 0:
            loss_dict["loss_iou"] = torch.tensor([0], device=dev)  # integer constant, loss.py:194-195
        else:
            loss_dict["loss_iou"] = losses[2]
        if training:
            return None, loss_dict
        return assemble(*path.postprocess()), loss_dict
