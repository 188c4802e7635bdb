"""Host-side schedule of the DRN dense-regression path on libdrn_sm100 kernels.

`DensePath` owns the device buffers for one (B, T) shape and enqueues, on the current CUDA stream, the forward
(model/main_model.py:48-74 after the query encoder: gates, position feature, prop_fc, backbone, FPN, FCOS head, losses)
and the hand-derived backward of the same graph.  Python here only marshals pointers: every FLOP and every byte is
moved by a kernel of the C-ABI library (include/drn_b200.h).  Layouts are channels-last; tensor-core operands are
split-BF16 planes (drn_b200/planes.py).
"""
import ctypes as C
import contextlib
import os

import torch

from . import lib as L
from . import ops
from .planes import Planes

K1 = ((0, 0, 0),)
K3 = ((-1, 0, 0), (0, 0, 1), (1, 0, 2))            # (time shift, parity, weight tap) of a k3/s1/p1 conv
K3S2 = ((-1, 1, 0), (0, 0, 1), (0, 1, 2))           # k3/s2/p1 on the [T/2][2] parity view: input row 2t+r-1
K3_DGRAD = ((1, 0, 0), (0, 0, 1), (-1, 0, 2))       # dX[u] = sum_r dY[u+1-r] W_r
S2_DGRAD = (((0, 0, 1),), ((1, 0, 0), (0, 0, 2)))   # per output parity u = 2j+p
BN_MOMENTUM, BN_EPS = 0.1, 1e-5


def _lib():
    return L.load()


def _st():
    return L.stream_ptr()


def _vp(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class ConvBN:
    """Buffers of one conv -> BatchNorm(train) -> ReLU application (one level of a shared head block = one ConvBN)."""

    def __init__(self, prefix, cin, cout, k, stride, B, t_in, dev, bn_parts=None):
        self.prefix, self.cin, self.cout, self.k, self.stride = prefix, cin, cout, k, stride
        self.t_in, self.t_out = t_in, t_in // stride
        self.rows = B * self.t_out
        self.y = torch.empty(B, self.t_out, cout, device=dev)
        self.y2 = None  # second K-split slice of the forward contraction (conv0 only)
        self.stats = None  # [stats_rows][2][cout] BatchNorm partial sums written by the contraction epilogue (fused_stats)
        self.stats_rows = 0
        self.coef = torch.empty(5, cout, device=dev)
        self.sums = torch.zeros(2, cout, dtype=torch.float64, device=dev)   # kept zero between uses by the kernels
        self.counter = torch.zeros(1, dtype=torch.int32, device=dev)
        self.bcoef = torch.empty(2, cout, device=dev)
        self.dy = Planes.empty(B, self.t_out, cout, dev)
        # bn_parts: [(param prefix, channel offset, channels)] -- the fused cls|bbox tower has two BatchNorm modules
        self.bn_parts = bn_parts or [(prefix + ".1", 0, cout)]


class DensePath:
    def __init__(self, cfg, B, T, device, L=10, qe_hidden=512, qe_embed=300):
        assert T % 4 == 0, "T must be a multiple of 4 (two stride-2 levels)"
        self.cfg, self.B, self.T, self.dev = cfg, B, T, device
        self.L, self.qe_H, self.qe_E = L, qe_hidden, qe_embed
        D = cfg[cfg["feature_type"]]["feature_dim"]
        c1, F = cfg["first_output_dim"], cfg["fpn_feature_dim"]
        self.D, self.C0, self.c = D, D + 256, (c1, 2 * c1, 4 * c1)
        self.F = F
        self.Tl = (T, T // 2, T // 4)
        self.P = sum(self.Tl)
        self.strides = [float(s) for s in cfg["fpn_stride"]]
        dev = device
        e = lambda *s: torch.empty(*s, device=dev)  # noqa: E731
        z = lambda *s: torch.zeros(*s, device=dev)  # noqa: E731
        # inputs / gates
        self.f_pl = Planes.empty(B, T, D, dev)
        # query encoder (drn_qe_forward / drn_qe_backward): static token / length buffers, commands, workspace
        self.tokens = torch.zeros(B, L, dtype=torch.int64, device=dev)
        self.lengths = torch.ones(B, dtype=torch.int64, device=dev)
        # drn_qe_stage: sticky error bits (1 = token id outside the vocabulary, 2 = length outside [1, L]) and the per-batch
        # poison (0 or NaN) that the model adds to the losses
        self.input_err = torch.zeros(1, dtype=torch.int32, device=dev)
        self.poison = torch.zeros(1, device=dev)
        self.cmd_dim = 2 * qe_hidden
        self.cmd = [e(B, self.cmd_dim) for _ in range(3)]
        self.qe_ws = torch.zeros(int(_lib().drn_qe_workspace_bytes(B, L, qe_hidden, qe_embed)), dtype=torch.uint8, device=dev)
        self.qdim = (D, c1, 2 * c1)
        self.q = [e(B, n) for n in self.qdim]
        # accumulators the backward re-zeroes with ONE fill: gate gradients dq_i [B, n_i] and the 8 scalar gradients of the loss
        self.one = torch.ones(1, device=dev)
        self.bwd_zero = z(B * sum(self.qdim) + 8 + 3 * B * self.cmd_dim)
        self.dq, o = [], 0
        for n in self.qdim:
            self.dq.append(self.bwd_zero[o:o + B * n].view(B, n))
            o += B * n
        self.pgrad = self.bwd_zero[o:o + 8]
        o += 8
        self.dcmd = [self.bwd_zero[o + i * B * self.cmd_dim:o + (i + 1) * B * self.cmd_dim].view(B, self.cmd_dim) for i in range(3)]
        self.pos_in = e(B * T, 3)
        self.Pre = e(B, T, D)                       # prop_fc output before gating
        self.X0 = Planes.empty(B, T, self.C0, dev)  # [q0 * prop_fc(f) | position feature]
        self.dX0 = e(B, T, self.C0)
        self.dP_pl = Planes.empty(B, T, D, dev)
        # backbone
        self.conv = [ConvBN("backbone_net.forward_conv0", self.C0, c1, 3, 1, B, T, dev),
                     ConvBN("backbone_net.forward_conv1", c1, 2 * c1, 3, 2, B, T, dev),
                     ConvBN("backbone_net.forward_conv2", 2 * c1, 4 * c1, 3, 2, B, T // 2, dev)]
        if B * T >= 4096:  # enough rows for the K-split of conv0 to pay (see forward_main)
            ys = torch.empty(2, B, T, c1, device=dev)
            self.conv[0].y, self.conv[0].y2 = ys[0], ys[1]
        self.Cact = [Planes.empty(B, self.Tl[i], self.c[i], dev) for i in range(3)]
        self.QC = [Planes.empty(B, self.Tl[i], self.c[i], dev) for i in range(2)]
        self.dC = [e(B, self.Tl[i], self.c[i]) for i in range(3)]
        self.dQC = [e(B, self.Tl[i], self.c[i]) for i in range(2)]
        # FPN
        self.inner = [ConvBN("fpn.fpn_inner%d" % (i + 1), self.c[i], F, 1, 1, B, self.Tl[i], dev) for i in range(3)]
        self.layer = [ConvBN("fpn.fpn_layer%d" % (i + 1), F, F, 3, 1, B, self.Tl[i], dev) for i in range(3)]
        self.I = [Planes.empty(B, self.Tl[i], F, dev) for i in range(3)]
        self.Pf = [Planes.empty(B, self.Tl[i], F, dev) for i in range(3)]
        self.dI = [e(B, self.Tl[i], F) for i in range(3)]
        self.dPf = [e(B, self.Tl[i], F) for i in range(3)]
        # head (weights shared across levels, statistics per level)
        h = "fcos.head."
        self.tower = [ConvBN(h + "towers", F, 2 * F, 3, 1, B, self.Tl[i], dev,
                             bn_parts=[(h + "cls_tower.1", 0, F), (h + "bbox_tower.1", F, F)]) for i in range(3)]
        self.mix = [ConvBN(h + "mix_fc", 2 * F, F, 1, 1, B, self.Tl[i], dev) for i in range(3)]
        self.iouc = [ConvBN(h + "iou_scores", F, F // 2, 3, 1, B, self.Tl[i], dev) for i in range(3)]
        self.TW = [Planes.empty(B, self.Tl[i], 2 * F, dev) for i in range(3)]
        self.MX = [Planes.empty(B, self.Tl[i], F, dev) for i in range(3)]
        self.HI = [Planes.empty(B, self.Tl[i], F // 2, dev) for i in range(3)]
        self.dTW = [e(B, self.Tl[i], 2 * F) for i in range(3)]
        self.dMX = [e(B, self.Tl[i], F) for i in range(3)]
        self.dHI = [e(B, self.Tl[i], F // 2) for i in range(3)]
        n = B * self.P
        self.cls_raw, self.box_raw, self.iou_raw = e(n), e(n, 2), e(n)
        self.bbox = e(n, 2)
        self.dcls, self.dbox, self.diou = e(n), e(n, 2), e(n)
        self.acc = torch.zeros(8, dtype=torch.float64, device=dev)
        self.losses = z(8)
        self.upstream = z(3)
        self.gt = e(B, 2)
        self.scales = e(3)
        # eval-mode candidates (drn_postprocess): [B][3][top_n] detections / scores / locations + counts
        self.top_n = int(cfg["fcos_pre_nms_top_n"])
        self.post_det, self.post_score = z(B, 3, self.top_n, 2), z(B, 3, self.top_n)
        self.post_loc = z(B, 3, self.top_n)
        self.post_count = torch.zeros(B, 3, dtype=torch.int32, device=dev)
        self.graphs = {}
        if ops.SCHEDULE != "static":
            ops.workspace(dev)  # workspace of the contraction kernel (hybrid / stream-K folds): allocated here, not inside a graph capture
        self.launches_stage = 0
        self.Tl_c = (C.c_int * 3)(*self.Tl)
        self.strides_c = (C.c_float * 3)(*self.strides)
        self.lvl_off = [B * sum(self.Tl[:i]) for i in range(3)]
        # packed weights (planes) and weight-gradient workspaces
        self.wp = {
            "prop_fc": Planes.empty(1, D, D, dev),
            "conv0": Planes.empty(3, c1, self.C0, dev), "conv1": Planes.empty(3, 2 * c1, c1, dev),
            "conv2": Planes.empty(3, 4 * c1, 2 * c1, dev),
            "towers": Planes.empty(3, 2 * F, F, dev), "mix": Planes.empty(1, F, 2 * F, dev),
            "iouc": Planes.empty(3, F // 2, F, dev),
        }
        for i in range(3):
            self.wp["inner%d" % i] = Planes.empty(1, F, self.c[i], dev)
            self.wp["layer%d" % i] = Planes.empty(3, F, F, dev)
        self.tower_bias = e(2 * F)
        # weight-gradient workspaces: [slices][k][cout][cin] fp32 partial sums (K-splits x pyramid levels), summed and
        # re-laid out to the parameter layout [cout][cin][k] by ONE drn_unpack_conv_wgrads launch at the end of the backward
        self.ws, self.ws_split = {}, {}

        def ws_alloc(name, blks):
            blk0 = blks[0]
            splits = [self._wgrad_split(blk) for blk in blks]
            self.ws_split[name] = splits
            self.ws[name] = e(sum(splits), blk0.k, blk0.cout, blk0.cin)

        for i in range(3):
            ws_alloc("conv%d" % i, [self.conv[i]])
            ws_alloc("inner%d" % i, [self.inner[i]])
            ws_alloc("layer%d" % i, [self.layer[i]])
        ws_alloc("towers", self.tower)
        ws_alloc("mix", self.mix)
        ws_alloc("iouc", self.iouc)
        # BatchNorm statistics fused into the contraction epilogue (drn_gemm_t.stats) for every conv that runs through the grouped
        # CTA-pair launches without a K-split: the statistics pass then reduces [rows/32][2][C] partial sums instead of re-reading
        # the conv output.  DRN_FUSED_STATS=0: statistics pass over y (A/B).
        self.fused_stats = os.environ.get("DRN_FUSED_STATS", "1") == "1"
        if self.fused_stats:
            for i in range(3):
                for blk, a_pl, w in ((self.inner[i], self.Cact[i], self.wp["inner%d" % i]), (self.layer[i], self.I[i], self.wp["layer%d" % i]),
                                     (self.tower[i], self.Pf[i], self.wp["towers"]), (self.mix[i], self.TW[i], self.wp["mix"]),
                                     (self.iouc[i], self.MX[i], self.wp["iouc"])):
                    rows = int(_lib().drn_gemm_stats_rows(C.byref(self._conv_desc(blk, a_pl, w, engine=2))))
                    if rows < 1:
                        raise RuntimeError("drn_gemm_stats_rows: " + _lib().drn_last_error().decode())
                    blk.stats, blk.stats_rows = torch.zeros(rows, 2, blk.cout, device=dev), rows
        self.iou_branch_on = not cfg["is_first_stage"]
        self.gamma = float(cfg["fcos_loss_gamma"][0] if isinstance(cfg["fcos_loss_gamma"], (list, tuple)) else cfg["fcos_loss_gamma"])
        self.alpha = float(cfg["fcos_loss_alpha"][0] if isinstance(cfg["fcos_loss_alpha"], (list, tuple)) else cfg["fcos_loss_alpha"])
        self.launches = self.launches_fwd = self.launches_bwd = self.launches_pre = 0
        # optional side branch (DRN_SIDE=1): the HBM-bound weight packing / weight-gradient unpacking are forked off where the
        # latency-bound LSTM recurrences start (their CTAs use no shared memory, so -- unlike a second contraction -- they can
        # co-reside with the LSTM).  Measured NEUTRAL on B200 (r01 v11 A/B, gpurun_out/ab_v11.log: 3.732 vs 3.737 ms per step:
        # what the packing gains the recurrence loses to the loaded memory system), so the default is ONE stream.
        self.side = torch.cuda.Stream(device=dev) if os.environ.get("DRN_SIDE", "0") == "1" else None

    # ---------------------------------------------------------------------------------------------------------------
    # small launch helpers
    # ---------------------------------------------------------------------------------------------------------------
    def _fork(self):
        """Context manager: the body is enqueued on the side stream, ordered after everything enqueued so far on the current
        stream (inside a CUDA-graph capture this becomes a parallel branch of the graph).  `_join()` orders the current stream
        after the side branch.  Without a side stream both are no-ops."""
        if self.side is None:
            return contextlib.nullcontext()
        self.side.wait_stream(torch.cuda.current_stream(self.dev))
        return torch.cuda.stream(self.side)

    def _join(self):
        if self.side is not None:
            torch.cuda.current_stream(self.dev).wait_stream(self.side)

    def _chk(self, rc, what):
        self.launches += 1
        L.check(rc, what)

    def _gemm(self, *a, **k):
        self.launches += 1
        ops.gemm(*a, **k)

    def _sgemm(self, A, sam, sak, Bm, sbk, sbn, Cm, ldc, M, N, K, bias=None, relu=0, accumulate=0):
        self._chk(_lib().drn_sgemm(_vp(A), C.c_int64(sam), C.c_int64(sak), _vp(Bm), C.c_int64(sbk), C.c_int64(sbn), _vp(Cm),
                                   C.c_int64(ldc), M, N, K, _vp(bias), relu, accumulate, _st()), "sgemm")

    def _qe_desc(self, p, grads=None):
        """drn_qe_t for this path (reference model/language_module.py:9-62 parameter names)."""
        q = L.Qe()
        q.B, q.L, q.H, q.E, q.tok_ld = self.B, self.L, self.qe_H, self.qe_E, self.L
        q.tokens, q.lengths = self.tokens.data_ptr(), self.lengths.data_ptr()
        pre = "query_encoder."
        g = (lambda n: grads[pre + n].data_ptr() if (pre + n) in grads else None) if grads is not None else (lambda n: None)
        q.emb, q.g_emb = p[pre + "embedding.weight"].data_ptr(), g("embedding.weight")
        for d, suf in enumerate(("", "_reverse")):
            for f, n in (("w_ih", "weight_ih_l0"), ("w_hh", "weight_hh_l0"), ("b_ih", "bias_ih_l0"), ("b_hh", "bias_hh_l0")):
                getattr(q, f)[d] = p[pre + "biLSTM." + n + suf].data_ptr()
                getattr(q, "g_" + f)[d] = g("biLSTM." + n + suf)
        q.w1, q.b1, q.g_w1, q.g_b1 = p[pre + "qInput.weight"].data_ptr(), p[pre + "qInput.bias"].data_ptr(), g("qInput.weight"), g("qInput.bias")
        for t in range(3):
            q.w2[t], q.b2[t] = p[pre + "qInput%d.weight" % t].data_ptr(), p[pre + "qInput%d.bias" % t].data_ptr()
            q.g_w2[t], q.g_b2[t] = g("qInput%d.weight" % t), g("qInput%d.bias" % t)
            q.cmd[t], q.dcmd[t] = self.cmd[t].data_ptr(), self.dcmd[t].data_ptr()
        q.wa, q.ba = p[pre + "cmd_inter2logits.weight"].data_ptr(), p[pre + "cmd_inter2logits.bias"].data_ptr()
        q.g_wa, q.g_ba = g("cmd_inter2logits.weight"), g("cmd_inter2logits.bias")
        q.workspace, q.workspace_bytes = self.qe_ws.data_ptr(), self.qe_ws.numel()
        return q

    def _split(self, src2d, dst, dst_col0=0):
        rows, Cn = src2d.shape
        self._chk(_lib().drn_split_planes(_vp(src2d), C.c_int64(rows), Cn, C.c_int64(src2d.stride(0)), _vp(dst.data),
                                          C.c_int64(dst.C), dst_col0, C.c_int64(dst.plane_stride), _st()), "split_planes")

    @staticmethod
    def _pack_item(w, dst, o0=0):
        """weight w [O][C][k] fp32 -> tap-major planes `dst` [k][Ototal][C] at row offset o0."""
        it = L.PackItem()
        it.src, it.planes, it.Ototal, it.plane_stride = w.data_ptr(), dst.data.data_ptr(), dst.T, dst.plane_stride
        it.O, it.C, it.k, it.o0 = w.shape[0], w.shape[1], (w.shape[2] if w.dim() == 3 else 1), o0
        return it

    def _unpack_item(self, grad, name, o0=0):
        """workspace `name` [slices][k][Ototal][C] fp32 -> gradient [O][C][k] = sum of the slices (rows o0 .. o0+O)."""
        ws = self.ws[name]
        it = L.PackItem()
        it.src, it.grad, it.Ototal = ws.data_ptr(), grad.data_ptr(), ws.shape[2]
        it.O, it.C, it.k, it.o0 = grad.shape[0], grad.shape[1], (grad.shape[2] if grad.dim() == 3 else 1), o0
        it.nslices, it.slice_stride = ws.shape[0], ws.stride(0)
        return it

    def pack_weights(self, p):
        """fp32 parameters -> tap-major split-BF16 planes (once per forward: the optimizer changes them every step)."""
        h = "fcos.head."
        items = [self._pack_item(p["prop_fc.weight"], self.wp["prop_fc"])]
        for i in range(3):
            items.append(self._pack_item(p["backbone_net.forward_conv%d.0.weight" % i], self.wp["conv%d" % i]))
            items.append(self._pack_item(p["fpn.fpn_inner%d.0.weight" % (i + 1)], self.wp["inner%d" % i]))
            items.append(self._pack_item(p["fpn.fpn_layer%d.0.weight" % (i + 1)], self.wp["layer%d" % i]))
        items.append(self._pack_item(p[h + "cls_tower.0.weight"], self.wp["towers"], 0))
        items.append(self._pack_item(p[h + "bbox_tower.0.weight"], self.wp["towers"], self.F))
        items.append(self._pack_item(p[h + "mix_fc.0.weight"], self.wp["mix"]))
        items.append(self._pack_item(p[h + "iou_scores.0.weight"], self.wp["iouc"]))
        arr = (L.PackItem * len(items))(*items)
        torch.cat([p[h + "cls_tower.0.bias"], p[h + "bbox_tower.0.bias"]], out=self.tower_bias)
        self._chk(_lib().drn_pack_conv_weights(len(items), arr, _st()), "pack_conv_weights")

    def _bn_job(self, blk, p, grads=None, da=None, out_a=None, up=None, gate=None, out_qa=None, y2=None, partials=False):
        """drn_bn_job_t of one conv block (model/basic_blocks.py:22-30): statistics / apply / backward operands.
        partials: the statistics come from the partial sums the contraction epilogue wrote into blk.stats."""
        j = L.BnJob()
        j.y, j.B, j.T, j.C = blk.y.data_ptr(), self.B, blk.t_out, blk.cout
        if y2 is not None:
            j.y2 = y2.data_ptr()
        if partials:
            j.partials, j.partial_rows = blk.stats.data_ptr(), blk.stats_rows
        j.nparts = len(blk.bn_parts)
        for i, (pre, c0, n) in enumerate(blk.bn_parts):
            a = j.parts[i]
            a.c0, a.n = c0, n
            a.gamma, a.beta = p[pre + ".weight"].data_ptr(), p[pre + ".bias"].data_ptr()
            a.running_mean, a.running_var = p[pre + ".running_mean"].data_ptr(), p[pre + ".running_var"].data_ptr()
            a.num_batches_tracked = p[pre + ".num_batches_tracked"].data_ptr()
            if grads is not None and (pre + ".weight") in grads:
                a.dgamma, a.dbeta = grads[pre + ".weight"].data_ptr(), grads[pre + ".bias"].data_ptr()
        j.coef, j.sums, j.counter, j.bcoef = blk.coef.data_ptr(), blk.sums.data_ptr(), blk.counter.data_ptr(), blk.bcoef.data_ptr()
        if up is not None:
            j.up, j.up_plane_stride = up.data.data_ptr(), up.plane_stride
        if gate is not None:
            j.gate = gate.data_ptr()
        if out_a is not None:
            j.out_a, j.a_plane_stride = out_a.data.data_ptr(), out_a.plane_stride
        if out_qa is not None:
            j.out_qa, j.qa_plane_stride = out_qa.data.data_ptr(), out_qa.plane_stride
        if da is not None:
            j.da = da.data_ptr()
        j.dy, j.dy_plane_stride = blk.dy.data.data_ptr(), blk.dy.plane_stride
        return j

    def _bn_fwd(self, jobs, training, shared=False):
        """Train-mode BatchNorm + ReLU of up to 3 independent conv blocks: ONE statistics launch + ONE apply launch.
        shared=True: the jobs are the pyramid levels of ONE head block (same BatchNorm modules); their running statistics are
        then updated by a separate ordered kernel, level after level as in the reference (fcos.py:93-102)."""
        arr = (L.BnJob * len(jobs))(*jobs)
        mode = 0 if not training else (2 if shared else 1)
        self._chk(_lib().drn_bn_stats_multi(len(jobs), arr, C.c_float(BN_MOMENTUM), C.c_float(BN_EPS), mode, _st()), "bn_stats")
        if mode == 2:
            self._chk(_lib().drn_bn_running_update(len(jobs), arr, C.c_float(BN_MOMENTUM), _st()), "bn_running_update")
        self._chk(_lib().drn_bn_relu_apply_multi(len(jobs), arr, _st()), "bn_relu_apply")

    def _bn_stats(self, jobs, training):
        arr = (L.BnJob * len(jobs))(*jobs)
        self._chk(_lib().drn_bn_stats_multi(len(jobs), arr, C.c_float(BN_MOMENTUM), C.c_float(BN_EPS), 1 if training else 0, _st()),
                  "bn_stats")

    def _bn_apply(self, jobs):
        arr = (L.BnJob * len(jobs))(*jobs)
        self._chk(_lib().drn_bn_relu_apply_multi(len(jobs), arr, _st()), "bn_relu_apply")

    def _bn_bwd(self, jobs):
        """BatchNorm + ReLU backward of up to 3 conv blocks: sums of g and g*xhat (+ dgamma / dbeta), then dy planes."""
        arr = (L.BnJob * len(jobs))(*jobs)
        self._chk(_lib().drn_bn_bwd_reduce_multi(len(jobs), arr, _st()), "bn_bwd_reduce")
        self._chk(_lib().drn_bn_bwd_apply_multi(len(jobs), arr, _st()), "bn_bwd_apply")

    def _conv_desc(self, blk, a_pl, w_pl, bias=None, engine=None, stats=False):
        par = blk.stride
        taps = K1 if blk.k == 1 else (K3 if blk.stride == 1 else K3S2)
        g = ops.desc(L.GEMM_ROWS, a_pl.desc(par), w_pl.desc(), self.B, blk.t_out, blk.cout, K=blk.cin, taps=taps,
                     out=blk.y, bias=bias, engine=engine)
        if stats:
            g.stats = blk.stats.data_ptr()
        return g

    def _conv(self, blk, a_pl, w_pl, bias=None):
        self.launches += 1
        g = self._conv_desc(blk, a_pl, w_pl, bias)
        ws, nb = ops._ws_args(None)
        L.check(_lib().drn_gemm_ws(C.byref(g), ws, nb, _st()), "drn_gemm")

    # ---------------------------------------------------------------------------------------------------------------
    # forward
    # ---------------------------------------------------------------------------------------------------------------
    def stage_query(self, tokens, lengths, gt, vocab=None):
        """Eager: the caller's query tokens [B, <= L] / lengths / ground truth -> the static buffers the replayable parts read.
        Token ids and lengths are validated on the way (drn_qe_stage): the kernels index the embedding table, its gradient and
        the LSTM output with them."""
        self.gt.copy_(gt)
        ncols = tokens.shape[1]
        if ncols > self.L or tokens.stride(1) != 1:
            raise ValueError("stage_query: tokens [B, %d] do not fit the path's query width L = %d" % (ncols, self.L))
        self._chk(_lib().drn_qe_stage(_vp(tokens), C.c_int64(tokens.stride(0)), ncols, _vp(lengths), self.B, self.L,
                                      int(vocab) if vocab is not None else (1 << 30), _vp(self.tokens), _vp(self.lengths),
                                      _vp(self.input_err), _vp(self.poison), _st()), "qe_stage")

    def stage_feats(self, p, feats, pse):
        """Eager: reads the caller's clip features and proposal boundaries and fills the static operand buffers (split planes
        of the clip features, position feature)."""
        lib, B, T = _lib(), self.B, self.T
        self.launches = 0
        # position feature -> X0[:, :, D:] (main_model.py:53-55, backbone.py:31-32)
        self._chk(lib.drn_pos_feature(_vp(pse), _vp(p["position_transform.weight"]), _vp(p["position_transform.bias"]),
                                      C.c_int64(B * T), 256, _vp(self.X0.data), C.c_int64(self.C0), self.D,
                                      C.c_int64(self.X0.plane_stride), _vp(self.pos_in), _st()), "pos_feature")
        self._split(feats.view(B * T, self.D), self.f_pl)
        self.launches_stage = self.launches

    def stage_inputs(self, p, tokens, lengths, feats, pse, gt):
        self.stage_query(tokens, lengths, gt, vocab=p["query_encoder.embedding.weight"].shape[0])
        self.stage_feats(p, feats, pse)

    def forward(self, p, tokens, lengths, feats, pse, gt, training):
        """p: name -> parameter/buffer tensor (fp32 on device); tokens [B,L] i64, lengths [B] i64 (device);
        feats [B,T,D] fp32, pse [B,T,2] f64, gt [B,2] f32.  Fills self.losses / raw head outputs.  (Single-stream order; the
        overlapped schedule is model/main_model.py:_run_forward.)"""
        self.stage_inputs(p, tokens, lengths, feats, pse, gt)
        self.forward_pre(p)
        self.forward_main(p, training)

    def forward_pre(self, p):
        """Replayable, reads parameters and the staged query only: query encoder and gates, with the weight packing (HBM-bound,
        364 MB of traffic, no dependence on the query) forked off as a side branch at the point where the latency-bound
        recurrence starts.  The branch begins with a tiny kernel (`torch.cat` of two bias vectors), so that the LSTM's CTAs are
        resident before the packing grid is dispatched: a grid launched EARLIER than the LSTM would have to be dispatched
        completely before the LSTM can start, i.e. it would run before it rather than under it."""
        lib, B = _lib(), self.B
        self.launches = 0
        qd = self._qe_desc(p)
        nq = lib.drn_qe_launch_count(B, self.L, self.qe_H, 0)  # kernels enqueued inside drn_qe_forward
        if self.side is None:
            self.pack_weights(p)
            self._chk(lib.drn_qe_forward(C.byref(qd), _st()), "qe_forward")
            self.launches += nq - 1
        else:
            # query encoder -> three command vectors (model/main_model.py:47, language_module.py:38-62)
            self._chk(lib.drn_qe_forward_part(C.byref(qd), 1, _st()), "qe_forward[1]")
            with self._fork():
                self.pack_weights(p)
            self._chk(lib.drn_qe_forward_part(C.byref(qd), 2, _st()), "qe_forward[2]")
            self.launches += nq - 2
        # gates q_i = qInput_i(cmd_i) (model/main_model.py:48-50): M = B rows only, weight-streaming (drn_linear_fwd_batch)
        K = self.cmd_dim
        jobs = (L.LinearJob * 3)()
        for i in range(3):
            j, n = jobs[i], self.qdim[i]
            j.x, j.ldx, j.W, j.ldw = self.cmd[i].data_ptr(), K, p["qInput%d.weight" % i].data_ptr(), K
            j.bias, j.out, j.ldo, j.B, j.N, j.K, j.relu = p["qInput%d.bias" % i].data_ptr(), self.q[i].data_ptr(), n, B, n, K, 0
        self._chk(lib.drn_linear_fwd_batch(3, jobs, _st()), "gates")
        self._join()
        self.launches_pre = self.launches

    def forward_main(self, p, training):
        """Replayable: prop_fc -> backbone -> FPN -> head -> losses, on the staged planes, packed weights and gates."""
        lib, B, T = _lib(), self.B, self.T
        h = "fcos.head."
        self.launches = self.launches_stage + self.launches_pre
        # prop_fc (main_model.py:59) with the level-0 gate fused in the epilogue: Pre = W f + b (kept for the gate gradient),
        # X0[:, :, :D] = planes(q0 * Pre) (backbone.py:28-30; the cat with the position channels is the layout).  Running
        # this contraction beside the query encoder was measured to buy nothing (the persistent kernel and the LSTM cannot
        # share an SM: shared memory, scripts/overlap_probe.py), while the separate gating pass cost 39 us.
        if os.environ.get("DRN_FUSE_GATE", "1") == "1":
            self._gemm(L.GEMM_ROWS, self.f_pl.desc(), self.wp["prop_fc"].desc(), B, T, self.D, K=self.D, bias=p["prop_fc.bias"],
                       out2=self.Pre, rowscale=self.q[0], outp=self.X0)
        else:  # A/B: separate gating pass
            self._gemm(L.GEMM_ROWS, self.f_pl.desc(), self.wp["prop_fc"].desc(), B, T, self.D, K=self.D, bias=p["prop_fc.bias"],
                       out=self.Pre)
            self._chk(lib.drn_gate_planes(_vp(self.Pre), _vp(self.q[0]), B, T, self.D, _vp(self.X0.data), C.c_int64(self.C0), 0,
                                          C.c_int64(self.X0.plane_stride), _st()), "gate_planes")
        # backbone (backbone.py:27-34)
        src = self.X0
        for i in range(3):
            blk = self.conv[i]
            y2 = None
            if i == 0 and training and blk.y2 is not None:
                # conv0: one 256-wide N tile and K = 3 x 4352: only 32 pair tiles -> split K in two slices (64 of the 74 SM
                # pairs busy); the BatchNorm statistics pass folds the second slice into y
                g = self._conv_desc(blk, src, self.wp["conv0"], engine=2)
                g.split_k, g.out_split_stride = 2, blk.y2.data_ptr() - blk.y.data_ptr() >> 2
                self._group([g])
                y2 = blk.y2
            else:
                self._conv(blk, src, self.wp["conv%d" % i])
            if i < 2:
                self._bn_fwd([self._bn_job(blk, p, out_a=self.Cact[i], gate=self.q[i + 1], out_qa=self.QC[i], y2=y2)], training)
                src = self.QC[i]
            else:
                self._bn_fwd([self._bn_job(blk, p, out_a=self.Cact[i])], training)
        # FPN top-down (FPN.py:54-69).  The three lateral 1x1 convs are independent -> one grouped launch; their applies run
        # top-down (the upsample-add needs the level above); then the three 3-tap output convs, again one launch.  Every FPN
        # block owns its BatchNorm module, so the order of the running-statistics updates is immaterial here.
        fs = training and self.fused_stats  # BatchNorm partial sums from the contraction epilogues (train mode only)
        self._group([self._conv_desc(self.inner[i], self.Cact[i], self.wp["inner%d" % i], engine=2, stats=fs) for i in range(3)])
        jobs = [self._bn_job(self.inner[i], p, out_a=self.I[i], up=self.I[i + 1] if i < 2 else None, partials=fs) for i in range(3)]
        self._bn_stats(jobs, training)
        for i in (2, 1, 0):  # the upsample-add reads the level above: applies run top-down
            self._bn_apply([jobs[i]])
        self._group([self._conv_desc(self.layer[i], self.I[i], self.wp["layer%d" % i], engine=2, stats=fs) for i in range(3)])
        self._bn_fwd([self._bn_job(self.layer[i], p, out_a=self.Pf[i], partials=fs) for i in range(3)], training)
        # head (fcos.py:93-102): shared weights, per-level batch statistics.  Each shared conv runs its three levels in ONE
        # grouped launch; the BatchNorm statistics kernels follow in level order, which keeps the order of the three
        # running-statistics updates of every shared module (levels ascending) exactly as in the reference.
        self._group([self._conv_desc(self.tower[l], self.Pf[l], self.wp["towers"], bias=self.tower_bias, engine=2, stats=fs)
                     for l in range(3)])
        self._bn_fwd([self._bn_job(self.tower[l], p, out_a=self.TW[l], partials=fs) for l in range(3)], training, shared=True)
        self._group([self._conv_desc(self.mix[l], self.TW[l], self.wp["mix"], bias=p[h + "mix_fc.0.bias"], engine=2, stats=fs)
                     for l in range(3)])
        self._bn_fwd([self._bn_job(self.mix[l], p, out_a=self.MX[l], partials=fs) for l in range(3)], training, shared=True)
        self._group([self._conv_desc(self.iouc[l], self.MX[l], self.wp["iouc"], bias=p[h + "iou_scores.0.bias"], engine=2, stats=fs)
                     for l in range(3)])
        self._bn_fwd([self._bn_job(self.iouc[l], p, out_a=self.HI[l], partials=fs) for l in range(3)], training, shared=True)
        # cls_logits / bbox_pred / iou_scores.3 on every level: one launch (fcos.py:95-102)
        self._chk(lib.drn_head_proj_fwd(C.byref(self._head_levels()), _vp(p[h + "cls_logits.weight"]), _vp(p[h + "cls_logits.bias"]),
                                        _vp(p[h + "bbox_pred.weight"]), _vp(p[h + "bbox_pred.bias"]),
                                        _vp(p[h + "iou_scores.3.weight"]), _vp(p[h + "iou_scores.3.bias"]), _vp(self.cls_raw),
                                        _vp(self.box_raw), _vp(self.iou_raw), _st()), "head_proj_fwd")
        torch.cat([p[h + "scales.%d.scale" % l] for l in range(3)], out=self.scales)
        self._chk(lib.drn_fcos_loss_fwd(3, B, self.Tl_c, self.strides_c, _vp(self.cls_raw), _vp(self.box_raw), _vp(self.iou_raw),
                                        _vp(self.scales), _vp(self.gt), C.c_float(self.gamma), C.c_float(self.alpha),
                                        1 if self.iou_branch_on else 0, _vp(self.bbox), _vp(self.acc), _vp(self.losses), _st()),
                  "fcos_loss_fwd")
        self.launches += 1  # finalize kernel inside drn_fcos_loss_fwd
        self.launches_fwd = self.launches

    def _head_levels(self):
        hl = L.HeadLevels()
        hl.nlevels, hl.B, hl.F = 3, self.B, self.F
        for l in range(3):
            hl.T[l] = self.Tl[l]
            hl.tower[l], hl.tower_plane_stride[l] = self.TW[l].data.data_ptr(), self.TW[l].plane_stride
            hl.iou_hidden[l], hl.iou_hidden_plane_stride[l] = self.HI[l].data.data_ptr(), self.HI[l].plane_stride
            hl.d_tower[l] = self.dTW[l].data_ptr()
        return hl

    def postprocess(self):
        """Eval only (fcos.py:172-191 -> inference.py:49-136): candidate selection for every (sample, level) in one launch.
        Returns the device tensors (det [B,3,K,2], score, loc [B,3,K], count [B,3]); model/inference.py:assemble compacts them."""
        cfg = self.cfg
        self._chk(_lib().drn_postprocess(3, self.B, self.Tl_c, self.strides_c, _vp(self.cls_raw), _vp(self.bbox), _vp(self.iou_raw),
                                         C.c_float(float(cfg["fcos_inference_thr"])), self.top_n,
                                         0 if cfg["is_first_stage"] else 1, _vp(self.post_det), _vp(self.post_score),
                                         _vp(self.post_loc), _vp(self.post_count), _st()), "postprocess")
        return self.post_det, self.post_score, self.post_loc, self.post_count

    # ---------------------------------------------------------------------------------------------------------------
    # backward
    # ---------------------------------------------------------------------------------------------------------------
    def _group(self, descs):
        """Independent contractions -> one launch of the persistent CTA-pair kernel per <= 6 problems."""
        self.launches += ops.gemm_group(descs)

    @staticmethod
    def _wgrad_split(blk):
        """K-splits of a weight gradient: ~2048 contraction rows (32 K-blocks) per 256 x 256 tile, so the tiles of the three
        pyramid levels (8192 / 4096 / 2048 rows at B=32, T=256) cost the same and fill the 74 SM pairs evenly.  The layers
        with few output tiles (conv1: 6, FPN laterals: 2-8 before the split) are cut twice as fine: their launch is then no
        longer bound by one 32-iteration tile per busy pair (makespan 32 -> 16 and 40 -> 24 k-iterations on 74 pairs)."""
        s = max(1, min(8, blk.rows // 2048))
        if blk.prefix.startswith("fpn.fpn_inner") or blk.prefix.endswith("forward_conv1"):
            # never more slices than 64-row K-blocks (drn_gemm rejects a slice that would stay unwritten)
            t, b = blk.t_out, blk.rows // blk.t_out
            if t >= 64:
                kblocks = b * ((t + 63) // 64)
            else:
                rk = 1 << max(0, (t - 1).bit_length())
                kblocks = (b + 64 // rk - 1) // (64 // rk)
            s = max(1, min(16, 2 * s, kblocks))
        return s

    def _wgrad_desc(self, blk, x_pl, name, idx=0):
        """Weight gradient of `blk` into its slices of the workspace `name` (idx = pyramid level for shared head convs)."""
        ws, splits = self.ws[name], self.ws_split[name]
        taps = K1 if blk.k == 1 else (K3 if blk.stride == 1 else K3S2)
        s0 = sum(splits[:idx])
        return ops.desc(L.GEMM_WGRAD, blk.dy.desc(), x_pl.desc(blk.stride), self.B, blk.t_out, blk.cin, M=blk.cout, taps=taps,
                        out=ws[s0], out_ld=blk.cin, out_tap_stride=blk.cout * blk.cin, out_split_stride=ws.stride(0),
                        split_k=splits[idx], engine=2)

    def _dgrad_descs(self, blk, w_pl, out, mode=L.OUT_STORE, rowscale=None, out2=None):
        if blk.stride == 1:
            taps = K1 if blk.k == 1 else K3_DGRAD
            return [ops.desc(L.GEMM_ROWS, blk.dy.desc(), w_pl.desc(), self.B, blk.t_out, blk.cin, K=blk.cout, taps=taps, b_mn=1,
                             out=out, out_mode=mode, rowscale=rowscale, out2=out2, engine=2)]
        return [ops.desc(L.GEMM_ROWS, blk.dy.desc(), w_pl.desc(), self.B, blk.t_out, blk.cin, K=blk.cout, taps=S2_DGRAD[par],
                         b_mn=1, out=out, out_mode=mode, rowscale=rowscale, out2=out2, out_T=blk.t_in, out_t_mul=2,
                         out_t_add=par, engine=2) for par in (0, 1)]

    def stored_grad_names(self, names):
        """Gradients a kernel of `backward` fully OVERWRITES (never accumulates into): prop_fc.weight (its contraction stores)
        and every conv weight that goes through the slice workspaces + drn_unpack_conv_wgrads."""
        h = "fcos.head."
        st = {"prop_fc.weight", h + "cls_tower.0.weight", h + "bbox_tower.0.weight", "qInput0.weight", "qInput1.weight", "qInput2.weight"}
        for i in range(3):
            st |= {"backbone_net.forward_conv%d.0.weight" % i, "fpn.fpn_inner%d.0.weight" % (i + 1), "fpn.fpn_layer%d.0.weight" % (i + 1)}
        if self.iou_branch_on:
            st |= {h + "mix_fc.0.weight", h + "iou_scores.0.weight"}
        return {n for n in names if n in st}

    def part2_grad_names(self, names):
        """Gradients produced by `backward_tail` (everything else is final when `backward(..., tail=False)` returns)."""
        return {n for n in names if n.startswith("qInput") or n.startswith("query_encoder.")}

    @staticmethod
    def early_grad_name(n):
        """Gradients that are FINAL once the head and the FPN have been differentiated (`backward(part="early")`): the
        data-parallel schedule starts their all-reduce while the backbone backward still runs."""
        return n.startswith("fcos.") or n.startswith("fpn.")

    def backward(self, p, grads, upstream, tail=True, propfc=True, part="all"):
        """grads: name -> zero-initialised fp32 tensor for every parameter that wants a gradient (filled in place).
        upstream: [3] fp32 device tensor = d(total)/d(loss_cls, loss_reg, loss_iou).  tail=False stops before `backward_tail`;
        propfc=False also leaves out the prop_fc weight gradient (`backward_propfc`): the data-parallel schedule runs the
        parts as separate graphs with an all-reduce started after each.  part="early": loss, head and FPN only (their weight
        gradients unpacked at once); part="late": the backbone, from the activations' gradients the early part left behind."""
        lib, B = _lib(), self.B
        h = "fcos.head."
        F = self.F
        self.launches = 0
        iou_on = self.iou_branch_on and (h + "mix_fc.0.weight") in grads
        lv = range(3)
        if part != "late":
            self._backward_head_fpn(p, grads, upstream, iou_on)
            if part == "early":
                self._unpack_wgrads(grads, iou_on, which="early")
                self.launches_bwd = self.launches
                return
        # backbone
        for i in (2, 1):
            blk = self.conv[i]
            self._bn_bwd([self._bn_job(blk, p, grads, da=self.dC[i])])
            self._group([self._wgrad_desc(blk, self.QC[i - 1], "conv%d" % i)] +
                        self._dgrad_descs(blk, self.wp["conv%d" % i], self.dC[i - 1], mode=L.OUT_ADD, rowscale=self.q[i],
                                          out2=self.dQC[i - 1]))
            a = self.Cact[i - 1]
            self._chk(lib.drn_gate_reduce(_vp(self.dQC[i - 1]), C.c_int64(a.C), _vp(a.data), C.c_int64(a.C),
                                          C.c_int64(a.plane_stride), 1, B, self.Tl[i - 1], a.C, _vp(self.dq[i]), None, None,
                                          C.c_int64(0), None, _st()), "gate_reduce")
        blk = self.conv[0]
        self._bn_bwd([self._bn_job(blk, p, grads, da=self.dC[0])])
        self._group([self._wgrad_desc(blk, self.X0, "conv0")] + self._dgrad_descs(blk, self.wp["conv0"], self.dX0))
        self._chk(lib.drn_gate_reduce(_vp(self.dX0), C.c_int64(self.C0), _vp(self.Pre), C.c_int64(self.D), C.c_int64(0), 0, B,
                                      self.T, self.D, _vp(self.dq[0]), _vp(self.q[0]), _vp(self.dP_pl.data),
                                      C.c_int64(self.dP_pl.plane_stride), _vp(grads["prop_fc.bias"]), _st()), "gate0_bwd")
        self._chk(lib.drn_pos_bwd(_vp(self.dX0), C.c_int64(self.C0), self.D, _vp(self.pos_in), C.c_int64(B * self.T), 256,
                                  _vp(grads["position_transform.weight"]), _vp(grads["position_transform.bias"]), _st()),
                  "pos_bwd")
        # partial sums (K-splits x levels) in tap-major workspaces -> parameter layout [O][C][k]: here, or (single-GPU schedule)
        # as a side branch under the BPTT of the tail.  [Forking it beside the prop_fc weight gradient below was measured
        # 50 us SLOWER than running it first: r01 v11 A/B.]
        defer = tail and self.side is not None and part == "all"
        if not defer:
            self._unpack_wgrads(grads, iou_on, which="late" if part == "late" else "all")
        if propfc:
            self.backward_propfc(grads)
        self.launches_bwd = self.launches
        if tail:
            self.backward_tail(p, grads, unpack=(iou_on,) if defer else None)

    def _backward_head_fpn(self, p, grads, upstream, iou_on):
        """Loss, head and FPN backward (everything above the backbone): leaves dC[0..2] for the backbone part."""
        lib, B = _lib(), self.B
        h = "fcos.head."
        F = self.F
        self.bwd_zero.zero_()
        self._chk(lib.drn_fcos_loss_bwd(3, B, self.Tl_c, self.strides_c, _vp(self.cls_raw), _vp(self.box_raw), _vp(self.iou_raw),
                                        _vp(self.scales), _vp(self.gt), C.c_float(self.gamma), C.c_float(self.alpha),
                                        1 if self.iou_branch_on else 0, _vp(self.acc), _vp(upstream), _vp(self.dcls),
                                        _vp(self.dbox), _vp(self.diou), _vp(self.pgrad), _st()), "fcos_loss_bwd")
        # Every layer's data- and weight-gradients (and, for the shared head / FPN convs, all three pyramid levels) depend only
        # on the BatchNorm backward before them: they go out as ONE grouped launch per layer.
        lv = range(3)
        if iou_on:
            for l in lv:
                hi, o = self.HI[l], self.lvl_off[l]
                self._chk(lib.drn_skinny_conv_bwd(_vp(self.diou[o:]), _vp(hi.data), C.c_int64(hi.plane_stride), hi.C, 0, F // 2,
                                                  B, self.Tl[l], 1, 1, _vp(p[h + "iou_scores.3.weight"]), _vp(self.dHI[l]), F // 2, 0,
                                                  _vp(grads[h + "iou_scores.3.weight"]), _st()), "iou3_bwd")
            self._bn_bwd([self._bn_job(self.iouc[l], p, grads, da=self.dHI[l]) for l in lv])
            self._group([self._wgrad_desc(self.iouc[l], self.MX[l], "iouc", l) for l in lv] +
                        [d for l in lv for d in self._dgrad_descs(self.iouc[l], self.wp["iouc"], self.dMX[l])])
            self._bn_bwd([self._bn_job(self.mix[l], p, grads, da=self.dMX[l]) for l in lv])
        self._chk(lib.drn_head_proj_bwd(C.byref(self._head_levels()), _vp(self.dcls), _vp(self.dbox), _vp(p[h + "cls_logits.weight"]),
                                        _vp(p[h + "bbox_pred.weight"]), _vp(grads[h + "cls_logits.weight"]),
                                        _vp(grads[h + "bbox_pred.weight"]), _st()), "head_proj_bwd")
        if iou_on:  # mix_fc: weight gradients + data gradients added onto the tower gradient (fcos.py:101 cat)
            self._group([self._wgrad_desc(self.mix[l], self.TW[l], "mix", l) for l in lv] +
                        [d for l in lv for d in self._dgrad_descs(self.mix[l], self.wp["mix"], self.dTW[l], mode=L.OUT_ADD)])
        self._bn_bwd([self._bn_job(self.tower[l], p, grads, da=self.dTW[l]) for l in lv])
        self._group([self._wgrad_desc(self.tower[l], self.Pf[l], "towers", l) for l in lv] +
                    [d for l in lv for d in self._dgrad_descs(self.tower[l], self.wp["towers"], self.dPf[l])])
        # FPN output convs
        self._bn_bwd([self._bn_job(self.layer[i], p, grads, da=self.dPf[i]) for i in lv])
        self._group([self._wgrad_desc(self.layer[i], self.I[i], "layer%d" % i) for i in lv] +
                    [d for i in lv for d in self._dgrad_descs(self.layer[i], self.wp["layer%d" % i], self.dI[i])])
        for i in (1, 2):  # backward of the top-down upsample-add chain (FPN.py:63-68)
            self._chk(lib.drn_pair_sum_add(_vp(self.dI[i]), _vp(self.dI[i - 1]), C.c_int64(B * self.Tl[i]), F, _st()),
                      "pair_sum_add")
        # FPN lateral convs
        self._bn_bwd([self._bn_job(self.inner[i], p, grads, da=self.dI[i]) for i in lv])
        self._group([self._wgrad_desc(self.inner[i], self.Cact[i], "inner%d" % i) for i in lv] +
                    [d for i in lv for d in self._dgrad_descs(self.inner[i], self.wp["inner%d" % i], self.dC[i])])

    def backward_propfc(self, grads, pair_clusters=0, chunk=None):
        """prop_fc weight gradient: [D x (B*T)] x [(B*T) x D], the largest contraction of the backward pass.  pair_clusters > 0
        confines the persistent kernel to that many SM pairs.  chunk = (i, n): only output rows [i*D/n, (i+1)*D/n) of the
        gradient (data parallel: the chunks are all-reduced as they complete; at D = 4096, n = 4 a chunk is 64
        tiles = one wave on 64 of the 74 SM pairs -- the whole gradient is 4 waves either way -- which leaves 20 SMs to NCCL)."""
        B = self.B
        if pair_clusters:
            _lib().drn_set_pair_clusters(pair_clusters)
        try:
            gw = grads["prop_fc.weight"]
            if chunk is None:
                self._gemm(L.GEMM_WGRAD, self.dP_pl.desc(), self.f_pl.desc(), B, self.T, self.D, M=self.D, out=gw, out_ld=self.D,
                           out_tap_stride=0)
            else:
                i, n = chunk
                rows = self.D // n
                assert rows * n == self.D and rows % 8 == 0
                # static schedule (no workspace): a chunk must stay on 64 SM pairs, the other 10 TPCs belong to the exchange kernel
                self._gemm(L.GEMM_WGRAD, self.dP_pl.desc(), self.f_pl.desc(), B, self.T, self.D, M=rows, a_c0=i * rows,
                           out=gw[i * rows:(i + 1) * rows], out_ld=self.D, out_tap_stride=0, engine=2, streamk=False)
        finally:
            if pair_clusters:
                _lib().drn_set_pair_clusters(0)
        self.launches_bwd = self.launches

    def _unpack_wgrads(self, grads, iou_on, which="all"):
        """Weight-gradient workspaces [slices][k][O][C] -> parameter layout [O][C][k] (sum of the K-split / level slices), one
        launch, after the scalar parameter gradients gathered by the loss kernel (tiny copies).  which = "early": head + FPN
        only; "late": backbone only (data-parallel schedule: two launches, an all-reduce starts in between)."""
        lib, h, F = _lib(), "fcos.head.", self.F
        items = []
        if which != "late":
            grads[h + "cls_logits.bias"].copy_(self.pgrad[0:1])
            grads[h + "bbox_pred.bias"].copy_(self.pgrad[1:3])
            if iou_on:
                grads[h + "iou_scores.3.bias"].copy_(self.pgrad[3:4])
            for l in range(3):
                grads[h + "scales.%d.scale" % l].copy_(self.pgrad[4 + l:5 + l])
            for i in range(3):
                items.append(self._unpack_item(grads["fpn.fpn_inner%d.0.weight" % (i + 1)], "inner%d" % i))
                items.append(self._unpack_item(grads["fpn.fpn_layer%d.0.weight" % (i + 1)], "layer%d" % i))
            items.append(self._unpack_item(grads[h + "cls_tower.0.weight"], "towers", 0))
            items.append(self._unpack_item(grads[h + "bbox_tower.0.weight"], "towers", F))
            if iou_on:
                items.append(self._unpack_item(grads[h + "mix_fc.0.weight"], "mix"))
                items.append(self._unpack_item(grads[h + "iou_scores.0.weight"], "iouc"))
        if which != "early":
            for i in range(3):
                items.append(self._unpack_item(grads["backbone_net.forward_conv%d.0.weight" % i], "conv%d" % i))
        arr = (L.PackItem * len(items))(*items)
        self._chk(lib.drn_unpack_conv_wgrads(len(items), arr, _st()), "unpack_conv_wgrads")

    def backward_tail(self, p, grads, unpack=None):
        """Second part of the backward: the gates and the query encoder -- a latency-bound chain of small kernels (~0.4 ms) that
        leaves most SMs idle, which is where a data-parallel run hides the all-reduce of everything the first part produced
        (model/main_model.py:_run_backward) and where the single-GPU schedule hides the weight-gradient unpacking
        (`unpack` = (iou_on,): side branch forked where the BPTT starts)."""
        lib, B = _lib(), self.B
        # gates: dW += dq^T cmd, db += colsum(dq), dcmd = dq W  -- nine small contractions, one launch (dcmd is zero-filled
        # together with dq / pgrad at the start of the backward)
        K = self.cmd_dim
        jobs = (L.SgemmJob * 9)()
        for i in range(3):
            n = self.qdim[i]
            dq, cmd, W = self.dq[i].data_ptr(), self.cmd[i].data_ptr(), p["qInput%d.weight" % i].data_ptr()
            jw, jb, jc = jobs[3 * i], jobs[3 * i + 1], jobs[3 * i + 2]
            jw.A, jw.sam, jw.sak, jw.B, jw.sbk, jw.sbn = dq, 1, n, cmd, K, 1
            jw.C, jw.ldc, jw.M, jw.N, jw.K = grads["qInput%d.weight" % i].data_ptr(), K, n, K, B
            jw.store = 1  # the only contribution to this gradient: plain stores instead of 4 M atomics
            jb.A, jb.sam, jb.sak, jb.B, jb.sbk, jb.sbn = dq, 1, n, self.one.data_ptr(), 0, 0
            jb.C, jb.ldc, jb.M, jb.N, jb.K = grads["qInput%d.bias" % i].data_ptr(), 1, n, 1, B
            jc.A, jc.sam, jc.sak, jc.B, jc.sbk, jc.sbn = dq, n, 1, W, K, 1
            jc.C, jc.ldc, jc.M, jc.N, jc.K = self.dcmd[i].data_ptr(), K, B, K, n
        self._chk(lib.drn_sgemm_batch(9, jobs, _st()), "gates_bwd")
        # query encoder backward (BPTT), gradients accumulated into the zeroed buffers
        qd = self._qe_desc(p, grads)
        nq = lib.drn_qe_launch_count(B, self.L, self.qe_H, 1)  # kernels enqueued inside drn_qe_backward
        if unpack is None:
            self._chk(lib.drn_qe_backward(C.byref(qd), _st()), "qe_backward")
            self.launches += nq - 1
        else:
            self._chk(lib.drn_qe_backward_part(C.byref(qd), 1, _st()), "qe_backward[1]")
            with self._fork():
                self._unpack_wgrads(grads, unpack[0])
            self._chk(lib.drn_qe_backward_part(C.byref(qd), 2, _st()), "qe_backward[2]")
            self._join()
            self.launches += nq - 2
        self.launches_bwd = self.launches
