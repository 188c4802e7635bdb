"""Recall@k with temporal NMS on the GPU (reference utils/evaluate_utils.py: `PostProcessRunner.run_evaluate(iou_topk_dict,
temporal_nms=True)` as main.py:450-454 calls it).  The reference walks every query in Python (stable sort, O(n^2) greedy NMS on
Python floats); here the whole result set is ONE launch of `drn_nms_recall` (one warp per query, IEEE double arithmetic in the
reference's operation order: bit-exact picks).  Thin marshalling only; no CPU fallback.

Two entry points:
  * `PostProcessRunner(raw_results).run_evaluate({"iou": [...], "topk": [...]}, temporal_nms=True)` -- the reference's
    interface on the `results_dict` main.py:411-447 builds (`{vid: [{"gt": [s, e], "node_predictions": [[s, e, score], ...]}]}`);
  * `recall_from_candidates(det, score, count, gt, ...)` -- straight from the device tensors `drn_postprocess` wrote
    ([B, levels, top_n] candidates), no host round trip: what the configs[4] bench times.
"""
import ctypes as C

import torch

from . import lib as L

MAX_N = 256


def _launch(det, score, count, gt, G, K, iou, topk, nms, empty_fallback, want_picks):
    dev = det.device
    if dev.type != "cuda":
        raise RuntimeError("drn_nms_recall runs on a B200 through libdrn_sm100.so only (no CPU fallback)")
    Q = gt.shape[0]
    topk_t = torch.tensor(list(topk), dtype=torch.int32, device=dev)
    picks = torch.empty(Q, G * K, dtype=torch.int32, device=dev) if want_picks else None
    npicks = torch.empty(Q, dtype=torch.int32, device=dev)
    hits = torch.empty(Q, len(topk), dtype=torch.int32, device=dev)
    correct = torch.zeros(len(topk), dtype=torch.int32, device=dev)
    overlap = float(iou) - 0.05  # evaluate_utils.py:152
    L.check(L.load().drn_nms_recall(L.ptr(det), L.ptr(score), L.ptr(count), L.ptr(gt), Q, G, K, 1 if nms else 0, C.c_double(overlap),
                                    C.c_double(float(iou)), L.ptr(topk_t), len(topk), 1 if empty_fallback else 0, L.ptr(picks),
                                    L.ptr(npicks), L.ptr(hits), L.ptr(correct), L.stream_ptr()), "nms_recall")
    return picks, npicks, hits, correct


def recall_from_candidates(det, score, count, gt, iou=0.5, topk=(1, 5), nms=True, want_picks=False, sync=True):
    """det [B, G, K, 2], score [B, G, K] fp32, count [B, G] int32 (device; `DensePath.postprocess()` output), gt [B, 2].
    Returns {"recall": {k: float}, "correct": int32 [ntopk] (device), "hits": [B, ntopk], "npicks": [B], "picks": [B, G*K] | None}.
    Reading "recall" is the only synchronisation (sync=False leaves it out: accumulate `correct` over batches on the device)."""
    B, G, K = score.shape
    gt64 = gt.to(device=det.device, dtype=torch.float64).contiguous()
    picks, npicks, hits, correct = _launch(det.contiguous(), score.contiguous(), count.to(torch.int32).contiguous(), gt64, G, K, iou,
                                           topk, nms, True, want_picks)
    out = {"correct": correct, "hits": hits, "npicks": npicks, "picks": picks}
    if sync:
        c = correct.tolist()
        out["recall"] = {k: c[i] / B for i, k in enumerate(topk)}
    return out


def nms_temporal(x1, x2, s, overlap, device="cuda"):
    """evaluate_utils.py:192-215 for ONE list of segments (test / debugging aid): returns the picks as a Python list."""
    n = len(s)
    if n == 0:
        return []
    if n > MAX_N:
        raise ValueError("at most %d segments per query" % MAX_N)
    dev = torch.device(device)
    det = torch.tensor(list(zip(x1, x2)), dtype=torch.float32, device=dev).view(1, n, 2)
    sc = torch.tensor(s, dtype=torch.float32, device=dev).view(1, n)
    cnt = torch.tensor([n], dtype=torch.int32, device=dev)
    gt = torch.zeros(1, 2, dtype=torch.float64, device=dev)
    topk = torch.tensor([1], dtype=torch.int32, device=dev)
    picks = torch.empty(1, n, dtype=torch.int32, device=dev)
    npicks = torch.empty(1, dtype=torch.int32, device=dev)
    L.check(L.load().drn_nms_recall(L.ptr(det), L.ptr(sc), L.ptr(cnt), L.ptr(gt), 1, 1, n, 1, C.c_double(float(overlap)), C.c_double(2.0),
                                    L.ptr(topk), 1, 0, L.ptr(picks), L.ptr(npicks), None, None, L.stream_ptr()), "nms_recall")
    return picks[0, :int(npicks[0])].tolist()


class PostProcessRunner:
    """The reference's metric object (utils/evaluate_utils.py:13-354) for the path main.py uses: no merging, temporal NMS.
    `do_merge` reads a hard-coded pickle of the authors' machine in the reference (evaluate_utils.py:58) and `do_viz` draws
    plotly figures: neither is part of the metric and both raise here."""

    def __init__(self, raw_results, device="cuda"):
        if not isinstance(raw_results, dict):
            import json
            raw_results = json.load(open(raw_results, "r"))
        self.raw_results = raw_results
        self.device = torch.device(device)

    def run_evaluate(self, iou_topk_dict, do_merge=False, update_score=False, score_weight=1.0, temporal_nms=False, viz_nms=True,
                     do_viz=""):
        assert isinstance(iou_topk_dict, dict)
        if do_merge or do_viz:
            raise NotImplementedError("PostProcessRunner: do_merge / do_viz are outside the metric path (evaluate_utils.py:58, 238-326)")
        ious, topks = iou_topk_dict["iou"], iou_topk_dict["topk"]
        queries = [qr for vid in self.raw_results.values() for qr in vid]
        Q = len(queries)
        n = max([len(qr["node_predictions"]) for qr in queries] + [1])
        if n > MAX_N:
            raise ValueError("at most %d predictions per query (got %d)" % (MAX_N, n))
        det = torch.zeros(Q, n, 2, dtype=torch.float32)
        score = torch.zeros(Q, n, dtype=torch.float32)
        count = torch.zeros(Q, dtype=torch.int32)
        gt = torch.zeros(Q, 2, dtype=torch.float64)
        for i, qr in enumerate(queries):
            p = torch.tensor(qr["node_predictions"], dtype=torch.float64).view(-1, 3)
            det[i, :p.shape[0]] = p[:, :2].float()   # main.py:425-430 produced these from fp32 tensors: exact
            score[i, :p.shape[0]] = p[:, 2].float()
            count[i] = p.shape[0]
            gt[i] = torch.tensor(qr["gt"], dtype=torch.float64)
        det, score, count, gt = (t.to(self.device) for t in (det, score, count, gt))
        accuracy_topks = []
        self.last = {}
        for iou in ious:
            _, npicks, hits, correct = _launch(det, score, count, gt, 1, n, iou, topks, temporal_nms, False, False)
            c = correct.tolist()
            self.last[iou] = {"hits": hits, "npicks": npicks}
            accuracy_topks += [c[i] / Q for i in range(len(topks))]  # evaluate_utils.py:340-347: iou-major, then topk
        return topks, accuracy_topks
