"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's recall metric (utils/evaluate_utils.py).

R@k at a tIoU threshold with greedy temporal NMS, as `PostProcessRunner.run_evaluate(iou_topk_dict={"iou": [0.5],
"topk": [1, 5]}, temporal_nms=True)` computes it for main.py:450-454:
  * predictions sorted by score, descending (evaluate_utils.py:97, `_postprocess_raw_results_no_merge`);
  * `nms_temporal(starts, ends, scores, iou - 0.05)` (evaluate_utils.py:152, 192-215): repeatedly pick the best remaining
    segment and drop every other whose overlap inter / (len_i + len_j - inter) exceeds the threshold;
  * a query is a hit if any of the first k picks has `calculate_IoU(gt, pred) >= iou` with
    IoU = (min(e) - max(s)) / (max(e) - min(s)), unclamped (evaluate_utils.py:175-184, 232-236).
Zero-duration detections make the reference's NMS divide by zero (hard part 3: at T > 32 both ends clamp
to 1); they are dropped here, identically for both sides of a comparison.  Used by tests/ and scripts/r1_parity.py only.
"""


def nms_temporal(x1, x2, s, overlap):
    """evaluate_utils.py:192-215 (same tie behaviour: stable sort by score, best last)."""
    pick = []
    if not x1:
        return pick
    length = [b - a for a, b in zip(x1, x2)]
    order = [i for i, _ in sorted(enumerate(s), key=lambda x: x[1])]
    while order:
        i = order[-1]
        pick.append(i)
        keep = []
        for j in order[:-1]:
            inter = max(0.0, min(x2[i], x2[j]) - max(x1[i], x1[j]))
            if inter / (length[i] + length[j] - inter) <= overlap:
                keep.append(j)
        order = keep
    return pick


def calculate_iou(a, b):
    """evaluate_utils.py:232-236 (no clamp: disjoint segments give a negative value)."""
    return 1.0 * (min(a[1], b[1]) - max(a[0], b[0])) / (max(a[1], b[1]) - min(a[0], b[0]))


def recall_at(results, gts, topks=(1, 5), iou=0.5, nms=True):
    """results: per query dict with `detections` [n,2] and `scores` [n] (mainModel eval output / oracle.postprocess);
    gts: per query (start, end).  Returns {k: recall}."""
    hits = {k: 0 for k in topks}
    for res, gt in zip(results, gts):
        det = res["detections"].detach().cpu().tolist()
        sc = res["scores"].detach().cpu().tolist()
        preds = [(d[0], d[1], s) for d, s in zip(det, sc) if d[1] - d[0] > 0]
        preds.sort(key=lambda x: x[2], reverse=True)
        starts, ends, scores = [p[0] for p in preds], [p[1] for p in preds], [p[2] for p in preds]
        picks = nms_temporal(starts, ends, scores, iou - 0.05) if nms else list(range(len(preds)))
        for k in topks:
            for idx in picks[:k]:
                if calculate_iou((float(gt[0]), float(gt[1])), (starts[idx], ends[idx])) >= iou:
                    hits[k] += 1
                    break
    n = max(len(gts), 1)
    return {k: hits[k] / n for k in topks}
