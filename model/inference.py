"""Eval-mode post-processing (reference model/inference.py:49-215): per level sigmoid -> threshold -> top-k -> decode ->
clamp, levels concatenated, fallback detection when nothing passes.  Operates on the small raw head outputs
([B, 1.75*T] values) after they are read back; the dense work happened in the kernels."""
import torch

DOWNSAMPLE = 32.0  # hard-coded in the reference (inference.py:45)


def postprocess(cls_raw, bbox, iou_raw, Tl, strides, cfg, B):
    """cls_raw [B*P], bbox [B*P,2], iou_raw [B*P] in level-major / sample / t order (CPU tensors)."""
    thr, top_n = cfg["fcos_inference_thr"], cfg["fcos_pre_nms_top_n"]
    first = cfg["is_first_stage"]
    results = []
    offs = [B * sum(Tl[:i]) for i in range(len(Tl))]
    for b in range(B):
        dets, scores, levels, locs = [], [], [], []
        for lvl, T in enumerate(Tl):
            s = float(strides[lvl])
            loc = torch.arange(0, T * s, step=s, dtype=torch.float32) + s / 2
            sl = slice(offs[lvl] + b * T, offs[lvl] + (b + 1) * T)
            c = torch.sigmoid(cls_raw[sl])
            cand = c > thr
            score = c if first else c * torch.sigmoid(iou_raw[sl])  # threshold applies to cls BEFORE the product
            idx = cand.nonzero().squeeze(1)
            sc = score[idx]
            k = min(int(cand.sum()), top_n)
            if idx.numel() > k:
                sc, top = sc.topk(k, sorted=False)
                idx = idx[top]
            reg = bbox[sl][idx]
            d = torch.stack([loc[idx] - reg[:, 0], loc[idx] + reg[:, 1]], dim=1) / DOWNSAMPLE
            d = d.clamp(min=0, max=1)
            d = d[(d[:, 1] - d[:, 0]) >= 0]
            if d.shape[0]:
                dets.append(d)
            if sc.numel():
                scores.append(torch.sqrt(sc))
                locs.append(loc[idx] / 32)
            levels.append([lvl] * d.shape[0])
        if not dets:  # inference.py:192-197
            results.append({"detections": torch.tensor([[0.0, 1.0]]), "labels": [], "scores": torch.tensor([1.0]),
                            "level": [[-1]], "locations": torch.tensor([0.5])})
        else:
            results.append({"detections": torch.cat(dets), "labels": [], "scores": torch.cat(scores), "level": levels,
                            "locations": torch.cat(locs)})
    return results
