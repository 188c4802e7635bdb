"""Eval-mode post-processing (reference model/inference.py:49-215).

The arithmetic -- per level sigmoid -> threshold -> top-k -> decode -> clamp -> sqrt score -- runs in ONE fixed-shape kernel
for all (sample, level) pairs (drn_postprocess, drn_b200/csrc/head.cu).  What is left here is the reference's list assembly
on the [B, 3, top_n] result: concatenate the levels of every sample (inference.py:167-215) and substitute the fallback
detection when nothing passed the threshold (inference.py:192-197)."""
import torch


def assemble(det, score, loc, count):
    """det [B,L,K,2], score [B,L,K], loc [B,L,K], count [B,L] (CPU tensors from ONE device->host copy) -> the reference's
    per-sample dicts {detections [n,2], labels [], scores [n], level list[list[int]], locations [n]}."""
    B, nl = count.shape
    counts = count.tolist()
    results = []
    for b in range(B):
        dets, scores, locs, levels = [], [], [], []
        for lvl in range(nl):
            n = counts[b][lvl]
            if n:
                dets.append(det[b, lvl, :n])
                scores.append(score[b, lvl, :n])
                locs.append(loc[b, lvl, :n])
            levels.append([lvl] * n)
        if not dets:  # inference.py:192-197
            results.append({"detections": torch.tensor([[0.0, 1.0]]), "labels": [], "scores": torch.tensor([1.0]),
                            "level": [[-1]], "locations": torch.tensor([0.5])})
        else:
            results.append({"detections": torch.cat(dets), "labels": [], "scores": torch.cat(scores), "level": levels,
                            "locations": torch.cat(locs)})
    return results
