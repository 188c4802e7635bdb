"""Pins oracle/drn_oracle.py against fixtures produced by the UNMODIFIED reference
(oracle/make_goldens.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from drn_b200 import spec as spec_mod
from drn_b200 import synthetic as S
from oracle import drn_oracle as O

CASES = list(S.GOLDEN_CASES)


def _summ(t):
    t = t.detach().to(torch.float64).reshape(-1)
    idx = torch.from_numpy(S.sample_indices(t.numel()))
    return np.concatenate([[float(t.norm()), float(t.sum())], t[idx].numpy()])


def _close(a, b, rtol, what, floor=1e-30):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max(), floor)
    err = np.abs(a - b).max() / scale
    assert err <= rtol, "%s: max err / max|ref| = %.3e > %.1e" % (what, err, rtol)


def test_spec_matches_golden_keys(golden_dir):
    g = np.load(os.path.join(golden_dir, "s3_train_b4_t32_crafted.npz"))
    names = {n for n, _ in spec_mod.state_dict_spec(S.default_config(stage=3))}
    for k in g.files:
        if k.startswith("grad/") or k.startswith("buf/"):
            assert k.split("/", 1)[1] in names, k


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name, golden_dir):
    torch.set_num_threads(os.cpu_count())
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    cfg0 = S.default_config(stage=S.GOLDEN_CASES[name][3])
    cfg, sd, batch, stage, training = S.golden_case(name, spec_mod.state_dict_spec(cfg0))
    leaf = {}
    for k, v in sd.items():
        v = v.clone()
        if v.is_floating_point() and "running_" not in k:
            v.requires_grad_(not (stage == 1 and O.frozen_in_stage1(k)))
        leaf[k] = v
    cap = {}
    boxes, loss_dict, new_buf = O.forward(leaf, cfg, batch, training=training, capture=cap)
    for k in ("loss_cls", "loss_reg", "loss_iou"):
        _close(loss_dict[k].detach().reshape(-1).numpy(), g["loss/" + k], 2e-5, k)
        assert str(loss_dict[k].dtype) == str(g["loss_dtype/" + k]), k
    for l in range(3):
        _close(cap["logits%d" % l].detach().numpy(), g["head/logits%d" % l], 2e-5, "logits%d" % l)
        _close(cap["bbox%d" % l].detach().numpy(), g["head/bbox%d" % l], 2e-5, "bbox%d" % l)
        _close(cap["iou%d" % l].detach().numpy(), g["head/iou%d" % l], 2e-5, "iou%d" % l)
    cap["P_btd"] = cap["P"].permute(0, 2, 1)
    for k in g.files:
        if k.startswith("cap/"):
            _close(_summ(cap[k[4:]]), g[k], 2e-5, k)
    if training:
        loss = O.total_loss(loss_dict, stage)
        loss.backward()
        n = 0
        for k in g.files:
            if k.startswith("grad/"):
                p = leaf[k[5:]]
                assert p.grad is not None, k
                mine = _summ(p.grad)
                if g[k][0] < 1e-6:
                    # mathematically zero gradients (conv bias in front of train-mode BN, softmax shift):
                    # only round-off on either side
                    assert mine[0] < 1e-5, k
                else:
                    _close(mine, g[k], 5e-4, k)
                n += 1
            elif k.startswith("buf/"):
                key = k[4:]
                _close(_summ(new_buf.get(key, sd[key]).to(torch.float64)), g[k], 2e-5, k)
        assert n > 50
        # parameters the reference leaves without a gradient must have none here either
        for k, v in leaf.items():
            if v.requires_grad and ("grad/" + k) not in g.files:
                assert v.grad is None or float(v.grad.abs().max()) == 0.0, k
    else:
        for b, d in enumerate(boxes):
            ref_det = g["det/%d/detections" % b]
            assert d["detections"].shape == ref_det.shape
            # top-k with sorted=False has no guaranteed order: compare as sets ordered by (score, start)
            def order(det, sc):
                return np.lexsort((det[:, 0], sc))
            o1 = order(d["detections"].numpy(), d["scores"].numpy())
            o2 = order(ref_det, g["det/%d/scores" % b])
            _close(d["detections"].numpy()[o1], ref_det[o2], 2e-5, "det")
            _close(d["scores"].numpy()[o1], g["det/%d/scores" % b][o2], 2e-5, "scores")
            _close(d["locations"].numpy()[o1], g["det/%d/locations" % b][o2], 2e-5, "locations")
            lv = np.array([x for l in d["level"] for x in l], dtype=np.int64)
            assert (np.sort(lv) == np.sort(g["det/%d/level" % b])).all()


def test_metric_restatement_matches_reference_golden(golden_dir):
    """oracle/metrics.py (temporal NMS + unclamped IoU of utils/evaluate_utils.py:192-236) against outputs of the reference's own
    functions on seeded random segments (oracle/make_metric_goldens.py), and recall@k on a hand-checkable case."""
    import json
    from oracle import metrics as M
    cases = json.load(open(os.path.join(golden_dir, "metric_nms.json")))
    assert len(cases) == 100
    for c in cases:
        assert M.nms_temporal(c["x1"], c["x2"], c["s"], 0.45) == c["picks"]
        for a, b, ref in zip(c["x1"], c["x2"], c["iou"]):
            assert abs(M.calculate_iou(c["gt"], (a, b)) - ref) < 1e-12
    res = [{"detections": torch.tensor([[0.0, 0.5], [0.5, 1.0], [0.0, 0.0]]), "scores": torch.tensor([0.9, 0.8, 0.99])},
           {"detections": torch.tensor([[0.0, 1.0]]), "scores": torch.tensor([1.0])}]
    r = M.recall_at(res, [(0.5, 1.0), (0.1, 0.2)])  # zero-duration detection dropped; 2nd pick hits query 0; query 1 misses
    assert r[1] == 0.0 and r[5] == 0.5


def _pool_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "pool_props.npz"))
    B = len(g["props_num"])
    feats = [g["feat_%d" % k] for k in range(B)]
    ps = [g["p_start_%d" % k] for k in range(B)]
    pe = [g["p_end_%d" % k] for k in range(B)]
    return g, feats, ps, pe


def test_pooling_oracle_matches_reference_golden(golden_dir):
    """oracle/pooling.py (dataset.py:105-155,180-206) against the output of the reference's own CharadesSTA.get_data +
    collate_data on synthetic feature files (oracle/make_pool_goldens.py): bit-exact (integer index arithmetic + max)."""
    from oracle import pooling as P
    g, feats, ps, pe = _pool_golden(golden_dir)
    out, pse = P.pool_and_pad(feats, ps, pe, [int(x) for x in g["num_frames"]], window=16, interval=8)
    assert out.shape == g["props_features"].shape and np.array_equal(out, g["props_features"])
    assert np.array_equal(pse, g["props_s_e"])
    # the fixtures exercise both branches (single window / range) and the clamp to the last available window
    lo_hi = [P.window_range(float(s), int(e), feats[k].shape[0], 16, 8) for k in range(len(feats)) for s, e in zip(ps[k], pe[k])]
    assert any(lo == hi for lo, hi in lo_hi) and any(hi > lo for lo, hi in lo_hi)
    assert any(hi == feats[k].shape[0] - 1 for k in range(len(feats)) for (lo, hi) in
               [P.window_range(float(ps[k][-1]), int(pe[k][-1]), feats[k].shape[0], 16, 8)])
