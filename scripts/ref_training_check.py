"""Build-container check (CPU, needs /root/reference): the UNMODIFIED reference and the oracle restatement trained side by side
with the reference recipe (Adam 1e-3, clip 0.5, first stage; main.py:124-140,236-243) from the same seeded weights on the same
batches, then evaluated on the same held-out pairs: per-step losses and R@1 / R@5.  Pins the oracle's TRAINING behaviour (not
just one forward/backward) to the reference, so that scripts/r1_parity_seeds.py -- CUDA vs oracle on the GPU box, where the
reference package cannot be a training arm next to `model/` -- is a statement about the reference.  Prints one JSON object.

    python scripts/ref_training_check.py [--steps 60] [--eval-batches 8] > profiles/r02_ref_training_check.json"""
import argparse
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from drn_b200 import synthetic as S  # noqa: E402
from oracle import drn_oracle as O  # noqa: E402
from oracle import metrics as M  # noqa: E402
from oracle import ref_loader  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--eval-batches", type=int, default=8)
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    os.chdir(tempfile.mkdtemp())  # the reference's eval forward writes ./total_points.pkl (fcos.py:182)
    cfg = S.default_config(stage=1)
    model = ref_loader.build_reference_model(cfg)
    sd = S.synth_state_dict([(k, tuple(v.shape)) for k, v in model.state_dict().items()], glove=True)
    model.load_state_dict(sd)
    emb = sd["query_encoder.embedding.weight"]
    B, T = 32, 32
    tb = [S.synth_batch(B, T, max_len=10, embedding=emb, seed=S.SEED + 100000 + i, queries="charades") for i in range(a.steps)]
    evalb = [S.synth_batch(B, T, max_len=10, embedding=emb, seed=S.SEED + 900000 + j, queries="charades", split="test") for j in range(a.eval_batches)]
    gts = [g for b in evalb for g in b["gt_start_end"].tolist()]
    # ---- reference ----
    model.train()
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=cfg["lr"])
    ref_loss = []
    for b in tb:
        _, ld = model(b["query_tokens"], b["query_length"], b["props_features"], b["props_start_end"], b["gt_start_end"], None, None)
        loss = sum(v for v in ld.values())
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), cfg["clip_gradient"])
        opt.step()
        ref_loss.append(float(loss.detach()))
    model.eval()
    res_ref = []
    with torch.no_grad():
        for b in evalb:
            boxes, _ = model(b["query_tokens"], b["query_length"], b["props_features"], b["props_start_end"], b["gt_start_end"], None, None)
            res_ref += boxes
    # ---- oracle ----
    leaf = {}
    for k, v in sd.items():
        v = v.detach().clone()
        if v.is_floating_point() and "running_" not in k:
            v.requires_grad_(not O.frozen_in_stage1(k))
        leaf[k] = v
    params = [v for v in leaf.values() if v.is_floating_point() and v.requires_grad]
    opt = torch.optim.Adam(params, lr=cfg["lr"])
    or_loss = []
    for b in tb:
        _, ld, newbuf = O.forward(leaf, cfg, b, training=True)
        loss = O.total_loss(ld, 1)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, cfg["clip_gradient"])
        opt.step()
        with torch.no_grad():
            for k, v in newbuf.items():
                leaf[k] = v.detach().clone()
        or_loss.append(float(loss.detach()))
    res_or = []
    with torch.no_grad():
        for b in evalb:
            rb, _, _ = O.forward(leaf, cfg, b, training=False)
            res_or += rb
    rr, ro = M.recall_at(res_ref, gts), M.recall_at(res_or, gts)
    gaps = [abs(x - y) / abs(x) for x, y in zip(ref_loss, or_loss)]
    print(json.dumps({"steps": a.steps, "B": B, "T": T, "eval_pairs": len(gts),
                      "reference": {"R@1": rr[1], "R@5": rr[5], "loss_first5": ref_loss[:5], "loss_last": ref_loss[-1]},
                      "oracle": {"R@1": ro[1], "R@5": ro[5], "loss_first5": or_loss[:5], "loss_last": or_loss[-1]},
                      "rel_loss_gap": {"step0": gaps[0], "step1": gaps[1], "step5": gaps[min(5, len(gaps) - 1)], "max": max(gaps), "last": gaps[-1]},
                      "R@1_diff_pp": 100 * (ro[1] - rr[1]), "R@5_diff_pp": 100 * (ro[5] - rr[5])}))


if __name__ == "__main__":
    main()
