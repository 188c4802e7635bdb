"""Accuracy parity on synthetic labels (north_star: R@1 IoU=0.5 within +-0.2 pp of the reference).

Trains the CUDA path (mainModel on libdrn_sm100) and the CPU oracle (the pinned restatement of the reference) from the same
seeded weights on the same seeded synthetic batches with the reference's recipe (main.py:124-140, 236-243: Adam lr 1e-3,
clip_grad_norm 0.5, first stage), at the real DRN geometry T = 32, then evaluates both on held-out synthetic pairs with the
reference's metric (oracle/metrics.py = utils/evaluate_utils.py: score sort, temporal NMS 0.45, R@{1,5} at tIoU 0.5).

    python scripts/r1_parity.py [--steps 150] [--eval-batches 16] [--batch 32] [--T 32]
Prints one JSON line; exit status 1 if |R@1 difference| > 0.2 pp... is NOT enforced here (two float trajectories of a
non-convex optimisation diverge; the line reports the numbers and the loss curves)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from drn_b200 import spec as spec_mod  # noqa: E402
from drn_b200 import synthetic as S  # noqa: E402
from oracle import drn_oracle as O  # noqa: E402
from oracle import metrics as M  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=150)
    ap.add_argument("--eval-batches", type=int, default=16)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--T", type=int, default=32)
    ap.add_argument("--no-oracle", action="store_true")
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    from model.main_model import mainModel
    cfg = S.default_config(stage=1)
    sd = S.synth_state_dict(spec_mod.state_dict_spec(cfg))
    emb = sd["query_encoder.embedding.weight"]
    B, T = a.batch, a.T

    def batch(i):
        return S.synth_batch(B, T, max_len=10, embedding=emb, seed=S.SEED + 1000 + i)

    # ---- CUDA path ----
    model = mainModel(1301, S.config_namespace(stage=1))
    model.load_state_dict(sd)
    for k, p in model.named_parameters():
        if O.frozen_in_stage1(k):
            p.requires_grad = False
    model = model.cuda().train()
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=cfg["lr"])
    ours_loss = []
    t0 = time.time()
    for i in range(a.steps):
        b = batch(i)
        _, ld = model(b["query_tokens"], b["query_length"], b["props_features"], b["props_start_end"], b["gt_start_end"], None, None)
        loss = sum(ld.values())
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), cfg["clip_gradient"])
        opt.step()
        ours_loss.append(float(loss))
    t_ours = time.time() - t0

    # ---- CPU oracle ----
    ref_loss, leaf = [], None
    if not a.no_oracle:
        leaf = {}
        for k, v in sd.items():
            v = v.detach().clone()
            if v.is_floating_point() and "running_" not in k:
                v.requires_grad_(not O.frozen_in_stage1(k))
            leaf[k] = v
        params = [v for v in leaf.values() if v.requires_grad]
        ropt = torch.optim.Adam(params, lr=cfg["lr"])
        t0 = time.time()
        for i in range(a.steps):
            b = batch(i)
            _, ld, newbuf = O.forward(leaf, cfg, b, training=True)
            loss = O.total_loss(ld, 1)
            ropt.zero_grad(set_to_none=True)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(params, cfg["clip_gradient"])
            ropt.step()
            with torch.no_grad():
                for k, v in newbuf.items():
                    leaf[k] = v.detach().clone()
            ref_loss.append(float(loss))
        t_ref = time.time() - t0

    # ---- held-out evaluation ----
    model.eval()
    res_o, res_r, gts = [], [], []
    with torch.no_grad():
        for j in range(a.eval_batches):
            b = S.synth_batch(B, T, max_len=10, embedding=emb, seed=S.SEED + 900000 + j)
            boxes, _ = model(b["query_tokens"], b["query_length"], b["props_features"], b["props_start_end"], b["gt_start_end"], None, None)
            res_o += boxes
            gts += b["gt_start_end"].tolist()
            if leaf is not None:
                rb, _, _ = O.forward(leaf, cfg, b, training=False)
                res_r += rb
    # ---- same-checkpoint parity: each side's trained weights evaluated by the OTHER implementation ----
    cross = {}
    if leaf is not None:
        ours_sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
        res_x = []
        with torch.no_grad():
            for j in range(a.eval_batches):
                b = S.synth_batch(B, T, max_len=10, embedding=emb, seed=S.SEED + 900000 + j)
                rb, _, _ = O.forward(ours_sd, cfg, b, training=False)
                res_x += rb
        cross["cuda_trained_weights_on_oracle"] = M.recall_at(res_x, gts)
        m2 = mainModel(1301, S.config_namespace(stage=1))
        m2.load_state_dict({k: v.detach() for k, v in leaf.items()})
        m2 = m2.cuda().eval()
        res_y = []
        with torch.no_grad():
            for j in range(a.eval_batches):
                b = S.synth_batch(B, T, max_len=10, embedding=emb, seed=S.SEED + 900000 + j)
                boxes, _ = m2(b["query_tokens"], b["query_length"], b["props_features"], b["props_start_end"], b["gt_start_end"], None, None)
                res_y += boxes
        cross["oracle_trained_weights_on_cuda"] = M.recall_at(res_y, gts)
    ro = M.recall_at(res_o, gts)
    out = {"steps": a.steps, "B": B, "T": T, "eval_pairs": len(gts), "ours": {"R@1": ro[1], "R@5": ro[5], "final_loss": ours_loss[-1],
           "loss_first_last10": [sum(ours_loss[:10]) / 10, sum(ours_loss[-10:]) / 10], "train_s": round(t_ours, 2)}}
    if leaf is not None:
        rr = M.recall_at(res_r, gts)
        out["oracle_cpu"] = {"R@1": rr[1], "R@5": rr[5], "final_loss": ref_loss[-1],
                             "loss_first_last10": [sum(ref_loss[:10]) / 10, sum(ref_loss[-10:]) / 10], "train_s": round(t_ref, 2)}
        out["independent_training_R@1_diff_pp"] = 100 * (ro[1] - rr[1])
        out["independent_training_R@5_diff_pp"] = 100 * (ro[5] - rr[5])
        out["same_weights"] = {
            "cuda_trained: R@1 cuda vs oracle": [ro[1], cross["cuda_trained_weights_on_oracle"][1]],
            "cuda_trained: R@5 cuda vs oracle": [ro[5], cross["cuda_trained_weights_on_oracle"][5]],
            "oracle_trained: R@1 cuda vs oracle": [cross["oracle_trained_weights_on_cuda"][1], rr[1]],
            "oracle_trained: R@5 cuda vs oracle": [cross["oracle_trained_weights_on_cuda"][5], rr[5]],
            "max_R@1_diff_pp": 100 * max(abs(ro[1] - cross["cuda_trained_weights_on_oracle"][1]),
                                         abs(cross["oracle_trained_weights_on_cuda"][1] - rr[1]))}
        out["loss_head"] = {"ours": [round(x, 5) for x in ours_loss[:8]], "oracle_cpu": [round(x, 5) for x in ref_loss[:8]]}
        out["rel_loss_gap_step0"] = abs(ours_loss[0] - ref_loss[0]) / abs(ref_loss[0])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
