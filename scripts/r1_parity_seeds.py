"""R@1 / R@5 (tIoU 0.5, temporal NMS) parity over SEEDS (north_star: within +-0.2 pp of the reference).

For each seed (= its own sequence of training batches) three runs start from the same seeded weights and follow the reference
recipe (main.py:124-140,236-243: Adam lr 1e-3, clip_grad_norm 0.5, first stage) at the real DRN geometry T = 32, stopped in the
regime where R@1 is 50-80 % (not saturated):
    cuda     mainModel on libdrn_sm100
    oracle   the CPU restatement of the reference (pinned to it by tests/test_oracle_golden.py)
    control  the oracle with every weight (once) and every batch's features perturbed by 2^-16 relative -- the precision class of
             the CUDA path (three BF16 products), i.e. how far two runs of the REFERENCE ITSELF drift apart under round-off-sized noise
and are evaluated on the same held-out pairs (real Charades test queries).  Adam makes trajectories chaotic (tests/
test_multistep_gpu.py), so single runs differ by several pp in all three pairs; the claim tested here is statistical:
|mean(cuda) - mean(oracle)| is within the spread the control shows.  The metric runs on the GPU (drn_b200/metric.py) for the CUDA
arm and through oracle/metrics.py for the CPU arms.  Prints one JSON object.

    python scripts/r1_parity_seeds.py [--seeds 5] [--steps 60] [--eval-batches 16]"""
import argparse
import json
import os
import statistics
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
    ap.add_argument("--seeds", type=int, default=5)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--eval-batches", type=int, default=16)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--T", type=int, default=32)
    ap.add_argument("--signal-scale", type=float, default=2.0, help="amplitude of the query signal in the features (2.0 = SURVEY 8d; smaller = harder task)")
    ap.add_argument("--arms", default="cuda,oracle,control", help="which arms to run here (the CPU arms can run in the build container, the cuda arm on the GPU box)")
    ap.add_argument("--merge", nargs="*", default=None, help="JSON files of earlier --arms runs of the SAME settings to merge into the summary")
    a = ap.parse_args()
    arms = [x for x in a.arms.split(",") if x]
    torch.set_num_threads(os.cpu_count())
    cfg = S.default_config(stage=1)
    sd = S.synth_state_dict(spec_mod.state_dict_spec(cfg), glove=True)
    emb = sd["query_encoder.embedding.weight"]
    B, T = a.batch, a.T
    evalb = [S.synth_batch(B, T, max_len=10, embedding=emb, seed=S.SEED + 900000 + j, queries="charades", split="test", signal_scale=a.signal_scale)
             for j in range(a.eval_batches)]
    gts = [g for b in evalb for g in b["gt_start_end"].tolist()]

    def train_batches(seed):
        return [S.synth_batch(B, T, max_len=10, embedding=emb, seed=S.SEED + 100000 * (seed + 1) + i, queries="charades", signal_scale=a.signal_scale)
                for i in range(a.steps)]

    def run_cuda(batches):
        from model.main_model import mainModel
        model = mainModel(1301, S.config_namespace(stage=1))
        model.load_state_dict(sd)
        for k, p in model.named_parameters():
            if O.frozen_in_stage1(k):
                p.requires_grad = False
        model = model.cuda().train()
        opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=cfg["lr"])
        losses = []
        for b in batches:
            _, ld = model(b["query_tokens"], b["query_length"], b["props_features"], b["props_start_end"], b["gt_start_end"], None, None)
            loss = sum(ld.values())
            opt.zero_grad(set_to_none=True)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), cfg["clip_gradient"])
            opt.step()
            losses.append(float(loss))
        model.eval()
        res = []
        with torch.no_grad():
            for b in evalb:
                boxes, _ = model(b["query_tokens"], b["query_length"], b["props_features"], b["props_start_end"], b["gt_start_end"], None, None)
                res += boxes
        return M.recall_at(res, gts), losses

    def run_oracle(batches, perturb, pseed):
        g = torch.Generator().manual_seed(1000 + pseed)
        leaf = {}
        for k, v in sd.items():
            v = v.detach().clone()
            if v.is_floating_point() and "running_" not in k:
                if perturb:
                    v = v * (1 + perturb * (2 * torch.rand(v.shape, generator=g) - 1))
                v.requires_grad_(not O.frozen_in_stage1(k))
            leaf[k] = v
        params = [v for v in leaf.values() if v.is_floating_point() and v.requires_grad]
        opt = torch.optim.Adam(params, lr=cfg["lr"])
        losses = []
        for b in batches:
            if perturb:
                b = dict(b)
                f = b["props_features"]
                b["props_features"] = f * (1 + perturb * (2 * torch.rand(f.shape, generator=g) - 1))
            _, ld, newbuf = O.forward(leaf, cfg, b, training=True)
            loss = O.total_loss(ld, 1)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(params, cfg["clip_gradient"])
            opt.step()
            with torch.no_grad():
                for k, v in newbuf.items():
                    leaf[k] = v.detach().clone()
            losses.append(float(loss))
        res = []
        with torch.no_grad():
            for b in evalb:
                rb, _, _ = O.forward(leaf, cfg, b, training=False)
                res += rb
        return M.recall_at(res, gts), losses

    rows = [{"seed": seed} for seed in range(a.seeds)]
    for f in a.merge or []:
        prev = json.load(open(f))
        assert (prev["steps"], prev["B"], prev["T"], prev["seeds"], prev.get("signal_scale", 2.0)) == (a.steps, B, T, a.seeds, a.signal_scale), f
        for r, pr in zip(rows, prev["per_seed"]):
            r.update({k: v for k, v in pr.items() if k in ("cuda", "oracle", "control")})
    t0 = time.time()
    for seed in range(a.seeds):
        tb = train_batches(seed)
        for arm in arms:
            rec, losses = run_cuda(tb) if arm == "cuda" else run_oracle(tb, 0.0 if arm == "oracle" else 2.0 ** -16, seed)
            rows[seed][arm] = {"R@1": rec[1], "R@5": rec[5], "final_loss": losses[-1], "loss_step0": losses[0]}
            sys.stderr.write("seed %d %s: R@1 %.4f R@5 %.4f (%.0f s)\n" % (seed, arm, rec[1], rec[5], time.time() - t0))
    have = [arm for arm in ("cuda", "oracle", "control") if all(arm in r for r in rows)]
    mean = lambda arm, k: statistics.mean(r[arm][k] for r in rows)  # noqa: E731
    sdev = lambda arm, k: statistics.pstdev(r[arm][k] for r in rows)  # noqa: E731
    out = {"steps": a.steps, "B": B, "T": T, "seeds": a.seeds, "signal_scale": a.signal_scale, "eval_pairs": len(gts),
           "queries": "Charades-STA (train / held-out test split), GloVe-300", "arms": have, "per_seed": rows}
    for k in ("R@1", "R@5"):
        out[k] = {"mean_pct": {arm: 100 * mean(arm, k) for arm in have}, "std_over_seeds_pp": {arm: 100 * sdev(arm, k) for arm in have}}
        if "cuda" in have and "oracle" in have:
            out[k]["mean_cuda_minus_oracle_pp"] = 100 * (mean("cuda", k) - mean("oracle", k))
            out[k]["mean_abs_cuda_vs_oracle_pp"] = 100 * statistics.mean(abs(r["cuda"][k] - r["oracle"][k]) for r in rows)
        if "control" in have and "oracle" in have:
            out[k]["mean_control_minus_oracle_pp"] = 100 * (mean("control", k) - mean("oracle", k))
            out[k]["mean_abs_control_vs_oracle_pp"] = 100 * statistics.mean(abs(r["control"][k] - r["oracle"][k]) for r in rows)
    if len(have) == 3:
        d, c = abs(out["R@1"]["mean_cuda_minus_oracle_pp"]), abs(out["R@1"]["mean_control_minus_oracle_pp"])
        sem = 100 * (sdev("oracle", "R@1") ** 2 + sdev("cuda", "R@1") ** 2) ** 0.5 / max(a.seeds, 1) ** 0.5
        out["verdict"] = {"abs_mean_diff_pp": d, "control_abs_mean_diff_pp": c, "standard_error_of_the_difference_pp": sem,
                          "within_0.2pp": d <= 0.2, "within_control_or_2_sem": d <= max(0.2, c, 2 * sem)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
