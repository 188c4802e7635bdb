"""Why do a CUDA run and an oracle run diverge after ONE Adam step although their step-0 gradients agree to ~1e-2 rel-L2?
Adam's first update is lr * g / (|g| + 1e-8): sign-like.  For each gradient tensor at B=4, T=64 this prints the rel-L2 error,
and the RMS difference of the Adam-normalised updates u = g / (|g| + 1e-8), for (CUDA vs oracle) and for the control (oracle
with 2^-16-perturbed features vs oracle).  JSON on stdout."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from drn_b200 import spec as spec_mod  # noqa: E402
from drn_b200 import synthetic as S  # noqa: E402
from oracle import drn_oracle as O  # noqa: E402


def oracle_grads(sd, cfg, batch, perturb=0.0):
    leaf = {}
    g = torch.Generator().manual_seed(3)
    for k, v in sd.items():
        v = v.clone()
        if v.is_floating_point() and "running_" not in k:
            v.requires_grad_(not O.frozen_in_stage1(k))
        leaf[k] = v
    b = dict(batch)
    if perturb:
        f = batch["props_features"]
        b["props_features"] = f * (1 + perturb * (2 * torch.rand(f.shape, generator=g) - 1))
    _, ld, _ = O.forward(leaf, cfg, b, training=True)
    O.total_loss(ld, 1).backward()
    return {k: v.grad for k, v in leaf.items() if v.requires_grad and v.grad is not None}


def main():
    B, T = int(os.environ.get("PB", 4)), int(os.environ.get("PT", 64))
    torch.set_num_threads(os.cpu_count())
    from model.main_model import mainModel
    cfg = S.default_config(stage=1)
    sd = S.synth_state_dict(spec_mod.state_dict_spec(cfg))
    batch = S.synth_batch(B, T, max_len=8, embedding=sd["query_encoder.embedding.weight"], seed=S.SEED + 50)
    model = mainModel(1301, S.config_namespace(stage=1))
    model.load_state_dict(sd)
    for k, p in model.named_parameters():
        if O.frozen_in_stage1(k):
            p.requires_grad = False
    model = model.cuda().train()
    _, ld = model(batch["query_tokens"], batch["query_length"], batch["props_features"], batch["props_start_end"], batch["gt_start_end"], None, None)
    sum(ld.values()).backward()
    torch.cuda.synchronize()
    ref = oracle_grads(sd, cfg, batch)
    ctl = oracle_grads(sd, cfg, batch, perturb=2.0 ** -16)
    u = lambda g: g / (g.abs() + 1e-8)  # noqa: E731
    rows = {}
    tot = {"cuda": 0.0, "ctl": 0.0, "n": 0}
    for k, r in ref.items():
        g = dict(model.named_parameters())[k].grad
        if g is None:
            continue
        g = g.cpu()
        n = float(r.norm())
        du_c = float((u(g) - u(r)).pow(2).sum())
        du_p = float((u(ctl[k]) - u(r)).pow(2).sum())
        rows[k] = {"numel": r.numel(), "ref_norm": n, "rel_cuda": float((g - r).norm()) / max(n, 1e-30), "rel_ctl": float((ctl[k] - r).norm()) / max(n, 1e-30),
                   "rms_du_cuda": (du_c / r.numel()) ** 0.5, "rms_du_ctl": (du_p / r.numel()) ** 0.5,
                   "frac_abs_lt_1e-8_ref": float((r.abs() < 1e-8).float().mean()), "frac_exact_zero_ref": float((r == 0).float().mean()),
                   "frac_exact_zero_cuda": float((g == 0).float().mean()),
                   "sign_flips_cuda": float(((g * r) < 0).float().mean()), "sign_flips_ctl": float(((ctl[k] * r) < 0).float().mean())}
        tot["cuda"] += du_c
        tot["ctl"] += du_p
        tot["n"] += r.numel()
    print(json.dumps({"B": B, "T": T, "rms_du_all_cuda": (tot["cuda"] / tot["n"]) ** 0.5, "rms_du_all_ctl": (tot["ctl"] / tot["n"]) ** 0.5,
                      "per_tensor": rows}, indent=1))


if __name__ == "__main__":
    main()
