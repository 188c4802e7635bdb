"""Multi-step parity of the CUDA path with the CPU oracle: several optimizer steps from the same initial weights on the same
seeded batches, with the reference's training loop (main.py:216-244: forward, loss, backward, clip_grad_norm_ over
model.parameters(), optimizer.step(), optimizer.zero_grad()).  What one-step tests cannot see: CUDA-graph replay after the
weights changed (they are re-packed inside the graph), BatchNorm running statistics and num_batches_tracked over steps, and
the stage-2 quirk that parameters outside the optimizer keep ACCUMULATING gradients (main.py:132-134,239).

Adam turns a gradient into +-lr wherever |g| >> 1e-8, so two float implementations of the same gradient diverge after ONE step
on every element whose gradient is round-off noise.  The bound is therefore calibrated by a CONTROL: the oracle against itself
with its input features perturbed by 2^-16 relative.  SGD (update proportional to the gradient) has no such amplification and
is held to a plain tolerance.

The control perturbs what the CUDA path perturbs: three BF16 products carry ~16 mantissa bits per operand, i.e. every layer sees
its weights and activations at ~2^-17 relative precision; scripts/bwd_debug.py shows the consequence at B=4, T=64 -- conv
outputs 5e-5 off after 7 layers, which flips ONE ReLU mask (|bn(y)| = 2.4e-5) among the 32 k elements of the level-3 bbox tower
and moves that layer's gradient by 2e-2 while every other element agrees to 1e-5.  A features-only perturbation flips 10x fewer
masks (scripts/adam_divergence_probe.py), so the control perturbs weights AND features by 2^-16."""
import os

import pytest
import torch

from drn_b200 import spec as spec_mod
from drn_b200 import synthetic as S
from oracle import drn_oracle as O

pytestmark = pytest.mark.gpu


def _batches(B, T, n, emb, crafted_sd=None):
    out = []
    for i in range(n):
        b = S.synth_batch(B, T, max_len=8, embedding=emb, seed=S.SEED + 50 + i)
        if crafted_sd is not None:
            _, b = S.craft_stage23(crafted_sd, b)
        out.append(b)
    return out


def _make_opt(kind, params, lr):
    return torch.optim.Adam(params, lr=lr) if kind == "adam" else torch.optim.SGD(params, lr=lr)


def _train_cuda(sd, stage, batches, kind, lr, opt_filter=None, clip=0.5):
    from model.main_model import mainModel
    model = mainModel(1301, S.config_namespace(stage=stage))
    model.load_state_dict(sd)
    if stage == 1:
        for k, p in model.named_parameters():
            if O.frozen_in_stage1(k):
                p.requires_grad = False
    model = model.cuda().train()
    named = [(k, p) for k, p in model.named_parameters() if p.requires_grad and (opt_filter is None or opt_filter(k))]
    opt = _make_opt(kind, [p for _, p in named], lr)
    opt.zero_grad()
    losses, norms = [], []
    for b in batches:
        _, ld = model(b["query_tokens"], b["query_length"], b["props_features"], b["props_start_end"], b["gt_start_end"], None, None)
        loss = ld["loss_iou"] if stage == 2 else sum(ld.values())
        losses.append(float(loss))
        if loss != 0:
            loss.backward()
        norms.append(float(torch.nn.utils.clip_grad_norm_(model.parameters(), clip)))
        opt.step()
        opt.zero_grad()  # only the optimizer's parameters: the rest keep their .grad (stage 2)
    torch.cuda.synchronize()
    return model, losses, norms


def _train_oracle(sd, cfg, stage, batches, kind, lr, opt_filter=None, clip=0.5, perturb=0.0, zero_all=False):
    leaf = {}
    g = torch.Generator().manual_seed(3)
    for k, v in sd.items():
        v = v.detach().clone()
        if v.is_floating_point() and "running_" not in k:
            if perturb:  # the control: every weight (once) and every batch's features perturbed by `perturb` relative
                v = v * (1 + perturb * (2 * torch.rand(v.shape, generator=g) - 1))
            v.requires_grad_(not (stage == 1 and O.frozen_in_stage1(k)))
        leaf[k] = v
    every = [v for v in leaf.values() if v.is_floating_point() and v.requires_grad]
    opt = _make_opt(kind, [v for k, v in leaf.items() if v.is_floating_point() and v.requires_grad and (opt_filter is None or opt_filter(k))], lr)
    losses, norms = [], []
    for b in batches:
        if perturb:
            b = dict(b)
            f = b["props_features"]
            b["props_features"] = f * (1 + perturb * (2 * torch.rand(f.shape, generator=g) - 1))
        _, ld, newbuf = O.forward(leaf, cfg, b, training=True)
        loss = O.total_loss(ld, stage)
        losses.append(float(loss))
        if loss.requires_grad:
            loss.backward()
        norms.append(float(torch.nn.utils.clip_grad_norm_(every, clip)))
        opt.step()
        opt.zero_grad()
        if zero_all:
            for v in every:
                v.grad = None
        with torch.no_grad():
            for k, v in newbuf.items():
                leaf[k] = v.detach().clone()
    return leaf, losses, norms


def _rel(a, b):
    return float((a.double() - b.double()).norm() / max(float(b.double().norm()), 1e-30))


# Adam at lr 1e-5 (the reference's stage-2 rate): at its stage-1 rate of 1e-3 every one of the 38 M parameters moves by +-1e-3 in
# the first step (the loss goes 2.60 -> 1.42 -> 3.2 at B=4) and the 0.4 % of gradient elements whose sign is round-off decide the
# trajectory: CUDA, oracle and perturbed oracle are 1e-1 apart in loss after three steps, all three pairs alike -- that regime is
# compared statistically over seeds by scripts/r1_parity.py, not step by step.
@pytest.mark.parametrize("kind,lr", [("sgd", 1e-2), ("adam", 1e-5)])
def test_five_optimizer_steps_stage1(kind, lr):
    torch.set_num_threads(os.cpu_count())
    cfg = S.default_config(stage=1)
    sd = S.synth_state_dict(spec_mod.state_dict_spec(cfg))
    batches = _batches(4, 64, 5, sd["query_encoder.embedding.weight"])
    model, lc, nc = _train_cuda(sd, 1, batches, kind, lr)
    ref, lo, no = _train_oracle(sd, cfg, 1, batches, kind, lr)
    ctl, lp, npert = _train_oracle(sd, cfg, 1, batches, kind, lr, perturb=2.0 ** -16)
    msd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    # step 0 is a pure forward/backward comparison
    assert abs(lc[0] - lo[0]) <= 1e-4 * abs(lo[0]) and abs(nc[0] - no[0]) <= 2e-3 * no[0]
    ctl_gap = 0.0
    for i in range(5):
        ctl_gap = max(ctl_gap, abs(lp[i] - lo[i]))  # the control's divergence so far
        bound = (1e-3 if kind == "sgd" else 2e-4) * abs(lo[i]) + 4.0 * ctl_gap
        assert abs(lc[i] - lo[i]) <= bound, "step %d: loss %.6f vs oracle %.6f (control gap %.2e); losses %s | %s | %s" % (
            i, lc[i], lo[i], ctl_gap, lc, lo, lp)
    # BatchNorm: five updates of every running statistic, num_batches_tracked exact (shared head modules: 3 levels x 5 steps)
    worst_w = worst_ctl = 0.0
    tot_e = tot_c = tot_u = 0.0
    for k, v in ref.items():
        if k.endswith("num_batches_tracked"):
            assert int(msd[k]) == int(v), k
            assert int(v) == (15 if k.startswith("fcos.head.") else 5), k
        elif "running_" in k:
            assert _rel(msd[k], v) <= 2e-4 + 4.0 * _rel(ctl[k], v), k
        elif v.is_floating_point() and float((v - sd[k]).norm()) > 0:
            upd = float((v.detach() - sd[k]).double().norm())  # size of the 5-step update of this tensor
            e = float((msd[k].double() - v.detach().double()).norm()) / upd
            c = float((ctl[k].detach().double() - v.detach().double()).norm()) / upd
            worst_w, worst_ctl = max(worst_w, e), max(worst_ctl, c)
            tot_e, tot_c, tot_u = tot_e + (e * upd) ** 2, tot_c + (c * upd) ** 2, tot_u + upd ** 2
            # per tensor: SGD is proportional to the gradient (plain bound); under Adam both e and c are chaotic, so the
            # per-tensor bound is loose and the aggregate below carries the comparison
            assert e <= (5e-3 + 4.0 * c if kind == "sgd" else 2e-2 + 10.0 * c), "%s: update error %.2e (control %.2e)" % (k, e, c)
    E, Cc = (tot_e / tot_u) ** 0.5, (tot_c / tot_u) ** 0.5
    print("five steps (%s): update error over all tensors %.2e (control %.2e); worst tensor %.2e (control %.2e)" % (kind, E, Cc, worst_w, worst_ctl))
    assert E <= 3.0 * Cc + 2e-3, "all tensors: update error %.2e vs control %.2e" % (E, Cc)


def test_stage2_stale_gradients_accumulate_like_the_reference():
    """Stage 2 (main.py:132-134): only mix_fc + iou_scores are in the optimizer, loss = loss_iou; every other parameter keeps
    requires_grad=True, is never zeroed and so ACCUMULATES its gradient step after step, inflating clip_grad_norm_'s total norm
    (main.py:239).  Three steps: the norms returned by clip_grad_norm_, the accumulated stale gradients and the optimized
    weights must follow the oracle."""
    torch.set_num_threads(os.cpu_count())
    cfg = S.default_config(stage=2)
    sd0 = S.synth_state_dict(spec_mod.state_dict_spec(cfg))
    batches = _batches(4, 32, 3, sd0["query_encoder.embedding.weight"], crafted_sd=sd0)
    sd, _ = S.craft_stage23(sd0, batches[0])
    flt = lambda k: "iou_scores" in k or "mix_fc" in k  # noqa: E731
    model, lc, nc = _train_cuda(sd, 2, batches, "sgd", 1e-3, opt_filter=flt)
    ref, lo, no = _train_oracle(sd, cfg, 2, batches, "sgd", 1e-3, opt_filter=flt)
    assert all(x > 0 for x in lo), "crafted batches must keep the IoU branch live"
    for i in range(3):
        assert abs(lc[i] - lo[i]) <= 1e-3 * abs(lo[i]), (i, lc[i], lo[i])
        assert abs(nc[i] - no[i]) <= 5e-3 * no[i], (i, nc[i], no[i])
    # the quirk is live: with every gradient zeroed each step (what the reference does NOT do) the clipped norms differ
    _, _, nz = _train_oracle(sd, cfg, 2, batches, "sgd", 1e-3, opt_filter=flt, zero_all=True)
    assert abs(nz[0] - no[0]) <= 1e-6 * no[0] and abs(nz[2] - no[2]) > 5e-3 * no[2], (nz, no)
    params = dict(model.named_parameters())
    ctl, _, _ = _train_oracle(sd, cfg, 2, batches, "sgd", 1e-3, opt_filter=flt, perturb=2.0 ** -16)
    for k in ("fcos.head.bbox_pred.weight", "fcos.head.bbox_tower.0.weight", "fpn.fpn_layer1.0.weight"):
        g, r = params[k].grad.cpu(), ref[k].grad  # accumulated over 3 steps, clipped in place each step
        # small-case bound of tests/test_model_gpu.py (one ReLU-mask flip is visible in a B=4, T=32 tensor)
        assert r is not None and _rel(g, r) <= 2e-3 + 8.0 * _rel(ctl[k].grad, r), (k, _rel(g, r), _rel(ctl[k].grad, r))
    for k in ("fcos.head.iou_scores.0.weight", "fcos.head.mix_fc.0.weight"):
        upd = float((ref[k].detach() - sd[k]).norm())
        e = float((params[k].detach().cpu() - ref[k].detach()).norm()) / upd
        c = float((ctl[k].detach() - ref[k].detach()).norm()) / upd  # includes the 2^-16 perturbation of the initial weight itself
        assert upd > 0 and e <= 5e-3 + 4.0 * c, (k, e, c)
