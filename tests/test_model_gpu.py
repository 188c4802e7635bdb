"""Parity of the CUDA path (mainModel -> libdrn_sm100.so through the C ABI) with the CPU oracle and with the fixtures the
UNMODIFIED reference produced (tests/golden), on the same seeded inputs.

Tolerances (floating-point path, north_star: 1e-3 relative to fp32):
  * forward quantities (head outputs, losses, BatchNorm running statistics): max|err| / max|ref| <= 1e-3.
  * gradients: rel-L2 <= 5e-4 + 4 x (full size) or 1e-3 + 8 x (the small golden cases) the oracle's OWN sensitivity (max over
    3 draws) to a 2^-16 relative perturbation of its inputs.  Train-mode BatchNorm followed by ReLU makes the early-layer gradients of this model ill-conditioned: at B=32,
    T=256 a 1.5e-5 relative perturbation of weights and features moves d prop_fc.weight of the fp32 oracle by 1.1e-2 (ReLU mask
    flips), so a fixed 1e-3 bound is not a property any 16-bit-mantissa (or TF32: the reference's own GPU default)
    implementation can have.  Measured (profiles/r02_grad_errors.json, all 68-73 live gradient tensors, stages 1 / 2 / 3 at
    B=32, T=256): err / sens <= 1.24 / 0.65 / 1.73 -- the CUDA path errs like a 2^-16 input perturbation, which is what three
    BF16 products are -- so the factor 4 is ~2x the worst observed ratio.  In the small golden cases (32 - 512 rows per level) ONE
    flipped ReLU mask is visible in a whole tensor (scripts/bwd_debug.py: B=4, T=64, one element of the level-3 bbox tower with
    |bn(y)| = 2.4e-5 flips and moves that layer's gradient by 2e-2 while all others agree to 1e-5), so err / sens scatters up
    to 5 there (profiles/r02_grad_errors_golden_cases.json: err / (1e-3 + sens) <= 3.9) and the bound is 1e-3 + 8 x sens.
    Well-conditioned gradients (the head) are additionally held to a plain bound: 2e-4 at full size (observed <= 4.5e-5),
    1e-3 in the small cases (observed <= 5e-4).
"""
import os

import numpy as np
import pytest
import torch

from drn_b200 import spec as spec_mod
from drn_b200 import synthetic as S
from oracle import drn_oracle as O

pytestmark = pytest.mark.gpu
FWD_TOL = 1e-3


def _build(name=None, B=None, T=None, L=10, stage=1, training=True, crafted=False):
    if name is not None:
        B, T, L, stage, training, crafted = S.GOLDEN_CASES[name]
    cfg0 = S.default_config(stage=stage)
    spec = spec_mod.state_dict_spec(cfg0)
    if name is not None:
        cfg, sd, batch, stage, training = S.golden_case(name, spec)
    else:
        cfg = cfg0
        sd = S.synth_state_dict(spec)
        batch = S.synth_batch(B, T, max_len=L, embedding=sd["query_encoder.embedding.weight"])
    return cfg, sd, batch, stage, training


def _cuda_model(sd, stage, training):
    from model.main_model import mainModel
    model = mainModel(1301, S.config_namespace(stage=stage))
    model.load_state_dict(sd)
    if stage == 1:
        for k, p in model.named_parameters():
            if O.frozen_in_stage1(k):
                p.requires_grad = False
    return model.cuda().train(training)


def _run_cuda(model, batch):
    return model(batch["query_tokens"], batch["query_length"], batch["props_features"], batch["props_start_end"],
                 batch["gt_start_end"], None, None)


def _head_outputs(model, B):
    path = list(model._paths.values())[-1]
    out = {}
    for i in range(3):
        o, Tl = path.lvl_off[i], path.Tl[i]
        out["logits%d" % i] = path.cls_raw[o:o + B * Tl].view(B, 1, Tl).cpu()
        out["bbox%d" % i] = path.bbox[o:o + B * Tl].view(B, Tl, 2).permute(0, 2, 1).cpu()
        out["iou%d" % i] = path.iou_raw[o:o + B * Tl].view(B, 1, Tl).cpu()
    return out


def _maxrel(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


def _oracle(sd, cfg, batch, stage, training, perturb=0.0, seed=7):
    leaf = {}
    g = torch.Generator().manual_seed(seed)
    for k, v in sd.items():
        v = v.clone()
        if v.is_floating_point() and "running_" not in k:
            if perturb:
                v = v * (1 + perturb * (2 * torch.rand(v.shape, generator=g) - 1))
            v.requires_grad_(not (stage == 1 and O.frozen_in_stage1(k)))
        leaf[k] = v
    b = dict(batch)
    if perturb:
        f = batch["props_features"]
        b["props_features"] = f * (1 + perturb * (2 * torch.rand(f.shape, generator=g) - 1))
    cap = {}
    boxes, ld, newbuf = O.forward(leaf, cfg, b, training=training, capture=cap)
    grads = None
    if training:
        loss = O.total_loss(ld, stage)
        if loss.requires_grad:
            loss.backward()
        grads = {k: v.grad for k, v in leaf.items() if v.requires_grad and v.grad is not None}
    return boxes, ld, newbuf, cap, grads


@pytest.mark.parametrize("name", list(S.GOLDEN_CASES))
def test_forward_matches_oracle_and_reference_golden(name, golden_dir):
    torch.set_num_threads(os.cpu_count())
    cfg, sd, batch, stage, training = _build(name)
    B = batch["props_features"].shape[0]
    model = _cuda_model(sd, stage, training)
    with torch.no_grad():
        boxes, ld = _run_cuda(model, batch)
    oboxes, old, newbuf, cap, _ = _oracle(sd, cfg, batch, stage, training)
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    mine = _head_outputs(model, B)
    for k, v in mine.items():
        assert _maxrel(v, cap[k]) <= FWD_TOL, (k, _maxrel(v, cap[k]))
        assert _maxrel(v, g["head/" + k]) <= FWD_TOL, ("golden", k)
    for k in ("loss_cls", "loss_reg", "loss_iou"):
        a = ld[k].detach().cpu().reshape(-1).double()
        assert a.shape == old[k].reshape(-1).shape
        assert _maxrel(a, old[k].detach().reshape(-1)) <= FWD_TOL or float(a.abs().max()) == float(old[k].abs().max()) == 0.0, k
        assert _maxrel(a, g["loss/" + k]) <= FWD_TOL or float(a.abs().max()) == 0.0, ("golden", k)
        assert str(ld[k].dtype) == str(g["loss_dtype/" + k]), k
    msd = model.state_dict()
    for k, v in newbuf.items():
        if v.is_floating_point():
            assert _maxrel(msd[k].cpu(), v) <= 1e-4, k
        else:
            assert int(msd[k]) == int(v), k
    if not training:
        assert len(boxes) == len(oboxes)
        for d, od in zip(boxes, oboxes):
            assert d["detections"].shape == od["detections"].shape
            o1 = np.lexsort((d["detections"][:, 0].cpu().numpy(), d["scores"].cpu().numpy()))
            o2 = np.lexsort((od["detections"][:, 0].numpy(), od["scores"].numpy()))
            assert _maxrel(d["detections"].cpu()[o1], od["detections"][o2]) <= FWD_TOL
            assert _maxrel(d["scores"].cpu()[o1], od["scores"][o2]) <= FWD_TOL
            assert sorted(x for l in d["level"] for x in l) == sorted(x for l in od["level"] for x in l)


WELL_CONDITIONED = ("fcos.head.cls_logits", "fcos.head.bbox_pred", "fcos.head.scales", "fcos.head.iou_scores")


@pytest.mark.parametrize("name", [n for n, c in S.GOLDEN_CASES.items() if c[4]] + ["full_b32_t256"])
def test_gradients_match_oracle(name):
    """`full_b32_t256` = BASELINE config 2 size: the only case where the weight gradients run with K-split slices and all
    grouped launches carry more tiles than SM pairs."""
    torch.set_num_threads(os.cpu_count())
    cfg, sd, batch, stage, training = _build(B=32, T=256) if name == "full_b32_t256" else _build(name)
    model = _cuda_model(sd, stage, True)
    _, ld = _run_cuda(model, batch)
    loss = ld["loss_iou"] if stage == 2 else sum(ld.values())
    assert loss.requires_grad
    loss.backward()
    torch.cuda.synchronize()
    _, _, _, _, ref = _oracle(sd, cfg, batch, stage, True)
    perts = [_oracle(sd, cfg, batch, stage, True, perturb=2.0 ** -16, seed=sd_)[4] for sd_ in (7, 8, 9)]
    params = dict(model.named_parameters())
    checked = 0
    for k, gref in ref.items():
        p = params[k]
        n = float(gref.norm())
        if n < 1e-6:  # mathematically zero (conv bias in front of train-mode BN, softmax shift): round-off on either side
            assert p.grad is None or float(p.grad.norm()) < 1e-5, k
            continue
        assert p.grad is not None, k
        err = float((p.grad.cpu().double() - gref.double()).norm()) / n
        sens = max(float((pt[k].double() - gref.double()).norm()) / n for pt in perts)
        full = name == "full_b32_t256"
        tol = 5e-4 + 4.0 * sens if full else 1e-3 + 8.0 * sens
        assert err <= tol, "%s: rel-L2 %.2e > %.2e (oracle sensitivity %.2e)" % (k, err, tol, sens)
        if k.startswith(WELL_CONDITIONED):
            assert err <= (2e-4 if full else 1e-3), (k, err)
        checked += 1
    assert checked > 40
    # parameters the reference leaves without a gradient stay without one (textualAttention, centerness; frozen in stage 1)
    for k, p in params.items():
        if k not in ref:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k


def test_full_size_forward_and_properties():
    """BASELINE config 2 size (B=32, T=256): forward parity with the oracle, plus size-independent properties of the backward:
    linearity in the upstream loss gradient, zero-sum of every train-mode BatchNorm input gradient."""
    torch.set_num_threads(os.cpu_count())
    cfg, sd, batch, stage, training = _build(B=32, T=256)
    model = _cuda_model(sd, 1, True)
    _, ld = _run_cuda(model, batch)
    (ld["loss_cls"] + ld["loss_reg"]).backward()
    g1 = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    _, old, _, cap, _ = _oracle(sd, cfg, batch, 1, True)
    for k in ("loss_cls", "loss_reg"):
        assert _maxrel(ld[k].detach().cpu(), old[k].detach()) <= FWD_TOL, k
    mine = _head_outputs(model, 32)
    for k, v in mine.items():
        assert _maxrel(v, cap[k]) <= FWD_TOL, (k, _maxrel(v, cap[k]))
    path = list(model._paths.values())[-1]
    for blk in path.conv + path.layer + path.inner + path.tower:
        dy = blk.dy.to_float().double()
        colsum = dy.sum(dim=(0, 1)).abs().max().item()
        assert colsum <= 1e-4 * max(dy.abs().sum(dim=(0, 1)).max().item(), 1e-30), blk.prefix
    for p in model.parameters():
        p.grad = None
    _, ld2 = _run_cuda(model, batch)  # BatchNorm batch statistics do not depend on the running buffers: same graph
    (2.0 * ld2["loss_cls"] + 2.0 * ld2["loss_reg"]).backward()
    for k, p in model.named_parameters():
        if p.grad is not None and k in g1 and float(g1[k].norm()) > 1e-6:
            assert float((p.grad - 2.0 * g1[k]).norm() / (2.0 * g1[k]).norm()) <= 1e-4, k


@pytest.mark.parametrize("first", [True, False])
def test_postprocess_kernel_matches_oracle(first):
    """drn_postprocess against oracle.postprocess (reference inference.py:49-215) on random head outputs: more candidates than
    top_n on some (sample, level) pairs, none on others, and an all-empty sample (fallback detection)."""
    import ctypes as C
    from drn_b200 import lib as L
    from drn_b200.dense import DensePath
    from model.inference import assemble
    torch.manual_seed(1)
    cfg = S.default_config(stage=1 if first else 3)
    B, T = 4, 256
    Tl = (T, T // 2, T // 4)
    logits = [torch.randn(B, 1, t) * 2 - 1 for t in Tl]
    bbox = [torch.rand(B, 2, t) * 8 for t in Tl]
    iou = [torch.randn(B, 1, t) for t in Tl]
    logits[0][2] = -20.0  # sample 2: nothing passes on level 0
    for l in range(3):
        logits[l][3] = -20.0  # sample 3: nothing passes anywhere
    ref = O.postprocess(O.compute_locations(T, cfg["fpn_stride"]), logits, bbox, iou, cfg)
    path = DensePath(cfg, B, T, torch.device("cuda"), L=4)
    path.cls_raw.copy_(torch.cat([x.permute(0, 2, 1).reshape(-1) for x in logits]))
    path.bbox.copy_(torch.cat([x.permute(0, 2, 1).reshape(-1, 2) for x in bbox]))
    path.iou_raw.copy_(torch.cat([x.permute(0, 2, 1).reshape(-1) for x in iou]))
    got = assemble(*path.postprocess())
    assert all(d["detections"].is_cuda for d in got)  # inference.py:193-196 returns CUDA tensors
    got = [{k: (v.cpu() if torch.is_tensor(v) else v) for k, v in d.items()} for d in got]
    assert max(len(x) for r in ref for x in r["level"]) == 32  # the top-k branch is exercised
    for g, r in zip(got, ref):
        assert g["detections"].shape == r["detections"].shape
        og = torch.argsort(g["scores"] + g["detections"][:, 0] * 1e-3)
        orf = torch.argsort(r["scores"] + r["detections"][:, 0] * 1e-3)
        assert torch.allclose(g["detections"][og], r["detections"][orf], atol=1e-6)
        assert torch.allclose(g["scores"][og], r["scores"][orf], atol=1e-6)
        assert torch.allclose(g["locations"][og], r["locations"][orf], atol=1e-6)
        assert g["level"] == r["level"]
    assert got[3]["detections"].tolist() == [[0.0, 1.0]] and got[3]["level"] == [[-1]]


@pytest.mark.parametrize("B,T", [(64, 64), (16, 512)])
def test_eval_forward_sweep_sizes(B, T):
    """BASELINE configs[4] (inference sweep T in {64..512}): eval-mode forward (BatchNorm from the running statistics) against
    the oracle at the sweep's extreme T values; detections compared after the reference's own ordering-free canonicalisation."""
    torch.set_num_threads(os.cpu_count())
    cfg, sd, batch, stage, _ = _build(B=B, T=T)
    model = _cuda_model(sd, 1, False)
    with torch.no_grad():
        boxes, ld = _run_cuda(model, batch)
    oboxes, old, _, cap, _ = _oracle(sd, cfg, batch, 1, False)
    mine = _head_outputs(model, B)
    for k, v in mine.items():
        assert _maxrel(v, cap[k]) <= FWD_TOL, (k, _maxrel(v, cap[k]))
    for k in ("loss_cls", "loss_reg"):
        assert _maxrel(ld[k].detach().cpu(), old[k].detach()) <= FWD_TOL, k
    assert len(boxes) == len(oboxes) == B
    same_shape = sum(1 for d, od in zip(boxes, oboxes) if d["detections"].shape == od["detections"].shape)
    assert same_shape >= B - 1  # a class score within 1e-7 of the 0.05 threshold may flip a single candidate
    # BatchNorm buffers are untouched in eval mode
    msd = model.state_dict()
    for k, v in sd.items():
        if "running_" in k or "num_batches" in k:
            assert torch.equal(msd[k].cpu(), v), k


@pytest.mark.parametrize("L", [7, 14])
def test_token_width_buckets(L):
    """The reference collate pads queries to the longest of the batch (dataset.py:186,198), so the token width changes from
    batch to batch; mainModel pads it up to a bucket (zero tokens beyond the lengths are masked).  Losses and head gradients
    must be those of the oracle run on the UNPADDED tokens, and widths of one bucket must share one DensePath."""
    torch.set_num_threads(os.cpu_count())
    cfg, sd, batch, stage, training = _build(B=4, T=64, L=L)
    assert batch["query_tokens"].shape[1] == L
    model = _cuda_model(sd, 1, True)
    _, ld = _run_cuda(model, batch)
    sum(ld.values()).backward()
    _, old, _, cap, grads = _oracle(sd, cfg, batch, 1, True)
    perts = [_oracle(sd, cfg, batch, 1, True, perturb=2.0 ** -16, seed=s_)[4] for s_ in (7, 8, 9)]
    for k in ("loss_cls", "loss_reg"):
        assert _maxrel(ld[k].detach().cpu(), old[k].detach()) <= FWD_TOL, k
    params = dict(model.named_parameters())
    for k in ("fcos.head.cls_logits.weight", "fcos.head.bbox_pred.weight", "query_encoder.embedding.weight"):
        g, r = params[k].grad.cpu(), grads[k]
        sens = max(float((pt[k] - r).norm() / r.norm()) for pt in perts)  # the embedding sits at the far end of the chain
        assert float((g - r).norm() / r.norm()) <= 1e-3 + 8.0 * sens, (k, float((g - r).norm() / r.norm()), sens)
    ge = params["query_encoder.embedding.weight"].grad.cpu()
    used = torch.zeros(ge.shape[0], dtype=torch.bool)
    for b in range(batch["query_tokens"].shape[0]):
        used[batch["query_tokens"][b, :int(batch["query_length"][b])]] = True
    assert float(ge[~used].abs().max()) == 0.0 and float(ge[0].abs().max()) == 0.0  # padding tokens contribute nothing
    # a narrower batch of the same bucket reuses the buffers and graphs
    n_paths = len(model._paths)
    b2 = dict(batch)
    b2["query_tokens"] = batch["query_tokens"][:, :L - 1].contiguous()
    b2["query_length"] = batch["query_length"].clamp(max=L - 1)
    _run_cuda(model, b2)
    assert len(model._paths) == n_paths


def _summary(t):
    """oracle/make_goldens.py:summarize -- [norm, sum, 16 sampled values] of a tensor in the reference's layout."""
    t = t.detach().to(torch.float64).reshape(-1).cpu()
    idx = torch.from_numpy(S.sample_indices(t.numel()))
    return torch.cat([torch.tensor([float(t.norm()), float(t.sum())], dtype=torch.float64), t[idx]]), float(t.abs().sum())


@pytest.mark.parametrize("name", list(S.GOLDEN_CASES))
def test_intermediates_match_reference_golden(name, golden_dir):
    """The goldens hold summaries of thirteen intermediates captured by forward hooks on the UNMODIFIED reference (gates, prop_fc
    output, backbone C1-C3, FPN laterals / outputs): the CUDA path's buffers must reproduce them -- a wrong layer cannot hide
    behind a matching loss."""
    cfg, sd, batch, stage, training = _build(name)
    model = _cuda_model(sd, stage, training)
    with torch.no_grad():
        _run_cuda(model, batch)
    path = list(model._paths.values())[-1]
    bct = lambda pl: pl.to_float().permute(0, 2, 1)  # noqa: E731  planes [B,T,C] -> reference layout [B,C,T]
    I = [bct(path.I[i]) for i in range(3)]
    mine = {"q0": path.q[0], "q1": path.q[1], "q2": path.q[2], "P_btd": path.Pre,
            "C1": bct(path.Cact[0]), "C2": bct(path.Cact[1]), "C3": bct(path.Cact[2]),
            "I3": I[2], "P3": bct(path.Pf[2]), "P2": bct(path.Pf[1]), "P1": bct(path.Pf[0]),
            # fused upsample-add (FPN.py:63-68): the lateral before the add is I_l - up2(I_{l+1})
            "L2": I[1] - I[2].repeat_interleave(2, dim=2), "L1": I[0] - I[1].repeat_interleave(2, dim=2)}
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    for k, v in mine.items():
        ref = torch.from_numpy(g["cap/" + k])
        got, l1 = _summary(v)
        scale = float(ref[2:].abs().max())
        assert abs(float(got[0] - ref[0])) <= 1e-3 * float(ref[0]), (k, "norm", float(got[0]), float(ref[0]))
        assert abs(float(got[1] - ref[1])) <= 1e-3 * l1, (k, "sum")
        assert float((got[2:] - ref[2:]).abs().max()) <= 1e-3 * max(scale, float(ref[0]) / v.numel() ** 0.5), (k, "samples")


def test_eval_parity_at_sweep_batch_256():
    """BASELINE configs[4] is quoted at batch 256: eval-mode forward + candidate selection at B=256, T=64 against the oracle
    (head outputs 1e-3; detections per sample after the ordering-free canonicalisation; at most a threshold flip or two)."""
    torch.set_num_threads(os.cpu_count())
    B, T = 256, 64
    cfg, sd, batch, stage, _ = _build(B=B, T=T)
    model = _cuda_model(sd, 1, False)
    with torch.no_grad():
        boxes, ld = _run_cuda(model, batch)
    oboxes, old, _, cap, _ = _oracle(sd, cfg, batch, 1, False)
    mine = _head_outputs(model, B)
    for k, v in mine.items():
        assert _maxrel(v, cap[k]) <= FWD_TOL, (k, _maxrel(v, cap[k]))
    for k in ("loss_cls", "loss_reg"):
        assert _maxrel(ld[k].detach().cpu(), old[k].detach()) <= FWD_TOL, k
    assert len(boxes) == len(oboxes) == B
    same = 0
    for d, od in zip(boxes, oboxes):
        if d["detections"].shape != od["detections"].shape:
            continue
        o1 = np.lexsort((d["detections"][:, 0].cpu().numpy(), d["scores"].cpu().numpy()))
        o2 = np.lexsort((od["detections"][:, 0].numpy(), od["scores"].numpy()))
        if _maxrel(d["detections"].cpu()[o1], od["detections"][o2]) <= FWD_TOL and _maxrel(d["scores"].cpu()[o1], od["scores"][o2]) <= FWD_TOL:
            same += 1
    assert same >= B - 2, same


def test_invalid_tokens_and_lengths_are_rejected_not_dereferenced():
    """token ids index the embedding table (and its gradient) and lengths index the LSTM output.  Host tensors raise
    like the reference (nn.Embedding IndexError / pack_padded_sequence); device tensors are sanitised by drn_qe_stage: the batch's
    losses are NaN, input_error() names the problem, nothing outside the buffers is touched and the next valid batch is exact."""
    cfg, sd, batch, stage, _ = _build(B=4, T=32, L=8)
    model = _cuda_model(sd, 1, True)
    _, ld = _run_cuda(model, batch)
    sum(ld.values()).backward()
    good = {k: float(v) for k, v in ld.items()}
    gemb = model.query_encoder.embedding.weight.grad.clone()
    assert model.input_error() is None
    bad_tok = batch["query_tokens"].clone()
    bad_tok[1, 0] = 1302          # one past the table
    bad_tok[2, 1] = -5
    with pytest.raises(IndexError):
        model(bad_tok, batch["query_length"], batch["props_features"], batch["props_start_end"], batch["gt_start_end"], None, None)
    bad_len = batch["query_length"].clone()
    bad_len[3] = 0
    with pytest.raises(RuntimeError):
        model(batch["query_tokens"], bad_len, batch["props_features"], batch["props_start_end"], batch["gt_start_end"], None, None)
    for tok, ln, what in ((bad_tok.cuda(), batch["query_length"].cuda(), "token"), (batch["query_tokens"].cuda(), bad_len.cuda(), "length"),
                          (batch["query_tokens"].cuda(), torch.full_like(bad_len, 99).cuda(), "length")):
        for p in model.parameters():
            p.grad = None
        _, ld2 = model(tok, ln, batch["props_features"], batch["props_start_end"], batch["gt_start_end"], None, None)
        assert all(torch.isnan(v).all() for k, v in ld2.items() if k != "loss_iou")
        (ld2["loss_cls"] + ld2["loss_reg"]).backward()   # must not fault: sanitised ids / lengths
        torch.cuda.synchronize()
        msg = model.input_error()
        assert msg is not None and what in msg, msg
        assert model.input_error() is None  # flags are cleared by the read
    for p in model.parameters():
        p.grad = None
    _, ld3 = model(batch["query_tokens"].cuda(), batch["query_length"].cuda(), batch["props_features"], batch["props_start_end"],
                   batch["gt_start_end"], None, None)
    sum(ld3.values()).backward()
    assert {k: float(v) for k, v in ld3.items()} == good
    assert torch.equal(model.query_encoder.embedding.weight.grad, gemb) or \
        float((model.query_encoder.embedding.weight.grad - gemb).abs().max()) <= 1e-6 * float(gemb.abs().max())
