"""Bring-up probe: runs golden cases through the CUDA mainModel and the CPU oracle, prints the error of every captured
intermediate, loss and gradient.  Never asserts."""
import os
import sys
import traceback

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from drn_b200 import spec as spec_mod  # noqa: E402
from drn_b200 import synthetic as S  # noqa: E402
from model.main_model import mainModel  # noqa: E402
from oracle import drn_oracle as O  # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    den = b.abs().max().item()
    return (a - b).abs().max().item() / max(den, 1e-30), ((a - b).norm() / max(b.norm().item(), 1e-30)).item()


def run(name):
    B, T, L, stage, training, crafted = S.GOLDEN_CASES[name]
    cfg0 = S.default_config(stage=stage)
    cfg, sd, batch, stage, training = S.golden_case(name, spec_mod.state_dict_spec(cfg0))
    print("=== %s B=%d T=%d stage=%d training=%s" % (name, B, T, stage, training), flush=True)
    model = mainModel(1301, S.config_namespace(stage=stage))
    model.load_state_dict(sd)
    if stage == 1:
        for k, p in model.named_parameters():
            if O.frozen_in_stage1(k):
                p.requires_grad = False
    model = model.cuda()
    model.train(training)
    boxes, ld = model(batch["query_tokens"], batch["query_length"], batch["props_features"], batch["props_start_end"],
                      batch["gt_start_end"], None, None)
    torch.cuda.synchronize()
    # oracle
    leaf = {}
    for k, v in sd.items():
        v = v.clone()
        if v.is_floating_point() and "running_" not in k:
            v.requires_grad_(not (stage == 1 and O.frozen_in_stage1(k)))
        leaf[k] = v
    cap = {}
    oboxes, old, newbuf = O.forward(leaf, cfg, batch, training=training, capture=cap)
    for k in ("loss_cls", "loss_reg", "loss_iou"):
        print("%-10s mine=%s oracle=%s" % (k, ld[k].detach().cpu().reshape(-1).tolist(), old[k].detach().reshape(-1).tolist()))
    path = list(model._paths.values())[0]
    chk = {"q0": path.q[0], "q1": path.q[1], "q2": path.q[2], "P": path.Pre.permute(0, 2, 1)}
    for i, n in enumerate(("C1", "C2", "C3")):
        chk[n] = path.Cact[i].to_float().permute(0, 2, 1)
        chk[n + ".y"] = path.conv[i].y.permute(0, 2, 1)
    chk["I3"] = path.I[2].to_float().permute(0, 2, 1)
    for i in range(3):
        chk["P%d" % (i + 1)] = path.Pf[i].to_float().permute(0, 2, 1)
        chk["Mx%d" % i] = path.MX[i].to_float().permute(0, 2, 1)
        chk["Hi%d" % i] = path.HI[i].to_float().permute(0, 2, 1)
        chk["Ct%d" % i] = path.TW[i].to_float()[:, :, :512].permute(0, 2, 1)
        chk["Bt%d" % i] = path.TW[i].to_float()[:, :, 512:].permute(0, 2, 1)
        o, Tl = path.lvl_off[i], path.Tl[i]
        chk["logits%d" % i] = path.cls_raw[o:o + B * Tl].view(B, 1, Tl)
        chk["bbox%d" % i] = path.bbox[o:o + B * Tl].view(B, Tl, 2).permute(0, 2, 1)
        chk["iou%d" % i] = path.iou_raw[o:o + B * Tl].view(B, 1, Tl)
    for k, v in chk.items():
        if k in cap:
            print("  fwd %-10s max/max %.2e  relL2 %.2e" % ((k,) + rel(v, cap[k])))
    msd = model.state_dict()
    worst = 0
    for k, v in newbuf.items():
        e = rel(msd[k].double(), v.double())[0]
        worst = max(worst, e)
        if e > 1e-4:
            print("  buf %-50s %.2e" % (k, e))
    print("  buffers worst %.2e" % worst)
    if training:
        loss = O.total_loss(old, stage)
        if loss.requires_grad:
            loss.backward()
        mine = ld["loss_iou"] if stage == 2 else sum(ld.values())
        if mine.requires_grad:
            mine.backward()
        torch.cuda.synchronize()
        for k, p in model.named_parameters():
            og = leaf[k].grad
            if p.grad is None and og is None:
                continue
            if p.grad is None or og is None:
                print("  grad %-50s mine %s oracle %s" % (k, p.grad is not None, og is not None))
                continue
            if og.abs().max() < 1e-6:
                print("  grad %-50s ~zero: mine max %.2e oracle max %.2e" % (k, p.grad.abs().max().item(), og.abs().max().item()))
                continue
            print("  grad %-50s max/max %.2e  relL2 %.2e" % ((k,) + rel(p.grad, og)))
    else:
        for b, (d, od) in enumerate(zip(boxes, oboxes)):
            print("  det %d n=%d/%d" % (b, d["detections"].shape[0], od["detections"].shape[0]))


if __name__ == "__main__":
    for name in sys.argv[1:] or ["s1_train_b4_t32", "s3_train_b4_t32_crafted", "s2_train_b4_t32_crafted", "c1_eval_b1_t64",
                                 "s1_train_b2_t256"]:
        try:
            run(name)
        except Exception:  # noqa: BLE001
            traceback.print_exc()
