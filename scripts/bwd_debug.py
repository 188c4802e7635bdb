"""Backward bring-up probe: gradients w.r.t. the ACTIVATIONS of the CUDA path (its dX buffers) against the oracle's autograd
(retain_grad on the captured intermediates), level by level.  Localises a gradient error to a layer.  Never asserts.
    PB=4 PT=64 python scripts/bwd_debug.py"""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from drn_b200 import spec as spec_mod  # noqa: E402
from drn_b200 import synthetic as S  # noqa: E402
from model.main_model import mainModel  # noqa: E402
from oracle import drn_oracle as O  # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / max(b.norm().item(), 1e-30)).item(), (a - b).abs().max().item(), b.abs().max().item()


def main():
    B, T = int(os.environ.get("PB", 4)), int(os.environ.get("PT", 64))
    seed = S.SEED + int(os.environ.get("PSEED", 50))
    torch.set_num_threads(os.cpu_count())
    cfg = S.default_config(stage=1)
    sd = S.synth_state_dict(spec_mod.state_dict_spec(cfg))
    batch = S.synth_batch(B, T, max_len=8, embedding=sd["query_encoder.embedding.weight"], seed=seed)
    model = mainModel(1301, S.config_namespace(stage=1))
    model.load_state_dict(sd)
    for k, p in model.named_parameters():
        if O.frozen_in_stage1(k):
            p.requires_grad = False
    model = model.cuda().train()
    os.environ["DRN_NO_GRAPHS"] = "1"
    _, ld = model(batch["query_tokens"], batch["query_length"], batch["props_features"], batch["props_start_end"], batch["gt_start_end"], None, None)
    sum(ld.values()).backward()
    torch.cuda.synchronize()
    path = list(model._paths.values())[0]
    leaf = {}
    for k, v in sd.items():
        v = v.clone()
        if v.is_floating_point() and "running_" not in k:
            v.requires_grad_(not O.frozen_in_stage1(k))
        leaf[k] = v
    cap = {}
    _, old, _ = O.forward(leaf, cfg, batch, training=True, capture=cap)
    for v in cap.values():
        if torch.is_tensor(v) and v.requires_grad:
            v.retain_grad()
    O.total_loss(old, 1).backward()
    g = lambda n: cap[n].grad  # noqa: E731  [B, C, T]
    btc = lambda t: t.permute(0, 2, 1)  # noqa: E731
    F = path.F
    print("B=%d T=%d  losses cuda %s oracle %s" % (B, T, [float(ld[k]) for k in ("loss_cls", "loss_reg")], [float(old[k]) for k in ("loss_cls", "loss_reg")]))
    n = B * path.P
    for l in range(3):
        o, Tl = path.lvl_off[l], path.Tl[l]
        print("level %d (T_l=%d)" % (l, Tl))
        rows = [("dcls (grad of logits)", path.dcls[o:o + B * Tl].view(B, Tl, 1), btc(g("logits%d" % l))),
                ("d tower cls half = dCt", path.dTW[l][:, :, :F], btc(g("Ct%d" % l))),
                ("d tower bbox half = dBt", path.dTW[l][:, :, F:], btc(g("Bt%d" % l))),
                ("dy tower cls (pre-BN)", path.tower[l].dy.to_float()[:, :, :F], btc(g("Ct%d.y" % l))),
                ("dy tower bbox (pre-BN)", path.tower[l].dy.to_float()[:, :, F:], btc(g("Bt%d.y" % l))),
                ("dP%d (grad of FPN output)" % (l + 1), path.dPf[l], btc(g("P%d" % (l + 1)))),
                ("dy layer%d" % (l + 1), path.layer[l].dy.to_float(), btc(g("P%d.y" % (l + 1)))),
                ("dy inner%d" % (l + 1), path.inner[l].dy.to_float(), btc(g(("I3" if l == 2 else "L%d" % (l + 1)) + ".y"))),
                ("dC%d" % (l + 1), path.dC[l], btc(g("C%d" % (l + 1)))),
                ("dy conv%d" % l, path.conv[l].dy.to_float(), btc(g("C%d.y" % (l + 1))))]
        # bbox raw gradient: oracle has grad of bbox = exp(raw * s); chain to raw: d raw = d bbox * bbox * s
        for name, mine, ref in rows:
            if ref is None:
                print("   %-28s oracle grad missing" % name)
                continue
            r, mx, sc = rel(mine, ref)
            print("   %-28s relL2 %.2e  max|err| %.2e  max|ref| %.2e" % (name, r, mx, sc))
        # element-level view of the worst tower-bbox dy errors: ReLU mask flip (isolated element, |bn(y)| ~ 0) or column-wide error?
        mine = path.tower[l].dy.to_float()[:, :, F:].detach().cpu()
        ref = btc(g("Bt%d.y" % l)).detach()
        err = (mine - ref).abs()
        ycu = path.tower[l].y[:, :, F:].detach().cpu()
        yor = btc(cap["Bt%d.y" % l]).detach()
        coef = path.tower[l].coef.detach().cpu()
        top = torch.topk(err.reshape(-1), 6).indices.tolist()
        colerr = err.sum(dim=(0, 1))
        print("   worst columns (sum |err|):", [(int(i), float(colerr[i])) for i in torch.topk(colerr, 4).indices])
        for idx in top:
            b_, t_, c_ = idx // (Tl * F), (idx // F) % Tl, idx % F
            cc = F + c_
            bn_cu = float(ycu[b_, t_, c_] * coef[0, cc] + coef[1, cc])
            act_or = float(btc(cap["Bt%d" % l])[b_, t_, c_])
            print("   (b=%d t=%d c=%d) dy cuda %.3e oracle %.3e | y cuda %.6f oracle %.6f | bn(y) cuda %.3e, oracle act %.3e | dBt cuda %.3e oracle %.3e"
                  % (b_, t_, c_, float(mine[b_, t_, c_]), float(ref[b_, t_, c_]), float(ycu[b_, t_, c_]), float(yor[b_, t_, c_]), bn_cu, act_or,
                     float(path.dTW[l][b_, t_, cc]), float(btc(g("Bt%d" % l))[b_, t_, c_])))
        db = path.dbox[o:o + B * Tl].view(B, Tl, 2)
        bb = cap["bbox%d" % l]
        ref = btc(bb.grad * bb.detach() * leaf["fcos.head.scales.%d.scale" % l].detach()) if bb.grad is not None else None
        if ref is not None:
            print("   %-28s relL2 %.2e  max|err| %.2e  max|ref| %.2e" % (("dbox (grad of bbox_pred out)",) + rel(db, ref)))


if __name__ == "__main__":
    main()
