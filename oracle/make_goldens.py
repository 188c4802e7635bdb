"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/<case>.npz by running the UNMODIFIED reference
(/root/reference, imported through oracle/shims) on the seeded cases of drn_b200.synthetic.GOLDEN_CASES.

    cd /tmp && python /root/repo/oracle/make_goldens.py          # needs /root/reference; CPU only

Weights and inputs are NOT stored: tests rebuild them from the seed (drn_b200.synthetic.golden_case).
Stored per case: the three losses, full head outputs, detections (eval), and for every gradient /
updated BatchNorm buffer / hooked intermediate its L2 norm, sum and 16 sampled elements.
"""
import os
import sys
import tempfile

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from drn_b200 import synthetic as S  # noqa: E402
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")

# reference module name -> oracle capture name ([B,C,T] outputs of the conv+BN+ReLU blocks)
HOOKS = {
    "backbone_net.forward_conv0": "C1", "backbone_net.forward_conv1": "C2", "backbone_net.forward_conv2": "C3",
    "fpn.fpn_inner3": "I3", "fpn.fpn_layer3": "P3", "fpn.fpn_inner2": "L2", "fpn.fpn_layer2": "P2",
    "fpn.fpn_inner1": "L1", "fpn.fpn_layer1": "P1", "prop_fc": "P_btd", "qInput0": "q0", "qInput1": "q1",
    "qInput2": "q2",
}


def summarize(t):
    t = t.detach().to(torch.float64).reshape(-1)
    idx = torch.from_numpy(S.sample_indices(t.numel()))
    return np.concatenate([[float(t.norm()), float(t.sum())], t[idx].numpy()])


def run_case(name):
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    B, T, L, stage, training, crafted = S.GOLDEN_CASES[name]
    cfg = S.default_config(stage=stage)
    model = ref_loader.build_reference_model(cfg)
    spec = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
    cfg, sd, batch, stage, training = S.golden_case(name, spec)
    model.load_state_dict(sd)
    model.train(training)

    out = {}
    caps = {}
    handles = []
    mods = dict(model.named_modules())
    for mname, cname in HOOKS.items():
        handles.append(mods[mname].register_forward_hook(
            lambda m, i, o, cname=cname: caps.__setitem__(cname, o)))
    head_out = {}
    orig_head_forward = model.fcos.head.forward

    def head_forward(x):
        r = orig_head_forward(x)
        head_out["r"] = r
        return r
    model.fcos.head.forward = head_forward

    boxes, loss_dict = model(batch["query_tokens"], batch["query_length"], batch["props_features"],
                             batch["props_start_end"], batch["gt_start_end"], None, None)
    for h in handles:
        h.remove()
    for k, v in loss_dict.items():
        out["loss/" + k] = v.detach().to(torch.float64).numpy().reshape(-1)
        out["loss_dtype/" + k] = np.array(str(v.dtype))
    logits, bbox, _, iou = head_out["r"]
    for l in range(3):
        out["head/logits%d" % l] = logits[l].detach().numpy()
        out["head/bbox%d" % l] = bbox[l].detach().numpy()
        out["head/iou%d" % l] = iou[l].detach().numpy()
    for cname, v in caps.items():
        out["cap/" + cname] = summarize(v)
    if training:
        if stage == 2:
            loss = loss_dict["loss_iou"]  # main.py:222-225
        else:
            loss = sum(loss_dict.values())
        if loss.requires_grad:
            loss.backward()
        for k, p in model.named_parameters():
            if p.grad is not None:
                out["grad/" + k] = summarize(p.grad)
        for k, v in model.state_dict().items():
            if "running_" in k or "num_batches" in k:
                out["buf/" + k] = summarize(v)
    else:
        for b, d in enumerate(boxes):
            out["det/%d/detections" % b] = d["detections"].detach().numpy()
            out["det/%d/scores" % b] = d["scores"].detach().numpy()
            out["det/%d/locations" % b] = d["locations"].detach().numpy()
            out["det/%d/level" % b] = np.array([x for lv in d["level"] for x in lv], dtype=np.int64)
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    os.chdir(tempfile.mkdtemp())  # eval forward writes ./total_points.pkl (fcos.py:182)
    names = sys.argv[1:] or list(S.GOLDEN_CASES)
    for name in names:
        out = run_case(name)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, {k: v.tolist() for k, v in out.items() if k.startswith("loss/")}, os.path.getsize(path))


if __name__ == "__main__":
    main()
