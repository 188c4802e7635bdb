"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the DRN dense-regression hot path.

A functional, plain-PyTorch fp32 restatement of the reference algorithm,
written from the math, operating on a `state_dict` with the reference's key names.  It exists so
that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg can check
and time the CUDA path on machines where /root/reference does not exist.  Nothing in the product
path (drn_b200/, model/) may import it.

Pinning: tests/test_oracle_golden.py checks this file against tests/golden/*.npz, which were
produced by running the UNMODIFIED reference (oracle/make_goldens.py) on the same seeded inputs.
The reference repo has no tests / golden vectors of its own, so those
reference-generated fixtures are the pin.

Each function cites the reference file:line it restates (paths relative to /root/reference).
"""
import math

import torch
import torch.nn.functional as F

INF = 100000000  # model/loss.py:19
DOWNSAMPLE = 32.0  # hard-coded in model/loss.py:98,178 and model/inference.py:45
SIZE_BANDS = ((-1.0, 6.0), (5.6, 11.0), (11.0, float(INF)))  # model/loss.py:47-51
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


# ----------------------------------------------------------------------------------------------
# query encoder (model/language_module.py:27-62, model/ops.py:74-85)
# ----------------------------------------------------------------------------------------------
def _lstm_direction(x, lengths, w_ih, w_hh, b_ih, b_hh, reverse):
    """One direction of a packed 1-layer LSTM written out per time step.
    x [B,L,E]; returns [B,L,H], zero where t >= length (pad_packed_sequence semantics,
    language_module.py:42-46).  Gate order i,f,g,o (torch.nn.LSTM)."""
    B, L, _ = x.shape
    H = w_hh.shape[1]
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    xg = x @ w_ih.t() + b_ih + b_hh  # [B,L,4H]
    out = [None] * L
    steps = range(L - 1, -1, -1) if reverse else range(L)
    for t in steps:
        gates = xg[:, t] + h @ w_hh.t()
        i, f, g, o = gates.chunk(4, dim=1)
        c_new = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h_new = torch.sigmoid(o) * torch.tanh(c_new)
        live = (lengths > t).to(x.dtype).unsqueeze(1)
        c = live * c_new + (1 - live) * c
        h = live * h_new + (1 - live) * h
        out[t] = live * h_new
    return torch.stack(out, dim=1)


def query_encoder(sd, tokens, lengths):
    """language_module.py:38-62 -> list of 3 command vectors [B,1024]."""
    p = "query_encoder."
    L = int(lengths.max())
    tokens = tokens[:, :L]
    emb = F.embedding(tokens, sd[p + "embedding.weight"], padding_idx=0)
    fw = _lstm_direction(emb, lengths, sd[p + "biLSTM.weight_ih_l0"], sd[p + "biLSTM.weight_hh_l0"],
                         sd[p + "biLSTM.bias_ih_l0"], sd[p + "biLSTM.bias_hh_l0"], False)
    bw = _lstm_direction(emb, lengths, sd[p + "biLSTM.weight_ih_l0_reverse"],
                         sd[p + "biLSTM.weight_hh_l0_reverse"], sd[p + "biLSTM.bias_ih_l0_reverse"],
                         sd[p + "biLSTM.bias_hh_l0_reverse"], True)
    H = torch.cat([fw, bw], dim=-1)  # [B,L,1024]
    B = H.shape[0]
    last = H[torch.arange(B), lengths - 1]  # language_module.py:52
    v = torch.cat([H[:, 0], last], dim=-1)  # [B,2048]
    hid = F.relu(F.linear(v, sd[p + "qInput.weight"], sd[p + "qInput.bias"]))
    pos = torch.arange(L).unsqueeze(0)
    mask = pos >= lengths.unsqueeze(1)  # ops.py:74-85
    cmds = []
    for t in range(3):
        c = F.linear(hid, sd[p + "qInput%d.weight" % t], sd[p + "qInput%d.bias" % t])  # [B,1024]
        raw = F.linear(c[:, None, :] * H, sd[p + "cmd_inter2logits.weight"],
                       sd[p + "cmd_inter2logits.bias"]).squeeze(-1)
        raw = raw.masked_fill(mask, -1e30)
        att = F.softmax(raw, dim=-1)
        cmds.append(torch.bmm(att[:, None, :], H).squeeze(1))
    return cmds, H


# ----------------------------------------------------------------------------------------------
# conv + BatchNorm + ReLU block (model/basic_blocks.py:5-33; fcos.py:31-38)
# ----------------------------------------------------------------------------------------------
class _BNState:
    """Collects running-stat updates in call order (momentum 0.1, unbiased var; torch BatchNorm1d)."""

    def __init__(self, sd, training):
        self.sd = sd
        self.training = training
        self.new = {}

    def get(self, key):
        return self.new.get(key, self.sd[key])

    def bn(self, y, prefix):
        gamma, beta = self.sd[prefix + ".weight"], self.sd[prefix + ".bias"]
        if self.training:
            n = y.shape[0] * y.shape[2]
            mean = y.mean(dim=(0, 2))
            var = y.var(dim=(0, 2), unbiased=False)
            with torch.no_grad():
                rm, rv = self.get(prefix + ".running_mean"), self.get(prefix + ".running_var")
                self.new[prefix + ".running_mean"] = (1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mean
                self.new[prefix + ".running_var"] = (1 - BN_MOMENTUM) * rv + BN_MOMENTUM * var * (n / max(n - 1, 1))
                self.new[prefix + ".num_batches_tracked"] = self.get(prefix + ".num_batches_tracked") + 1
        else:
            mean, var = self.sd[prefix + ".running_mean"], self.sd[prefix + ".running_var"]
        xhat = (y - mean[None, :, None]) * torch.rsqrt(var[None, :, None] + BN_EPS)
        return xhat * gamma[None, :, None] + beta[None, :, None]


def _conv_bn_relu(st, x, prefix, stride=1, cap=None, name=None):
    w = st.sd[prefix + ".0.weight"]
    b = st.sd.get(prefix + ".0.bias")
    y = F.conv1d(x, w, b, stride=stride, padding=(w.shape[2] - 1) // 2)
    a = F.relu(st.bn(y, prefix + ".1"))
    if cap is not None and name is not None:
        cap[name + ".y"] = y
        cap[name] = a
    return a


# ----------------------------------------------------------------------------------------------
# whole forward (model/main_model.py:42-81)
# ----------------------------------------------------------------------------------------------
def compute_locations(T, strides):
    """fcos.py:193-211: loc_l[t] = s_l*t + s_l/2."""
    locs = []
    for lvl, s in enumerate(strides):
        t_l = T // (2 ** lvl)
        locs.append(torch.arange(0, t_l * s, step=s, dtype=torch.float32) + s / 2)
    return locs


def forward(sd, cfg, batch, training=True, capture=None):
    """Returns (box_lists or None, loss_dict, new_buffers).  `capture`: optional dict that receives
    named intermediates in the reference's [B,C,T] layout."""
    cap = capture
    st = _BNState(sd, training)
    tokens, lengths = batch["query_tokens"], batch["query_length"]
    feats, pse, gt = batch["props_features"], batch["props_start_end"], batch["gt_start_end"]

    cmds, _ = query_encoder(sd, tokens, lengths)
    q = [F.linear(cmds[i], sd["qInput%d.weight" % i], sd["qInput%d.bias" % i]) for i in range(3)]
    # position feature (main_model.py:53-55); only level 0 is consumed (backbone.py:31-32)
    dur = (pse[:, :, 1] - pse[:, :, 0]).unsqueeze(-1)
    pos_in = torch.cat((pse, dur), dim=-1).float()
    pos = F.linear(pos_in, sd["position_transform.weight"], sd["position_transform.bias"]).permute(0, 2, 1)
    P = F.linear(feats, sd["prop_fc.weight"], sd["prop_fc.bias"]).permute(0, 2, 1)  # [B,D,T]
    if cap is not None:
        cap.update(q0=q[0], q1=q[1], q2=q[2], cmd0=cmds[0], cmd1=cmds[1], cmd2=cmds[2], P=P, pos=pos)

    # backbone (backbone.py:27-34)
    x = torch.cat([q[0][:, :, None] * P, pos], dim=1)
    C1 = _conv_bn_relu(st, x, "backbone_net.forward_conv0", 1, cap, "C1")
    C2 = _conv_bn_relu(st, q[1][:, :, None] * C1, "backbone_net.forward_conv1", 2, cap, "C2")
    C3 = _conv_bn_relu(st, q[2][:, :, None] * C2, "backbone_net.forward_conv2", 2, cap, "C3")

    # FPN (FPN.py:54-69); BN update order inner3, layer3, inner2, layer2, inner1, layer1
    I3 = _conv_bn_relu(st, C3, "fpn.fpn_inner3", 1, cap, "I3")
    P3 = _conv_bn_relu(st, I3, "fpn.fpn_layer3", 1, cap, "P3")
    I2 = _conv_bn_relu(st, C2, "fpn.fpn_inner2", 1, cap, "L2") + I3.repeat_interleave(2, dim=2)
    P2 = _conv_bn_relu(st, I2, "fpn.fpn_layer2", 1, cap, "P2")
    I1 = _conv_bn_relu(st, C1, "fpn.fpn_inner1", 1, cap, "L1") + I2.repeat_interleave(2, dim=2)
    P1 = _conv_bn_relu(st, I1, "fpn.fpn_layer1", 1, cap, "P1")
    feats_l = [P1, P2, P3]

    # head (fcos.py:87-105), shared weights, per-level BN batch stats
    h = "fcos.head."
    logits, bbox, iou = [], [], []
    for l, f in enumerate(feats_l):
        ct = _conv_bn_relu(st, f, h + "cls_tower", 1, cap, "Ct%d" % l)
        bt = _conv_bn_relu(st, f, h + "bbox_tower", 1, cap, "Bt%d" % l)
        logits.append(F.conv1d(ct, sd[h + "cls_logits.weight"], sd[h + "cls_logits.bias"], padding=1))
        raw = F.conv1d(bt, sd[h + "bbox_pred.weight"], sd[h + "bbox_pred.bias"], padding=1)
        bbox.append(torch.exp(raw * sd[h + "scales.%d.scale" % l]))
        mix = _conv_bn_relu(st, torch.cat([ct, bt], dim=1), h + "mix_fc", 1, cap, "Mx%d" % l)
        hid = _conv_bn_relu(st, mix, h + "iou_scores", 1, cap, "Hi%d" % l)
        iou.append(F.conv1d(hid, sd[h + "iou_scores.3.weight"], sd[h + "iou_scores.3.bias"]))
    if cap is not None:
        for l in range(3):
            cap["logits%d" % l], cap["bbox%d" % l], cap["iou%d" % l] = logits[l], bbox[l], iou[l]

    T = feats.shape[1]
    locations = compute_locations(T, cfg["fpn_stride"])
    losses = fcos_losses(locations, logits, bbox, gt.float(), iou, cfg["is_first_stage"],
                         cfg["fcos_loss_gamma"], cfg["fcos_loss_alpha"])
    loss_dict = {"loss_cls": losses[0], "loss_reg": losses[1], "loss_iou": losses[2]}
    boxes = None
    if not training:
        boxes = postprocess(locations, logits, bbox, iou, cfg)
    return boxes, loss_dict, st.new


# ----------------------------------------------------------------------------------------------
# targets + losses (model/loss.py:40-239, layers/iou_loss.py:5-24, layers/sigmoid_focal_loss.py:40-52)
# ----------------------------------------------------------------------------------------------
def targets_for_locations(locations, gt):
    """loss.py:90-127 in closed form for the single GT segment per sample.
    locations: list of [T_l]; gt [B,2] fp32.  Returns per level (labels [B,T_l], reg [B,T_l,2])."""
    out = []
    S = (gt[:, 0] * DOWNSAMPLE)[:, None]
    E = (gt[:, 1] * DOWNSAMPLE)[:, None]
    for lvl, loc in enumerate(locations):
        l = loc[None, :] - S
        r = E - loc[None, :]
        inside = torch.minimum(l, r) > 0
        m = torch.maximum(l, r)
        lo, hi = SIZE_BANDS[lvl]
        cared = (m >= lo) & (m <= hi)
        labels = (inside & cared).to(torch.float32)
        out.append((labels, torch.stack([l, r], dim=-1)))
    return out


def sigmoid_focal_loss_sum(logits, labels, gamma, alpha):
    """Spec = sigmoid_focal_loss_cpu (sigmoid_focal_loss.py:40-52) with one foreground class,
    evaluated with the numerically stable log-sigmoid the reference's CUDA path uses."""
    p = torch.sigmoid(logits)
    pos = -alpha * (1 - p) ** gamma * F.logsigmoid(logits)
    neg = -(1 - alpha) * p ** gamma * F.logsigmoid(-logits)
    return (labels * pos + (1 - labels) * neg).sum()


def iou_loss_mean(pred, target):
    """layers/iou_loss.py:5-24."""
    inter = torch.minimum(pred[:, 1], target[:, 1]) + torch.minimum(pred[:, 0], target[:, 0])
    union = target[:, 0] + target[:, 1] + pred[:, 0] + pred[:, 1] - inter
    return (-torch.log((inter + 1e-8) / (union + 1e-8))).mean()


def segment_tiou(a, b):
    """loss.py:241-256."""
    inter = torch.clamp(torch.minimum(a[..., 1], b[..., 1]) - torch.maximum(a[..., 0], b[..., 0]), min=0)
    union = torch.clamp(torch.maximum(a[..., 1], b[..., 1]) - torch.minimum(a[..., 0], b[..., 0]), min=0)
    return inter / (union + 1e-6)


def fcos_losses(locations, logits, bbox, gt, iou, is_first_stage, gamma=2.0, alpha=0.25):
    """loss.py:134-239.  logits/bbox/iou: per-level lists of [B,1|2|1,T_l]."""
    if isinstance(gamma, (list, tuple)):
        gamma, alpha = gamma[0], alpha[0]
    B = logits[0].shape[0]
    tg = targets_for_locations(locations, gt)
    # flatten order: level-major, then image, then t (loss.py:159-163)
    lab = torch.cat([t[0].reshape(-1) for t in tg])
    reg_t = torch.cat([t[1].reshape(-1, 2) for t in tg])
    cls_f = torch.cat([x.permute(0, 2, 1).reshape(-1) for x in logits])
    reg_f = torch.cat([x.permute(0, 2, 1).reshape(-1, 2) for x in bbox])

    iou_loss = None
    if not is_first_stage:  # loss.py:168-198
        merged = torch.cat(bbox, dim=-1).transpose(2, 1)  # [B,P,2], level-major P
        loc = torch.cat(locations)[None, :]
        pred = torch.stack([loc - merged[:, :, 0], loc + merged[:, :, 1]], dim=-1) / DOWNSAMPLE
        first = pred[:, 0].clamp(min=0, max=1)  # quirk: clamps the first LOCATION (loss.py:180-181)
        pred = torch.cat([first[:, None, :], pred[:, 1:]], dim=1)
        tiou = segment_tiou(pred, gt[:, None, :])
        iou_pred = torch.cat(iou, dim=-1).squeeze().sigmoid()
        mask = tiou > 0.9
        if int(mask.sum()) == 0:
            iou_loss = torch.tensor([0])  # int64, no grad (loss.py:194-195)
        else:
            iou_loss = F.smooth_l1_loss(iou_pred[mask], tiou[mask])

    pos = lab > 0
    n_pos = int(pos.sum())
    cls_loss = sigmoid_focal_loss_sum(cls_f, lab, gamma, alpha) / (n_pos + B)
    if n_pos > 0:
        reg_loss = iou_loss_mean(reg_f[pos], reg_t[pos])
    else:
        reg_loss = reg_f[pos].sum()
    if is_first_stage:
        return cls_loss, reg_loss, torch.zeros(1)  # FloatTensor([0]) (loss.py:239)
    return cls_loss, reg_loss, iou_loss


# ----------------------------------------------------------------------------------------------
# eval post-processing (model/inference.py:49-215)
# ----------------------------------------------------------------------------------------------
def postprocess(locations, logits, bbox, iou, cfg):
    thr, top_n = cfg["fcos_inference_thr"], cfg["fcos_pre_nms_top_n"]
    first = cfg["is_first_stage"]
    B = logits[0].shape[0]
    results = []
    for b in range(B):
        dets, scores, levels, locs = [], [], [], []
        for lvl, loc in enumerate(locations):
            c = torch.sigmoid(logits[lvl][b, 0])
            cand = c > thr
            score = c if first else c * torch.sigmoid(iou[lvl][b, 0])
            idx = cand.nonzero().squeeze(1)
            s = score[idx]
            k = min(int(cand.sum()), top_n)
            if idx.numel() > k:
                s, top = s.topk(k, sorted=False)
                idx = idx[top]
            reg = bbox[lvl][b][:, idx]  # [2,n]
            d = torch.stack([loc[idx] - reg[0], loc[idx] + reg[1]], dim=1) / DOWNSAMPLE
            d = d.clamp(min=0, max=1)
            d = d[(d[:, 1] - d[:, 0]) >= 0]
            if d.shape[0]:
                dets.append(d)
            if s.numel():
                scores.append(torch.sqrt(s))
                locs.append(loc[idx] / 32)
            levels.append([lvl] * d.shape[0])
        if not dets:  # inference.py:192-197
            results.append({"detections": torch.tensor([[0.0, 1.0]]), "labels": [], "scores": torch.tensor([1.0]),
                            "level": [[-1]], "locations": torch.tensor([0.5])})
        else:
            results.append({"detections": torch.cat(dets), "labels": [], "scores": torch.cat(scores),
                            "level": levels, "locations": torch.cat(locs)})
    return results


# ----------------------------------------------------------------------------------------------
# helpers for tests / bench
# ----------------------------------------------------------------------------------------------
def total_loss(loss_dict, stage):
    """main.py:222-225."""
    if stage == 2:
        return loss_dict["loss_iou"]
    return sum(v for v in loss_dict.values())


def frozen_in_stage1(name):
    """main.py:126-128."""
    return "iou_scores" in name or "mix_fc" in name


def forward_backward(sd_values, cfg, batch, stage=1):
    """Convenience: leaf-ify floating tensors, run forward + backward, return (loss_dict, grads, new_buffers)."""
    sd = {}
    for k, v in sd_values.items():
        v = v.detach().clone()
        if v.is_floating_point() and "running_" not in k:
            v.requires_grad_(not (stage == 1 and frozen_in_stage1(k)))
        sd[k] = v
    _, loss_dict, new_buf = forward(sd, cfg, batch, training=True)
    loss = total_loss(loss_dict, stage)
    if loss.requires_grad:
        loss.backward()
    grads = {k: v.grad for k, v in sd.items() if v.requires_grad and v.grad is not None}
    return loss_dict, grads, new_buf
