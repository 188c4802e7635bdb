"""state_dict contract of the reference `mainModel` (names, shapes, order) -- SURVEY.md section 2a / 8b.

Checkpoints written by the reference's main.py:369-373 must load into the drop-in model and vice versa
(main.py:104-111 does a key-matched partial load), so the names below are part of the boundary.
"""


def _conv_bn(prefix, c_out, c_in, k, bias=False):
    out = [(prefix + ".0.weight", (c_out, c_in, k))]
    if bias:
        out.append((prefix + ".0.bias", (c_out,)))
    out += [
        (prefix + ".1.weight", (c_out,)),
        (prefix + ".1.bias", (c_out,)),
        (prefix + ".1.running_mean", (c_out,)),
        (prefix + ".1.running_var", (c_out,)),
        (prefix + ".1.num_batches_tracked", ()),
    ]
    return out


def state_dict_spec(cfg, vocab_size=1301, hidden_dim=512, embed_dim=300):
    """Ordered (name, shape) list equal to `mainModel(...).state_dict()` of the reference
    (model/main_model.py:14-40, language_module.py:10-25,65-75, backbone.py:5-15, FPN.py:26-43, fcos.py:27-85)."""
    D = cfg[cfg["feature_type"]]["feature_dim"]
    c1 = cfg["first_output_dim"]
    F_ = cfg["fpn_feature_dim"]
    H = hidden_dim
    s = [("query_encoder.embedding.weight", (vocab_size + 1, embed_dim))]
    for suf in ("", "_reverse"):
        s += [
            ("query_encoder.biLSTM.weight_ih_l0" + suf, (4 * H, embed_dim)),
            ("query_encoder.biLSTM.weight_hh_l0" + suf, (4 * H, H)),
            ("query_encoder.biLSTM.bias_ih_l0" + suf, (4 * H,)),
            ("query_encoder.biLSTM.bias_hh_l0" + suf, (4 * H,)),
        ]
    s += [
        ("query_encoder.textualAttention.W1.weight", (1, 1024)), ("query_encoder.textualAttention.W1.bias", (1,)),
        ("query_encoder.textualAttention.W2.weight", (1024, 2048)), ("query_encoder.textualAttention.W2.bias", (1024,)),
        ("query_encoder.textualAttention.W3.weight", (2048, 2048)), ("query_encoder.textualAttention.W3.bias", (2048,)),
        ("query_encoder.qInput.weight", (H, 4 * H)), ("query_encoder.qInput.bias", (H,)),
    ]
    for t in range(3):
        s += [("query_encoder.qInput%d.weight" % t, (2 * H, H)), ("query_encoder.qInput%d.bias" % t, (2 * H,))]
    s += [("query_encoder.cmd_inter2logits.weight", (1, 2 * H)), ("query_encoder.cmd_inter2logits.bias", (1,))]
    s += _conv_bn("backbone_net.forward_conv0", c1, D + 256, 3)
    s += _conv_bn("backbone_net.forward_conv1", 2 * c1, c1, 3)
    s += _conv_bn("backbone_net.forward_conv2", 4 * c1, 2 * c1, 3)
    for i, c_in in enumerate((256, 512, 1024), 1):
        s += _conv_bn("fpn.fpn_inner%d" % i, 512, c_in, 1)
        s += _conv_bn("fpn.fpn_layer%d" % i, 512, 512, 3)
    h = "fcos.head."
    s += _conv_bn(h + "cls_tower", F_, F_, 3, bias=True)
    s += _conv_bn(h + "bbox_tower", F_, F_, 3, bias=True)
    s += [
        (h + "cls_logits.weight", (cfg["fcos_num_class"] - 1, F_, 3)), (h + "cls_logits.bias", (cfg["fcos_num_class"] - 1,)),
        (h + "bbox_pred.weight", (2, F_, 3)), (h + "bbox_pred.bias", (2,)),
        (h + "centerness.weight", (1, F_, 3)), (h + "centerness.bias", (1,)),
    ]
    s += _conv_bn(h + "mix_fc", F_, 2 * F_, 1, bias=True)
    s += _conv_bn(h + "iou_scores", F_ // 2, F_, 3, bias=True)
    s += [(h + "iou_scores.3.weight", (1, F_ // 2, 1)), (h + "iou_scores.3.bias", (1,))]
    s += [(h + "scales.%d.scale" % l, (1,)) for l in range(3)]
    s += [("prop_fc.weight", (D, D)), ("prop_fc.bias", (D,)),
          ("position_transform.weight", (256, 3)), ("position_transform.bias", (256,)),
          ("qInput0.weight", (D, 1024)), ("qInput0.bias", (D,)),
          ("qInput1.weight", (c1, 1024)), ("qInput1.bias", (c1,)),
          ("qInput2.weight", (2 * c1, 1024)), ("qInput2.bias", (2 * c1,))]
    return s
