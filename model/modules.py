"""Parameter containers with the reference's module tree, initialisers and state_dict keys (reference
model/backbone.py:5-15, model/FPN.py:26-43, model/fcos.py:19-85,114-124, model/basic_blocks.py:5-33).

These nn.Modules only HOLD parameters/buffers (so `.cuda()`, `state_dict()`, `named_parameters()`, `nn.DataParallel` and
the optimizer see exactly what the reference exposes); their torch `forward` is never used -- the math runs in
drn_b200.dense.DensePath on the CUDA kernels.
"""
import math

import torch
from torch import nn


def _no_forward(self, *a, **k):
    raise RuntimeError("parameter container: the dense path runs through drn_b200.dense.DensePath (CUDA kernels), "
                       "there is no torch forward / CPU fallback")


def conv_bn_relu(cin, cout, k, stride=1, bias=False):
    """conv_with_kaiming_uniform(use_bn=True, use_relu=True) of the reference (basic_blocks.py:5-33)."""
    conv = nn.Conv1d(cin, cout, kernel_size=k, stride=stride, padding=(k - 1) // 2, bias=bias)
    nn.init.kaiming_uniform_(conv.weight, a=1)
    return nn.Sequential(conv, nn.BatchNorm1d(cout), nn.ReLU(inplace=True))


class Backbone(nn.Module):
    def __init__(self, channels_list):
        super().__init__()
        self.num_layers = len(channels_list)
        for idx, (cin, cout, k, s) in enumerate(channels_list):
            self.add_module("forward_conv%d" % idx, conv_bn_relu(cin, cout, k, s))
    forward = _no_forward


class FPN(nn.Module):
    def __init__(self, in_channels_list, out_channels):
        super().__init__()
        for idx, cin in enumerate(in_channels_list, 1):
            self.add_module("fpn_inner%d" % idx, conv_bn_relu(cin, out_channels, 1))
            self.add_module("fpn_layer%d" % idx, conv_bn_relu(out_channels, out_channels, 3, 1))
    forward = _no_forward


class Scale(nn.Module):
    def __init__(self, init_value=1.0):
        super().__init__()
        self.scale = nn.Parameter(torch.FloatTensor([init_value]))
    forward = _no_forward


class FCOSHead(nn.Module):
    def __init__(self, cfg, in_channels):
        super().__init__()
        num_classes = cfg["fcos_num_class"] - 1
        if cfg["fcos_conv_layers"] != 1:
            raise NotImplementedError("the B200 path fuses the 1-layer towers of the reference config (fcos_conv_layers=1)")
        C_ = in_channels
        self.cls_tower = nn.Sequential(nn.Conv1d(C_, C_, 3, 1, 1), nn.BatchNorm1d(C_), nn.ReLU())
        self.bbox_tower = nn.Sequential(nn.Conv1d(C_, C_, 3, 1, 1), nn.BatchNorm1d(C_), nn.ReLU())
        self.cls_logits = nn.Conv1d(C_, num_classes, 3, 1, 1)
        self.bbox_pred = nn.Conv1d(C_, 2, 3, 1, 1)
        self.centerness = nn.Conv1d(C_, 1, 3, 1, 1)  # built, never applied (fcos.py:53-56,97)
        self.mix_fc = nn.Sequential(nn.Conv1d(2 * C_, C_, 1, 1), nn.BatchNorm1d(C_), nn.ReLU())
        self.iou_scores = nn.Sequential(nn.Conv1d(C_, C_ // 2, 3, 1, 1), nn.BatchNorm1d(C_ // 2), nn.ReLU(),
                                        nn.Conv1d(C_ // 2, 1, 1, 1))
        for modules in [self.cls_tower, self.bbox_tower, self.cls_logits, self.bbox_pred, self.centerness, self.iou_scores,
                        self.mix_fc]:
            for l in modules.modules():
                if isinstance(l, nn.Conv1d):
                    torch.nn.init.normal_(l.weight, std=0.01)
                    torch.nn.init.constant_(l.bias, 0)
        prior_prob = cfg["fcos_prior_prob"]
        torch.nn.init.constant_(self.cls_logits.bias, -math.log((1 - prior_prob) / prior_prob))
        self.scales = nn.ModuleList([Scale(1.0) for _ in range(3)])
    forward = _no_forward


class _LossEvaluatorStub:
    """Keeps the attribute the reference exposes (`loss_evaluator.total_points`, loss.py:38,193) without its leak."""

    def __init__(self):
        self.total_points = []


class FCOSModule(nn.Module):
    def __init__(self, cfg, in_channels):
        super().__init__()
        self.head = FCOSHead(cfg, in_channels)
        self.is_first_stage = cfg["is_first_stage"]
        self.fpn_strides = cfg["fpn_stride"]
        self.loss_evaluator = _LossEvaluatorStub()
    forward = _no_forward
