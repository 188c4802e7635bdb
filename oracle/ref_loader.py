"""TEST INFRASTRUCTURE ONLY.  Imports the UNMODIFIED reference (`/root/reference`) on CPU through the
import shims in oracle/shims (recipe: SURVEY.md Appendix B).  Only oracle/make_goldens.py and the
CPU tests that pin the oracle may use this; it cannot run on the GPU box (no /root/reference there).
"""
import os
import sys

import torch

REFERENCE_ROOT = os.environ.get("DRN_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "main_model.py"))


def import_reference():
    """Returns the reference's `model.main_model` module.  The reference package is called `model`,
    the same name as this repo's drop-in package, so it must be imported in a process where the
    repo's `model` has not been imported (make_goldens.py runs standalone)."""
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    if "model" in sys.modules and not sys.modules["model"].__file__.startswith(REFERENCE_ROOT):
        raise RuntimeError("the repo's own `model` package is already imported in this process")
    sys.dont_write_bytecode = True  # reference dir is read-only
    for p in (REFERENCE_ROOT, _SHIMS):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, REFERENCE_ROOT)
    sys.path.insert(0, _SHIMS)
    if not torch.cuda.is_available():
        # loss.py:239 and inference.py:193-196 call .cuda() unconditionally
        torch.Tensor.cuda = lambda self, *a, **k: self
    import model.main_model as mm  # noqa: E402
    return mm


def reference_config(cfg):
    """CPU focal loss quirk: sigmoid_focal_loss_cpu does gamma[0] / alpha[0]
    (model/layers/sigmoid_focal_loss.py:42-43) so the CPU oracle passes 1-element lists."""
    from argparse import Namespace
    cfg = dict(cfg)
    if not isinstance(cfg["fcos_loss_gamma"], (list, tuple)):
        cfg["fcos_loss_gamma"] = [cfg["fcos_loss_gamma"]]
        cfg["fcos_loss_alpha"] = [cfg["fcos_loss_alpha"]]
    return Namespace(**cfg)


def build_reference_model(cfg, vocab_size=1301):
    mm = import_reference()
    model = mm.mainModel(vocab_size, reference_config(cfg))
    if cfg.get("is_first_stage"):
        # main.py:126-128
        for name, p in model.named_parameters():
            if "iou_scores" in name or "mix_fc" in name:
                p.requires_grad = False
    return model
