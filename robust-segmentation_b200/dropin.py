"""Swap the B200 modules into a checkout of the reference.

    import robseg_b200.dropin as dropin        # (alias from __graft_entry__.load_package)
    dropin.install("/path/to/Robust-Segmentation")
    import tools.infer                         # now binds to the B200 attacker / evalSEA

``install`` imports the reference's own ``semseg`` package (models, datasets, configs stay
theirs), then rebinds exactly the names SURVEY.md section 8b lists: the ``semseg.attacker``
module, ``semseg.val.{Pgd_Attack, Pgd_Attack_1, evaluate}``, ``semseg.metrics.Metrics``,
``semseg.losses.{CrossEntropy, get_loss}`` and ``tools.worse_only.evalSEA``.
"""
import importlib
import sys


def install(reference_root=None):
    from .semseg import attacker, losses, metrics, val
    from .tools import worse_only

    if reference_root is not None and reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    ref = importlib.import_module("semseg")
    sys.modules["semseg.attacker"] = attacker
    ref.attacker = attacker
    for modname, names, src in (
        ("semseg.val", ("Pgd_Attack", "Pgd_Attack_1", "evaluate"), val),
        ("semseg.metrics", ("Metrics",), metrics),
        ("semseg.losses", ("CrossEntropy", "get_loss"), losses),
        ("tools.worse_only", ("evalSEA",), worse_only),
    ):
        try:
            m = importlib.import_module(modname)
        except Exception:  # optional pieces of the reference may not import (missing deps)
            continue
        for n in names:
            setattr(m, n, getattr(src, n))
    return attacker


def fast_logit_upsample(model, head=False):
    """Route the final logit up-sampling of a reference model through robseg's kernels.

    Works for models shaped like the reference's ``UperNetForSemanticSegmentation``
    (semseg/models/uperforseg.py:382-439: ``backbone`` -> ``decode_head`` -> bilinear
    up-sampling to the input size).  Only the eval-mode forward the attack uses is replaced;
    the training forward (loss + aux head) is left alone.

    ``head=True`` also routes the bilinear up-samplings inside the decode head (PSP, top-down
    path, pyramid fusion: uperforseg.py:193-198,282-303), which call
    ``nn.functional.interpolate`` by name: that name is rebound to ``ops.interpolate`` for the
    duration of the head's forward only (other modes / dtypes / devices fall through to the
    stock function)."""
    import types

    from . import ops

    if not (hasattr(model, "backbone") and hasattr(model, "decode_head")):
        raise TypeError("fast_logit_upsample expects a model with .backbone and .decode_head")
    stock_forward = model.forward

    def forward(self, input=None, lbl=None):
        if lbl is not None or self.training or not input.is_cuda:
            return stock_forward(input, lbl)
        feats = self.backbone(input)
        if head:
            with ops.patched_interpolate():
                low = self.decode_head(feats)
        else:
            low = self.decode_head(feats)
        return ops.upsample_bilinear(low.float(), input.shape[2:])

    model.forward = types.MethodType(forward, model)
    return model
