"""Import shims so the UNMODIFIED reference can be imported in the build container.

TEST INFRASTRUCTURE ONLY (used by tests/golden/make_golden.py and, when a copy of the
reference travels under baseline/_ref, by ``bench.py --impl reference``).  The reference
needs two packages that are not installed here (requirements.txt:3 ``autoattack`` pinned at
fra31/auto-attack@a392200, requirements.txt:68 ``timm==0.6.5``).  Only trivially small
pieces of them are reachable on the Linf SEA path: three norm helpers and a
print logger from ``autoattack.other_utils`` (semseg/attacker.py:6) and a handful of names
``semseg.models`` pulls from ``timm`` at import time.  They are restated here from their
published behaviour and injected into ``sys.modules`` before the reference is imported.
"""
import sys
import types


def install():
    import torch
    import torch.nn as nn

    if "autoattack.other_utils" in sys.modules and "timm.models.layers" in sys.modules:
        return

    def _flat(x):
        return x.reshape(x.shape[0], -1)

    def _keep(z, x, keepdim):
        return z.view(-1, *[1] * (x.dim() - 1)) if keepdim else z

    def L1_norm(x, keepdim=False):
        return _keep(_flat(x.abs()).sum(-1), x, keepdim)

    def L2_norm(x, keepdim=False):
        return _keep(_flat(x ** 2).sum(-1).sqrt(), x, keepdim)

    def L0_norm(x):
        return _flat(x != 0.0).sum(-1)

    class Logger:
        def __init__(self, log_path=None):
            self.log_path = log_path

        def log(self, str_to_log):
            print(str_to_log)
            if self.log_path is not None:
                with open(self.log_path, "a") as f:
                    f.write(str_to_log + "\n")

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    ou = mod("autoattack.other_utils", L0_norm=L0_norm, L1_norm=L1_norm, L2_norm=L2_norm,
             Logger=Logger)
    mod("autoattack", other_utils=ou)

    class DropPath(nn.Module):
        def __init__(self, drop_prob=0.0):
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            if self.drop_prob == 0.0 or not self.training:
                return x
            keep = 1.0 - self.drop_prob
            mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
            return x * mask / keep

    layers = mod("timm.models.layers", DropPath=DropPath, trunc_normal_=nn.init.trunc_normal_)
    registry = mod("timm.models.registry", register_model=lambda fn: fn)
    vit = mod("timm.models.vision_transformer", _create_vision_transformer=None,
              default_cfgs={}, _load_weights=None)
    models = mod("timm.models", layers=layers, registry=registry, vision_transformer=vit)
    optim = mod("timm.optim", create_optimizer=None)
    sched = mod("timm.scheduler", create_scheduler=None)
    mod("timm", models=models, optim=optim, scheduler=sched)
    _ = torch


def import_reference(root="/root/reference"):
    """Returns the reference's (attacker, val, metrics, losses, worse_only) modules."""
    import importlib

    install()
    if root not in sys.path:
        sys.path.insert(0, root)
    names = ["semseg.attacker", "semseg.val", "semseg.metrics", "semseg.losses",
             "tools.worse_only"]
    return tuple(importlib.import_module(n) for n in names)
