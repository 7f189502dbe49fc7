"""Swap the B200 modules into a checkout of the reference and run ITS drivers on them.

    import robseg_b200.dropin as dropin        # (alias from __graft_entry__.load_package)
    dropin.install("/path/to/Robust-Segmentation")
    import tools.infer                         # now binds to the B200 attacker / evalSEA
    ns = dropin.run_infer_main(["--cfg", "configs/ade20k_convnext.yaml", "--eps", "8"])

``install`` imports the reference's own ``semseg`` package (models, datasets, configs stay
theirs), then rebinds exactly the names SURVEY.md section 8b lists: the ``semseg.attacker``
module, ``semseg.val.{Pgd_Attack, Pgd_Attack_1, evaluate}``, ``semseg.metrics.Metrics``,
``semseg.losses.{CrossEntropy, get_loss}``, ``tools.worse_only.evalSEA`` and -- so that the
reference's own SEA driver gets the device-resident bookkeeping -- ``tools.infer.{attacker,
evalSEA, evaluate, eval_performance, check_imgs}`` (tools/infer.py:16,22,39-155).
``uninstall`` puts every original back.

``run_infer_main`` executes the reference's ``tools/infer.py`` ``__main__`` block (:220-413)
unmodified, inside the rebound ``tools.infer`` namespace.  ``runpy.run_module("tools.infer",
run_name="__main__")`` would re-execute the module's ``def evaluate`` / ``def eval_performance``
and shadow the rebinding, so only the main block is compiled from the file's own source.
"""
import ast
import importlib
import os
import sys
import types

_saved = []  # (container, key, had_it, old_value, is_sys_modules)


def shim_missing_deps():
    """Make a reference checkout importable where ``timm==0.6.5`` / ``autoattack`` (fra31/auto-attack
    @a392200, requirements.txt:3,68) are not installed.  Only names the reference touches at import
    time are provided: a print/append ``Logger`` and three norm helpers
    (semseg/attacker.py:6), ``DropPath`` / ``trunc_normal_`` / ``register_model`` and a few ``None``
    factories (semseg/models/*).  Packages that do import are left alone."""
    import torch.nn as nn

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    def importable(name):
        if name in sys.modules:
            return True
        try:
            importlib.import_module(name)
            return True
        except Exception:
            return False

    if not importable("autoattack.other_utils"):
        def _flat(x):
            return x.reshape(x.shape[0], -1)

        def _keep(z, x, keepdim):
            return z.view(-1, *[1] * (x.dim() - 1)) if keepdim else z

        class Logger:
            def __init__(self, log_path=None):
                self.log_path = log_path

            def log(self, str_to_log):
                print(str_to_log)
                if self.log_path is not None:
                    with open(self.log_path, "a") as f:
                        f.write(str_to_log + "\n")

        ou = mod("autoattack.other_utils",
                 L0_norm=lambda x: _flat(x != 0.0).sum(-1),
                 L1_norm=lambda x, keepdim=False: _keep(_flat(x.abs()).sum(-1), x, keepdim),
                 L2_norm=lambda x, keepdim=False: _keep(_flat(x ** 2).sum(-1).sqrt(), x, keepdim),
                 Logger=Logger)
        mod("autoattack", other_utils=ou)

    if not importable("timm.models.layers"):
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
        vit = mod("timm.models.vision_transformer", _create_vision_transformer=None, default_cfgs={},
                  _load_weights=None)
        models = mod("timm.models", layers=layers, registry=registry, vision_transformer=vit)
        mod("timm", models=models, optim=mod("timm.optim", create_optimizer=None),
            scheduler=mod("timm.scheduler", create_scheduler=None))


def _rebind(container, key, value, is_modules=False):
    if is_modules:
        _saved.append((container, key, key in container, container.get(key), True))
        container[key] = value
    else:
        _saved.append((container, key, hasattr(container, key), getattr(container, key, None), False))
        setattr(container, key, value)


def install(reference_root=None, shims=True, accelerate_models=False):
    """Rebind the hot-path names of an importable reference checkout to the B200 modules.
    Works whether or not ``tools.infer`` / ``tools.train_rob_seg`` were imported before.

    ``accelerate_models=True`` also wraps the model factories the drivers call
    (``UperNetForSemanticSegmentation``, ``create_segmenter``: tools/infer.py:256-264,
    tools/train_rob_seg.py) so that every model they build has its bilinear up-samplings on the
    robseg kernels (:func:`accelerate`).  The model classes themselves stay the
    reference's."""
    from .semseg import attacker, losses, metrics, val
    from .tools import infer, worse_only

    if _saved:
        uninstall()
    if shims:
        shim_missing_deps()
    if reference_root is not None and reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    ref = importlib.import_module("semseg")
    _rebind(sys.modules, "semseg.attacker", attacker, is_modules=True)
    _rebind(ref, "attacker", attacker)
    for modname, names, src in (
        ("semseg.val", ("Pgd_Attack", "Pgd_Attack_1", "evaluate"), val),
        ("semseg.metrics", ("Metrics",), metrics),
        ("semseg.losses", ("CrossEntropy", "get_loss"), losses),
        ("tools.worse_only", ("evalSEA",), worse_only),
        ("tools.infer", ("evalSEA",), worse_only),
        ("tools.infer", ("evaluate", "eval_performance", "check_imgs"), infer),
        ("tools.train_rob_seg", ("Pgd_Attack", "evaluate"), val),
        ("tools.train_rob_seg", ("get_loss",), losses),
    ):
        if modname.startswith("tools.train") and modname not in sys.modules:
            continue  # the trainer binds at import time: importing it after install() is enough
        try:
            m = importlib.import_module(modname)
        except Exception:  # optional pieces of the reference may not import (missing deps)
            continue
        for n in names:
            _rebind(m, n, getattr(src, n))
    for modname in ("tools.infer", "tools.train_rob_seg"):
        m = sys.modules.get(modname)
        if m is not None and hasattr(m, "attacker"):
            _rebind(m, "attacker", attacker)
        if m is not None and accelerate_models:
            for n in ("UperNetForSemanticSegmentation", "create_segmenter"):
                if hasattr(m, n):
                    _rebind(m, n, _accelerating(getattr(m, n)))
    return attacker


def _accelerating(factory):
    def build(*a, **k):
        return accelerate(factory(*a, **k))

    build.__name__ = getattr(factory, "__name__", "build")
    build.__wrapped__ = factory
    return build


def uninstall():
    """Undo ``install``: every rebound name gets the reference's own object back."""
    while _saved:
        container, key, had, old, is_modules = _saved.pop()
        if is_modules:
            if had:
                container[key] = old
            else:
                container.pop(key, None)
        elif had:
            setattr(container, key, old)
        else:
            delattr(container, key)


def run_infer_main(argv, overrides=None):
    """Run the reference's SEA driver -- the ``if __name__ == "__main__":`` block of ITS
    ``tools/infer.py`` (:220-413), compiled from the checkout's own source -- in the rebound
    ``tools.infer`` namespace.  ``install()`` must have been called.  ``overrides`` are extra
    names placed in that namespace first (tests swap ``get_data`` for a synthetic dataset).
    Returns the namespace after the run (``clean_stats``, ``evall.saveDict``, ...)."""
    if not _saved:
        raise RuntimeError("dropin.install(reference_root) first")
    mod = importlib.import_module("tools.infer")
    with open(mod.__file__) as f:
        tree = ast.parse(f.read(), mod.__file__)
    main = [n for n in tree.body if isinstance(n, ast.If) and isinstance(n.test, ast.Compare)
            and getattr(n.test.left, "id", None) == "__name__"]
    if len(main) != 1:
        raise RuntimeError(f"{mod.__file__}: expected exactly one __main__ block")
    code = compile(ast.Module(body=main[0].body, type_ignores=[]), mod.__file__, "exec")
    ns = dict(vars(mod))
    ns.update(overrides or {})
    old_argv = sys.argv
    sys.argv = ["tools.infer"] + list(argv)
    try:
        exec(code, ns)
    finally:
        sys.argv = old_argv
    return ns


_MISSING = object()


def run_train_main(cfg, overrides=None, gpu=0, require_install=True):
    """Run the reference's PIR-AT trainer -- ``Trainer(gpu, cfg)`` + ``Trainer.main()`` of ITS
    ``tools/train_rob_seg.py`` (:63-145 model / DDP / loaders / optimiser, :270-352 the training loop with the
    eval-mode inner attack, :353-440 periodic ``evaluate`` + checkpoints), unmodified -- in the rebound
    ``tools.train_rob_seg`` namespace, i.e. with ``Pgd_Attack`` / ``attacker.apgd_train`` / ``evaluate`` /
    ``get_loss`` being the B200 modules.  ``install()`` must have been called (the trainer binds these names when it
    is imported; ``install`` rebinds them if that happened earlier).  ``overrides`` are extra names set in that
    module for the duration of the run (tests swap ``get_segmentation_dataset`` for a synthetic dataset, and set
    the reference's own classes back for the comparison run with ``require_install=False``).  The process group the
    trainer creates (``setup_distributed``, world size = visible GPUs) is destroyed afterwards.  Returns the
    ``Trainer`` (``save_path``, ``model``, ``optimizer``, ...)."""
    if require_install and not _saved:
        raise RuntimeError("dropin.install(reference_root) first")
    import torch.distributed as dist

    mod = importlib.import_module("tools.train_rob_seg")
    saved = {}
    for k, v in (overrides or {}).items():
        saved[k] = getattr(mod, k, _MISSING)
        setattr(mod, k, v)
    trainer = None
    try:
        trainer = mod.Trainer(gpu=gpu, cfg=cfg)
        trainer.main()
        return trainer
    finally:
        for k, v in saved.items():
            if v is _MISSING:
                delattr(mod, k)
            else:
                setattr(mod, k, v)
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()


def _train_worker(gpu, cfg, reference_root, shims):
    """Body of one spawned trainer process: the rebinding lives in module state, so a fresh interpreter has to
    install it again before ITS import of tools.train_rob_seg builds the Trainer."""
    if reference_root is not None and os.getcwd() != reference_root:
        os.chdir(reference_root)  # the reference opens ./configs/... relative to its checkout (semseg/utils/utils.py:259)
    install(reference_root, shims=shims)
    run_train_main(cfg, gpu=gpu)


def launch_train(cfg, reference_root, world_size=None, shims=True):
    """``Trainer.launch_from_args(world_size, cfg)`` (tools/train_rob_seg.py:455-462: one spawned process per GPU) with
    the drop-in installed in every process.  ``torch.multiprocessing.spawn`` starts fresh interpreters, which import the
    reference's modules unpatched -- calling the reference's own launcher after ``install()`` would therefore train on
    the reference's attack in the children.  This launcher spawns ``__graft_entry__.dropin_train_worker`` instead (a
    name a fresh interpreter can import: this package's directory name is not an identifier), which loads the package,
    installs the drop-in and only then builds the reference's ``Trainer``.  ``world_size`` defaults to the visible GPUs
    (the Trainer sizes its process group the same way, :76)."""
    import torch
    import torch.multiprocessing as mp

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    ge = importlib.import_module("__graft_entry__")
    n = torch.cuda.device_count() if world_size is None else int(world_size)
    if n < 1:
        raise RuntimeError("launch_train needs at least one CUDA device (no CPU fallback)")
    mp.spawn(ge.dropin_train_worker, args=(cfg, reference_root, shims), nprocs=n, join=True)


def fast_logit_upsample(model, head=False, fuse_loss=False):
    """Route the final logit up-sampling of a reference model through robseg's kernels.

    ``fuse_loss=True`` additionally gives the model a ``forward_lowres(x)`` method returning the
    logits BEFORE that up-sampling; ``apgd_train`` then runs the loss kernel that interpolates on the
    fly (``ops.loss_upsampled_fwd_bwd``) and back-propagates from the low-resolution gradient, so
    the [B,C,H,W] logits and their gradient never exist during an attack.

    Works for models shaped like the reference's ``UperNetForSemanticSegmentation``
    (semseg/models/uperforseg.py:382-439: ``backbone`` -> ``decode_head`` -> bilinear
    up-sampling to the input size).  Only the eval-mode forward the attack uses is replaced;
    the training forward (loss + aux head) is left alone.

    ``head=True`` also routes the bilinear up-samplings inside the decode head (PSP, top-down
    path, pyramid fusion: uperforseg.py:193-198,282-303), which call
    ``nn.functional.interpolate`` by name: that name is rebound to ``ops.interpolate`` for the
    duration of the head's forward only (other modes / dtypes / devices fall through to the
    stock function)."""
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

    def forward_lowres(self, input):
        if self.training or not input.is_cuda:
            return None
        feats = self.backbone(input)
        if head:
            with ops.patched_interpolate():
                return self.decode_head(feats).float()
        return self.decode_head(feats).float()

    model.forward = types.MethodType(forward, model)
    if fuse_loss:
        model.forward_lowres = types.MethodType(forward_lowres, model)
    return model


def fast_interpolate(model, fuse_loss=False):
    """Any other reference model (``SegMenter``: semseg/models/segmenter.py:193-231, x16 bilinear
    up-sampling of the class masks at :228; ``PSPNet``): its ``F.interpolate`` calls resolve
    ``torch.nn.functional.interpolate`` at call time, so the whole eval-mode forward runs with that
    name rebound to ``ops.interpolate`` (bilinear / align_corners=False / fp32 / CUDA -> robseg
    kernels, everything else -> the stock function)."""
    from . import ops

    stock_forward = model.forward

    def forward(self, *a, **k):
        if self.training:
            return stock_forward(*a, **k)
        with ops.patched_interpolate():
            return stock_forward(*a, **k)

    def forward_lowres(self, im):
        # SegMenter.forward up to the class masks (semseg/models/segmenter.py:214-226); the x16
        # interpolation of :228 is what the fused loss kernel takes over.  Padded inputs (:216) are
        # cropped after the interpolation by the reference, which the fused kernel does not model.
        H, W = im.shape[2:]
        if self.training or not im.is_cuda or H % self.patch_size or W % self.patch_size:
            return None
        with ops.patched_interpolate():
            x = self.encoder(im, pre_neck=True)
            n_extra = 0 if "SAM" in self.backbone else 1 + self.encoder.distilled
            return self.decoder(x[:, n_extra:], (H, W)).float()

    model.forward = types.MethodType(forward, model)
    if fuse_loss and all(hasattr(model, n) for n in ("encoder", "decoder", "patch_size", "backbone")):
        model.forward_lowres = types.MethodType(forward_lowres, model)
    return model


def accelerate(model, head=True, fuse_loss="auto"):
    """``fast_logit_upsample`` for UperNet-shaped models, ``fast_interpolate`` otherwise.

    ``fuse_loss="auto"`` follows the measurements (profiles/r02_fused_upsample_loss.md, B200, 16 x 150 x
    512^2): interpolating inside the loss kernel is 3.8x faster than the three-kernel path at x16
    (SegMenter: 0.61 vs 2.34 ms per iteration) and 2.6x at x8, but only 1.2x at x4 (UperNet: 1.47 vs
    1.81 ms) where the separate kernels already stream at 83-93 % of the HBM roofline -- so it is
    switched on for SegMenter-shaped models and left opt-in (``fuse_loss=True``) for UperNet."""
    if hasattr(model, "backbone") and hasattr(model, "decode_head"):
        return fast_logit_upsample(model, head=head, fuse_loss=fuse_loss is True)
    return fast_interpolate(model, fuse_loss=bool(fuse_loss))


def reference_model(kind, variant, n_cls, image_size=512):
    """Random-init instance of one of the REFERENCE's model classes (the checkout must be importable:
    ``install`` / ``shim_missing_deps`` + sys.path).  ``kind="upernet"``:
    ``UperNetForSemanticSegmentation(f"ConvNeXt-{variant}_CVST", n_cls, None)`` (tools/infer.py:262);
    ``kind="segmenter"``: ``SegMenter`` over ``VisionTransformer`` + ``MaskTransformer`` built
    directly with the arguments ``load_config_segmenter`` / ``create_vit`` / ``create_decoder`` derive
    for ``vit_{small,base,large}_patch16`` (semseg/utils/utils.py:258-277, semseg/models/segmenter.py:
    265-342), because ``create_vit`` insists on a checkpoint file (:299)."""
    if kind == "upernet":
        from semseg.models import UperNetForSemanticSegmentation

        return UperNetForSemanticSegmentation(f"ConvNeXt-{variant}_CVST", n_cls, None)
    if kind != "segmenter":
        raise ValueError(kind)
    from semseg.models.segmenter import MaskTransformer, SegMenter, VisionTransformer

    name, d, heads, layers = {"S": ("vit_small_patch16_224", 384, 6, 12), "B": ("vit_base_patch16_384", 768, 12, 12),
                              "L": ("vit_large_patch16_384", 1024, 16, 24)}[variant]
    enc = VisionTransformer(image_size=(image_size, image_size), patch_size=16, n_layers=layers, d_model=d,
                            d_ff=4 * d, n_heads=heads, n_cls=1000, dropout=0.0, drop_path_rate=0.1, distilled=False)
    dec = MaskTransformer(n_cls=n_cls, patch_size=16, d_encoder=d, n_layers=2, n_heads=d // 64, d_model=d,
                          d_ff=4 * d, drop_path_rate=0.0, dropout=0.1)
    return SegMenter(enc, dec, n_cls=n_cls, backbone=name)
