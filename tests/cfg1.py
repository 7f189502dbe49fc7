"""BASELINE config 1 driven through the REFERENCE's own ``tools/infer.py`` flow
(``evaluate`` -> ``eval_performance`` -> ``evalSEA``, tools/infer.py:136-155,56-133,332-403).

TEST INFRASTRUCTURE, shared by ``tests/golden/make_golden_config1.py`` (unmodified reference on the
CPU of the build container -> ``tests/golden/config1_sea.npz``) and the ``-m gpu`` drop-in tests
(same code with the B200 modules swapped in by ``dropin.install``).  Nothing here is imported by the
product.

Config 1: ``torch.manual_seed(0)``; ``UperNetForSemanticSegmentation(
"ConvNeXt-T_CVST", 21, None).eval()``; ``x = rand(2,3,512,512)``; ``y = randint(0,21,(2,512,512))``;
eps = 4/255, n_iter = 10 (stages 3/3/4), ``use_rs=True, early_stop=True, track_loss="ce-avg"``.
The fixture runs all three SEA losses so that ``evalSEA`` sees its three attacks; "Mask-ce only"
is the ``mask-ce-avg`` entry.

The random start is drawn by ``torch.rand_like`` on the device of ``x`` (semseg/attacker.py:292);
CPU and CUDA generators produce different streams, so both sides draw it from a seeded CPU
generator through ``seeded_rand_like`` -- a test-only patch of the RNG source, not of the code
under test.
"""
import contextlib
import os
import random
import tempfile
from functools import partial

import numpy as np
import torch

LOSSES = ["mask-ce-bal", "mask-ce-avg", "js-avg"]
N_CLS, EPS, N_ITER, SIZE, N_IMG = 21, 4.0, 10, 512, 2


def build_inputs(model_ctor, n_cls=N_CLS, size=SIZE, n_img=N_IMG):
    """Seed-0 model + images + labels, exactly in SURVEY 8d's order (CPU generator)."""
    torch.manual_seed(0)
    model = model_ctor("ConvNeXt-T_CVST", n_cls, None).eval()
    x = torch.rand(n_img, 3, size, size)
    y = torch.randint(0, n_cls, (n_img, size, size))
    return model, x, y


class SynthData(torch.utils.data.Dataset):
    """Items shaped like the reference's datasets: (img, target[H,W] int64, name)."""

    def __init__(self, x, y):
        self.x, self.y = x, y

    def __len__(self):
        return self.x.shape[0]

    def __getitem__(self, i):
        return self.x[i], self.y[i], f"img{i}"


@contextlib.contextmanager
def seeded_rand_like(seed=1234):
    g = torch.Generator().manual_seed(seed)
    real = torch.rand_like

    def fake(t, *a, **k):
        return torch.rand(t.shape, generator=g, dtype=torch.float32).to(device=t.device, dtype=t.dtype)

    torch.rand_like = fake
    try:
        yield
    finally:
        torch.rand_like = real


@contextlib.contextmanager
def cuda_means_cpu():
    """CPU-only run of code that hard-codes ``.to("cuda")`` / ``.cuda()`` (tools/infer.py:82,143-144)."""
    real_to, real_cuda = torch.Tensor.to, torch.Tensor.cuda

    def to(self, *a, **k):
        a = tuple("cpu" if (isinstance(v, str) and v.startswith("cuda")) else v for v in a)
        if isinstance(k.get("device"), str) and k["device"].startswith("cuda"):
            k["device"] = "cpu"
        return real_to(self, *a, **k)

    torch.Tensor.to = to
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.to, torch.Tensor.cuda = real_to, real_cuda


class _Args:
    norm = "Linf"


def sea_flow(TI, model, x, y, weights, losses=LOSSES, n_cls=N_CLS, eps=EPS, n_iter=N_ITER, bs=N_IMG,
             rand_seed=1234, record=None):
    """The body of the reference's ``__main__`` loop (tools/infer.py:313-403) on a synthetic
    dataset, calling whatever ``TI.attacker / TI.evaluate / TI.eval_performance / TI.evalSEA``
    currently are (the reference's own objects, or the B200 ones after ``dropin.install``).

    Returns a dict of plain python / numpy results."""
    data = SynthData(x, y)
    loader = torch.utils.data.DataLoader(data, batch_size=bs, shuffle=False, num_workers=0)
    out = {}
    clean_stats, _ = TI.eval_performance(model, loader, n_batches=-1, n_cls=n_cls, ignore_index=-1)
    out["clean"] = clean_stats
    loss_wise, accs = [], {}
    for loss_ in losses:
        acc_rec = []
        inner = partial(TI.attacker.apgd_largereps, norm="Linf", eps=eps / 255.0, n_iter=n_iter, n_restarts=1,
                        use_rs=True, loss=loss_, verbose=False, track_loss="ce-avg", log_path=None,
                        num_classes=n_cls, early_stop=True)
        # the drop-in's evaluate() recognises ITS apgd_largereps behind a functools.partial (return_pred);
        # keep that shape and record the per-image accuracy through a thin subclass of partial
        attack_fn = _RecordingPartial(inner, acc_rec)
        with seeded_rand_like(rand_seed):
            adv_loader = TI.evaluate(loader, model, attack_fn, -1, _Args(), weights)
        adv_stats, l_outs = TI.eval_performance(model, adv_loader, -1, n_cls=n_cls, ignore_index=-1)
        out[loss_] = adv_stats
        accs[loss_] = torch.cat(acc_rec).numpy()
        loss_wise.append(l_outs.detach().cpu())
        if record is not None:
            record[loss_] = adv_loader
    out["acc"] = accs
    out["l_outs"] = torch.stack(loss_wise).numpy()
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "test_results"))
        save = {"seed": 225, "worst_Acc": 0, "worst_Acc_indiv": 0, "final_miou": 0}
        random.seed(225)  # module-import seeding of tools/worse_only.py:14-15
        np.random.seed(225)
        ev = TI.evalSEA(val_data=data, l_outs=loss_wise, eps=eps, n_cls=n_cls, addendum="SEA_cfg1",
                        saveDir=d, saveDict=save, modelName="UperNet_ConvNeXt-T_CVST")
        ev.worse_case_eval(bs=bs, n_batches=-1)
        ev.worst_case_miou()
    out["worst_Acc"] = float(save["worst_Acc"])
    out["worst_Acc_indiv"] = np.asarray(save["worst_Acc_indiv"], dtype=np.float32)
    out["final_miou"] = float(save["final_miou"])
    return out


class _RecordingPartial(partial):
    """functools.partial of apgd_largereps that also notes the returned per-image accuracy."""

    def __new__(cls, inner, rec):
        self = super().__new__(cls, inner.func, *inner.args, **inner.keywords)
        self._rec = rec
        return self

    def __call__(self, *a, **k):
        res = super().__call__(*a, **k)
        self._rec.append(res[2].detach().cpu())
        return res
