"""Generate the golden fixtures in tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference has no tests or golden vectors, so parity is pinned to
outputs of the reference's own functions on small seeded inputs.  Fixtures are small (a
few hundred KB in total) and committed; nothing at test / bench time reads
/root/reference.  Each fixture stores the inputs and what the reference returned.
"""
import os
import random
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_shims  # noqa: E402

A, V, M, L, W = ref_shims.import_reference("/root/reference")
torch.set_num_threads(1)


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, {k: v.shape for k, v in out.items()})


# ------------------------------------------------------------------ (i) losses
def make_logits(B, C, H, Wd, seed, sigma=3.0, frac_correct=0.5, frac_ignore=0.1, ties=True):
    g = torch.Generator().manual_seed(seed)
    z = sigma * torch.randn(B, C, H, Wd, generator=g)
    y = torch.randint(0, C, (B, H, Wd), generator=g)
    am = z.argmax(1)
    use = torch.rand(B, H, Wd, generator=g) < frac_correct
    y = torch.where(use, am, y)
    if ties:  # exact ties of the maximum between two channels at a few pixels
        for b in range(B):
            for k in range(3):
                h, w = int(torch.randint(0, H, (1,), generator=g)), int(torch.randint(0, Wd, (1,), generator=g))
                c0, c1 = sorted(torch.randperm(C, generator=g)[:2].tolist())
                top = z[b, :, h, w].max() + 0.5
                z[b, c0, h, w] = top
                z[b, c1, h, w] = top
                y[b, h, w] = c0 if k % 2 == 0 else c1
    ign = torch.rand(B, H, Wd, generator=g) < frac_ignore
    y = torch.where(ign, torch.full_like(y, -1), y)
    return z, y


def golden_losses():
    for tag, (B, C, H, Wd, seed, fi) in {
        "c7": (2, 7, 5, 6, 1, 0.1),
        "c21": (2, 21, 9, 8, 2, 0.0),
        "c151": (1, 151, 4, 7, 3, 0.15),
    }.items():
        z, y = make_logits(B, C, H, Wd, seed, frac_ignore=fi)
        g = torch.Generator().manual_seed(100 + seed)
        w = 0.5 + torch.rand(C, generator=g)
        out = dict(logits=z, labels=y, weights=w)
        mask_bg = 1 - (y == -1).float()
        for kind in ["mask-ce-avg", "mask-ce-bal", "js-avg", "ce-avg"]:
            zz = z.clone().requires_grad_()
            fn = A.criterion_dict[kind]
            lp = fn(zz, y) if kind == "ce-avg" else fn(zz, y, w)
            li = A.pixel_to_img_loss(lp, mask_bg)
            (gz,) = torch.autograd.grad(li.sum(), [zz])
            # arbitrary upstream gradient for the per-pixel criterion (compat op backward)
            zz2 = z.clone().requires_grad_()
            lp2 = fn(zz2, y) if kind == "ce-avg" else fn(zz2, y, w)
            up = torch.rand(lp2.shape, generator=g)
            (gz2,) = torch.autograd.grad((lp2 * up).sum(), [zz2])
            k = kind.replace("-", "_")
            out.update({f"{k}__loss_pix": lp, f"{k}__loss_img": li, f"{k}__dlogits": gz,
                        f"{k}__upstream": up, f"{k}__dlogits_up": gz2})
        pred = z.max(1)[1]
        out["pred"] = pred
        out["acc_step0"] = (pred == y).float().view(B, -1).mean(-1)
        pr = pred == y
        pr[y == -1] = True
        out["acc_loop"] = pr.float().view(B, -1).mean(-1)
        save(f"loss_{tag}", **out)


# ------------------------------------------------------------------ tiny consumer model
class TinySeg(torch.nn.Module):
    def __init__(self, C, seed):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.c1 = torch.nn.Conv2d(3, 8, 3, padding=1)
        self.c2 = torch.nn.Conv2d(8, C, 3, padding=1)
        with torch.no_grad():
            for p in self.parameters():
                p.copy_(torch.randn(p.shape, generator=g) * (0.6 if p.dim() > 1 else 0.1))

    def forward(self, x):
        return self.c2(torch.tanh(self.c1(x - 0.5)))


class Recorder(torch.nn.Module):
    def __init__(self, m):
        super().__init__()
        self.m, self.inputs = m, []

    def forward(self, x):
        self.inputs.append(x.detach().clone())
        return self.m(x)


def tiny_problem(C, B, H, Wd, seed, frac_ignore=0.0):
    model = TinySeg(C, seed).eval()
    g = torch.Generator().manual_seed(seed + 7)
    x = torch.rand(B, 3, H, Wd, generator=g)
    with torch.no_grad():
        y = model(x).argmax(1)
    flip = torch.rand(B, H, Wd, generator=g) < 0.2
    y = torch.where(flip, torch.randint(0, C, (B, H, Wd), generator=g), y)
    if frac_ignore:
        y = torch.where(torch.rand(B, H, Wd, generator=g) < frac_ignore, torch.full_like(y, -1), y)
    w = 0.5 + torch.rand(C, generator=g)
    return model, x, y, w


def model_arrays(model):
    return {"w_" + k.replace(".", "_"): v for k, v in model.m.state_dict().items()} if isinstance(
        model, Recorder) else {"w_" + k.replace(".", "_"): v for k, v in model.state_dict().items()}


def golden_apgd():
    for tag, (C, B, H, Wd, seed, loss, n_iter, eps, fi) in {
        "maskce": (5, 3, 12, 10, 11, "mask-ce-avg", 25, 8 / 255, 0.0),
        "maskbal_ign": (6, 2, 10, 12, 12, "mask-ce-bal", 20, 8 / 255, 0.1),
        "js": (5, 2, 12, 10, 13, "js-avg", 20, 4 / 255, 0.0),
    }.items():
        model, x, y, w = tiny_problem(C, B, H, Wd, seed, fi)
        rec = Recorder(model).eval()
        torch.manual_seed(1000 + seed)
        noise = []
        # replay the RNG draws the reference will make (one rand_like per stage, §9-Q10)
        st = torch.get_rng_state()
        for _ in range(3):
            noise.append(2 * torch.rand_like(x) - 1)
        torch.set_rng_state(st)
        x_adv, _, acc = A.apgd_largereps(
            rec, x.clone(), y, w, norm="Linf", eps=eps, n_iter=n_iter, loss=loss,
            track_loss="ce-avg", use_rs=True, early_stop=True, num_classes=C,
            log_path=None)
        save(f"apgd_{tag}", x=x, y=y, weights=w, eps=np.float64(eps), n_iter=n_iter,
             noise=torch.stack(noise), x_adv=x_adv, acc=acc, trace=torch.stack(rec.inputs),
             **model_arrays(rec), C=C, seed=seed)
    # a single apgd_train call, long enough for several step-size checks, with all outputs
    C, B, H, Wd, seed = 5, 4, 10, 10, 21
    model, x, y, w = tiny_problem(C, B, H, Wd, seed)
    rec = Recorder(model).eval()
    torch.manual_seed(77)
    st = torch.get_rng_state()
    noise = 2 * torch.rand_like(x) - 1
    torch.set_rng_state(st)
    logger = sys.modules["autoattack.other_utils"].Logger(None)
    x_best, acc, loss_best, x_best_adv = A.apgd_train(
        rec, x.clone(), y, "Linf", 8 / 255, n_iter=40, use_rs=True, loss="mask-ce-avg",
        track_loss="ce-avg", early_stop=False, logger=logger, num_classes=C, weights=w)
    save("apgd_train40", x=x, y=y, weights=w, eps=np.float64(8 / 255), n_iter=40, noise=noise,
         x_best=x_best, acc=acc, loss_best=loss_best, x_best_adv=x_best_adv,
         trace=torch.stack(rec.inputs), **model_arrays(rec), C=C, seed=seed)


# ------------------------------------------------------------------ (iii) metrics
def golden_metrics():
    g = torch.Generator().manual_seed(5)
    C, B, H, Wd = 9, 3, 17, 13
    target = torch.randint(0, C - 1, (B, H, Wd), generator=g)  # class C-1 never a target
    pred = torch.where(torch.rand(B, H, Wd, generator=g) < 0.6, target,
                       torch.randint(0, C, (B, H, Wd), generator=g))
    target = torch.where(torch.rand(B, H, Wd, generator=g) < 0.1, torch.full_like(target, -1), target)
    m_acc, a_acc, m_iou = A.compute_iou_acc(pred.clone(), target, C)
    logits = torch.randn(B, C, H, Wd, generator=g)
    met = M.Metrics(C, -1, "cpu")
    met.update(logits, target)
    hist1 = met.hist.clone()
    met.update(torch.nn.functional.one_hot(pred, C).permute(0, 3, 1, 2).float(), target)
    ious, miou = met.compute_iou()
    met2 = M.Metrics(C, -1, "cpu"); met2.hist = met.hist.clone()  # finalisers mutate in place
    hist = met.hist.clone()
    met = M.Metrics(C, -1, "cpu"); met.hist = hist.clone()
    ious, miou = met.compute_iou()
    met = M.Metrics(C, -1, "cpu"); met.hist = hist.clone()
    f1, mf1 = met.compute_f1()
    met = M.Metrics(C, -1, "cpu"); met.hist = hist.clone()
    acc, macc, aacc = met.compute_pixel_acc()
    save("metrics", pred=pred, target=target, C=C, m_acc=m_acc, a_acc=a_acc, m_iou=m_iou,
         logits=logits, hist_after_logits=hist1, hist=hist, ious=np.array(ious), miou=miou,
         f1=np.array(f1), mf1=mf1, acc=np.array(acc), macc=macc, aacc=np.float32(aacc))


# ------------------------------------------------------------------ (iv) evalSEA
class FakeVal(torch.utils.data.Dataset):
    def __init__(self, targets):
        self.t = targets

    def __len__(self):
        return self.t.shape[0]

    def __getitem__(self, i):
        return torch.zeros(1), self.t[i], str(i)


def golden_sea():
    g = torch.Generator().manual_seed(9)
    C, N, H, Wd, NA = 6, 8, 12, 11, 3
    target = torch.randint(0, C, (N, H, Wd), generator=g)
    l_outs = []
    for a in range(NA):
        keep = torch.rand(N, 1, 1, generator=g) * 0.8 + 0.1  # per-image accuracy differs per attack
        p = torch.where(torch.rand(N, H, Wd, generator=g) < keep, target,
                        torch.randint(0, C, (N, H, Wd), generator=g))
        l_outs.append(p)
    tmp = tempfile.mkdtemp()
    os.makedirs(os.path.join(tmp, "test_results"))
    sd = {}
    ev = W.evalSEA(FakeVal(target), [t.clone() for t in l_outs], 8, C, "x", tmp, sd, "m")
    # single-process loaders: the fixture does not depend on worker processes
    ev.get_loader = lambda bs=1: torch.utils.data.DataLoader(ev.val_data, batch_size=bs, shuffle=False)
    ev.worse_case_eval(bs=4)
    random.seed(225)
    ev.worst_case_miou()
    stats = torch.load(os.path.join(tmp, "test_results", "stats_x_8.pt"))
    save("sea", target=target, l_outs=torch.stack(l_outs), C=C, bs=4,
         worst_Acc=np.float64(sd["worst_Acc"]), worst_Acc_indiv=sd["worst_Acc_indiv"],
         final_miou=np.float64(sd["final_miou"]),
         cons_ints=stats["run_int_imwise"], cons_unions=stats["run_union_imwise"])


# ------------------------------------------------------------------ (v) PIR-AT PGD
def golden_pgd():
    torch.Tensor.cuda = lambda self, *a, **k: self  # reference hard-codes .cuda() (val.py:141,143,192)
    for tag, (cls, los, kw) in {
        "pgd1_pgd": (V.Pgd_Attack_1, "pgd", dict(epsilon=4 / 255)),
        "pgd_maskce": (V.Pgd_Attack, "mask-ce-avg", dict(eps=4 / 255)),
        "pgd_js": (V.Pgd_Attack, "js-avg", dict(eps=4 / 255)),
    }.items():
        C, B, H, Wd, seed = 5, 3, 10, 12, 31
        model, x, y, w = tiny_problem(C, B, H, Wd, seed)
        rec = Recorder(model).eval()
        atk = cls(alpha=1e-2, num_iter=3, los=los, **kw)
        torch.manual_seed(5)
        st = torch.get_rng_state()
        d0 = torch.zeros_like(x).uniform_(-4 / 255, 4 / 255)
        torch.set_rng_state(st)
        x_adv = atk.adv_attack(rec, x.clone(), y)[0]
        save(tag, x=x, y=y, eps=np.float64(4 / 255), alpha=np.float64(1e-2), num_iter=3,
             delta0=d0, x_adv=x_adv, trace=torch.stack(rec.inputs), **model_arrays(rec), C=C, seed=seed)


if __name__ == "__main__":
    golden_losses()
    golden_apgd()
    golden_metrics()
    golden_sea()
    golden_pgd()
