"""Pin the CPU oracle (oracle/robseg_oracle.py) against the golden vectors the reference
itself produced (tests/golden/make_golden.py).  CPU only."""
import random

import numpy as np
import pytest
import torch

import robseg_oracle as O

KINDS = ["mask-ce-avg", "mask-ce-bal", "js-avg", "ce-avg"]


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.mark.parametrize("tag", ["c7", "c21", "c151"])
@pytest.mark.parametrize("kind", KINDS)
def test_loss_and_grad_match_reference(golden, tag, kind):
    g = golden("loss_" + tag)
    z, y, w = g["logits"], g["labels"], g["weights"]
    B, C = z.shape[:2]
    k = kind.replace("-", "_")
    o = O.loss_fwd_bwd(z.reshape(B, C, -1), y.reshape(B, -1), kind, w)
    assert _rel(o["loss_pix"].reshape(y.shape), g[k + "__loss_pix"]) < 2e-6
    assert _rel(o["loss_img"], g[k + "__loss_img"]) < 2e-6
    assert _rel(o["dlogits"].reshape(z.shape), g[k + "__dlogits"]) < 2e-6
    d_up = O.loss_pixel_bwd(z.reshape(B, C, -1), y.reshape(B, -1), kind,
                            g[k + "__upstream"].reshape(B, -1), w)
    assert _rel(d_up.reshape(z.shape), g[k + "__dlogits_up"]) < 2e-6
    # integer side: argmax with ties -> lowest index; both accuracy conventions
    assert np.array_equal(o["pred"].reshape(y.shape), g["pred"])
    P = y[0].size
    assert np.array_equal((o["correct"] / np.float32(P)).astype(np.float32), g["acc_step0"])
    loop = ((o["correct"] + (P - o["valid"])) / np.float32(P)).astype(np.float32)
    assert np.array_equal(loop, g["acc_loop"])


def test_track_loss_is_plain_ce(golden):
    g = golden("loss_c7")
    z, y = g["logits"], g["labels"]
    B, C = z.shape[:2]
    o = O.loss_fwd_bwd(z.reshape(B, C, -1), y.reshape(B, -1), "mask-ce-avg", g["weights"])
    assert _rel(o["track_img"], g["ce_avg__loss_img"]) < 2e-6


class _Tiny(torch.nn.Module):
    def __init__(self, g):
        super().__init__()
        C = int(g["C"])
        self.c1 = torch.nn.Conv2d(3, 8, 3, padding=1)
        self.c2 = torch.nn.Conv2d(8, C, 3, padding=1)
        sd = {k[2:].replace("_", ".", 1): torch.from_numpy(v) for k, v in g.items() if k.startswith("w_")}
        self.load_state_dict(sd)

    def forward(self, x):
        return self.c2(torch.tanh(self.c1(x - 0.5)))


def tiny_model(g):
    return _Tiny(g).eval()


def _trace_agreement(tr_o, tr_g, eps):
    """Fraction of elements whose value differs; sign(grad) can only flip where the
    reference's autograd chain and the closed form round differently (|grad| ~ 0)."""
    n = min(len(tr_o), len(tr_g))
    bad = [(np.abs(tr_o[i] - tr_g[i]) > 1e-6).mean() for i in range(n)]
    return n, bad


@pytest.mark.parametrize("tag,kind", [("maskce", "mask-ce-avg"), ("maskbal_ign", "mask-ce-bal"),
                                      ("js", "js-avg")])
def test_apgd_largereps_trajectory(golden, tag, kind):
    g = golden("apgd_" + tag)
    torch.set_num_threads(1)
    model = O.TorchModelAdapter(tiny_model(g))
    trace = []
    x_adv, loss_best, acc = O.apgd_largereps(
        model, g["x"], g["y"], g["weights"], eps=float(g["eps"]), n_iter=int(g["n_iter"]),
        loss=kind, early_stop=True, use_rs=True, rand_ts=list(g["noise"]), trace=trace)
    assert len(trace) == len(g["trace"])
    n, bad = _trace_agreement(trace, g["trace"], float(g["eps"]))
    # the first model input of every stage and the first update must be exact
    assert bad[0] == 0.0 and bad[1] <= 0.01
    assert max(bad) <= 0.05, bad
    assert np.abs(acc - g["acc"]).max() <= 1.5 / g["y"][0].size
    assert (np.abs(x_adv - g["x"]) <= float(g["eps"]) + 1e-6).all()


def test_apgd_train_outputs(golden):
    g = golden("apgd_train40")
    torch.set_num_threads(1)
    model = O.TorchModelAdapter(tiny_model(g))
    trace = []
    x_best, acc, loss_best, x_best_adv = O.apgd_train(
        model, g["x"], g["y"], float(g["eps"]), n_iter=int(g["n_iter"]), use_rs=True,
        loss="mask-ce-avg", early_stop=False, weights=g["weights"], rand_t=g["noise"], trace=trace)
    assert len(trace) == len(g["trace"]) == 41
    n, bad = _trace_agreement(trace, g["trace"], float(g["eps"]))
    assert bad[0] == 0.0 and max(bad) <= 0.05, bad
    assert _rel(loss_best, g["loss_best"]) < 1e-3
    assert np.abs(acc - g["acc"]).max() <= 1.5 / g["y"][0].size
    assert (np.abs(x_best - g["x_best"]) > 1e-6).mean() <= 0.05
    assert (np.abs(x_best_adv - g["x_best_adv"]) > 1e-6).mean() <= 0.05


def test_apgd_step_teacher_forced(golden):
    """Given the reference's own consecutive model inputs, one oracle step reproduces the
    next input bit-exactly wherever the reference did not restart a row."""
    g = golden("apgd_train40")
    model = O.TorchModelAdapter(tiny_model(g))
    tr, x, y = g["trace"], g["x"], g["y"]
    B = x.shape[0]
    yf = y.reshape(B, -1)
    eps = float(g["eps"])
    step = np.float32(2 * eps) * np.ones(B, np.float32)
    exact = 0
    for i in range(0, 6):  # before the first step-size check (k = 8)
        logits = model.forward(tr[i])
        o = O.loss_fwd_bwd(logits, yf, "mask-ce-avg", g["weights"])
        grad = model.vjp(o["dlogits"])
        x_old = tr[i - 1] if i > 0 else tr[0]
        nxt = O.apgd_step(x, tr[i], x_old, grad, step, eps, 0.75 if i > 0 else 1.0)
        frac_bad = (nxt != tr[i + 1]).mean()
        assert frac_bad <= 0.01, (i, frac_bad)
        exact += frac_bad == 0.0
    assert exact >= 3


def test_schedule_matches_reference_constants():
    # semseg/attacker.py:322-329 for n_iter = 300*0.4 = 120: k starts at 26, shrinks by 3, floor 7
    checks = O.apgd_schedule(120)
    assert checks[0] == (25, 26) and checks[1] == (48, 23)
    assert all(k >= 7 for _, k in checks)
    assert O.apgd_schedule(3) == [(0, 1), (1, 1), (2, 1)]


def test_metrics_match_reference(golden):
    g = golden("metrics")
    C = int(g["C"])
    h = O.pixel_hist(g["pred"], g["target"], C)
    m_acc, a_acc, m_iou = O.iou_acc_from_counts(h["inter"].sum(0), h["tgt"].sum(0), h["prd"].sum(0))
    assert np.float32(g["m_acc"]) == m_acc and np.float32(g["a_acc"]) == a_acc
    assert np.float32(g["m_iou"]) == m_iou
    pred_logits = g["logits"].argmax(1)
    h1 = O.pixel_hist(pred_logits, g["target"], C)["hist"].sum(0)
    assert np.array_equal(h1.astype(np.float32), g["hist_after_logits"])
    hist = h1 + h["hist"].sum(0)
    assert np.array_equal(hist.astype(np.float32), g["hist"])
    f = O.metrics_from_hist(hist)
    np.testing.assert_allclose(f["ious"], g["ious"], rtol=0, atol=0.011, equal_nan=True)
    np.testing.assert_allclose(f["f1"], g["f1"], rtol=0, atol=0.011, equal_nan=True)
    np.testing.assert_allclose(f["acc"], g["acc"], rtol=0, atol=0.011, equal_nan=True)
    assert abs(f["miou"] - float(g["miou"])) <= 0.011
    assert abs(f["mf1"] - float(g["mf1"])) <= 0.011
    assert abs(f["macc"] - float(g["macc"])) <= 0.011
    assert abs(f["aacc"] - float(g["aacc"])) <= 0.011


def test_sea_aggregation_matches_reference(golden):
    g = golden("sea")
    C = int(g["C"])
    A, N = g["l_outs"].shape[:2]
    cnt = O.pixel_hist(g["l_outs"].reshape(A * N, -1), np.tile(g["target"].reshape(N, -1), (A, 1)), C)
    inter = cnt["inter"].reshape(A, N, C)
    tgt = cnt["tgt"].reshape(A, N, C)
    union = (cnt["tgt"] + cnt["prd"] - cnt["inter"]).reshape(A, N, C)
    assert np.array_equal(inter.astype(np.float32), g["cons_ints"])
    assert np.array_equal(union.astype(np.float32), g["cons_unions"])
    acc_an = O.sea_image_acc(inter, tgt)
    worst, per_attack = O.sea_worst_acc(acc_an)
    assert abs(worst - float(g["worst_Acc"])) <= 1e-7
    np.testing.assert_allclose(per_attack, g["worst_Acc_indiv"], rtol=0, atol=1e-7)
    random.seed(225)
    final, sel = O.sea_worst_miou(inter, union)
    assert final == float(g["final_miou"])  # bit-exact python double


def test_config1_full_size_aggregation_matches_reference(golden):
    """BASELINE config 1 at full size (2 x 512^2, 21 classes): the reference's own tools/infer.py flow
    (tests/golden/make_golden_config1.py) produced these argmax maps, per-attack statistics, worst-case
    aACC and worst-case mIoU; the oracle's counters / finalisers / greedy must reproduce them exactly."""
    g = golden("config1_sea")
    C = 21
    l_outs, y = g["l_outs"].astype(np.int64), g["y"].astype(np.int64)
    A, N = l_outs.shape[:2]
    cnt = O.pixel_hist(l_outs.reshape(A * N, -1), np.tile(y.reshape(N, -1), (A, 1)), C)
    inter, tgt, prd = (cnt[k].reshape(A, N, C) for k in ("inter", "tgt", "prd"))
    for a, k in enumerate(["mask-ce-bal", "mask-ce-avg", "js-avg"]):
        m_acc, a_acc, m_iou = O.iou_acc_from_counts(inter[a].sum(0), tgt[a].sum(0), prd[a].sum(0))
        # the class means are float32 sums in torch's reduction order; the oracle rounds the exact
        # mean once -> equal to within one float32 ulp; the ratio of totals (aAcc) is exact
        np.testing.assert_allclose([float(m_acc), float(a_acc), float(m_iou)], g["stats__" + k], rtol=1.5e-7)
        assert float(a_acc) == g["stats__" + k][1], k
        # per-image accuracy returned by the attack == accuracy of the re-forwarded argmax map
        np.testing.assert_allclose(inter[a].sum(-1) / tgt[a].sum(-1), g["acc__" + k], atol=2e-5)
    worst, per_attack = O.sea_worst_acc(O.sea_image_acc(inter, tgt))
    assert worst == float(g["worst_Acc"])
    assert np.array_equal(np.asarray(per_attack, dtype=np.float32), g["worst_Acc_indiv"])
    random.seed(225)
    final, _ = O.sea_worst_miou(inter, tgt + prd - inter)
    assert final == float(g["final_miou"])


@pytest.mark.parametrize("tag,kind,rs,clamp,best", [
    ("pgd1_pgd", "pgd", True, False, False),
    ("pgd_maskce", "mask-ce-avg", False, True, True),
    ("pgd_js", "js-avg", False, True, True)])
def test_pgd_attack_matches_reference(golden, tag, kind, rs, clamp, best):
    g = golden(tag)
    torch.set_num_threads(1)
    model = O.TorchModelAdapter(tiny_model(g))
    x_adv = O.pgd_attack(model, g["x"], g["y"], float(g["eps"]), float(g["alpha"]),
                         int(g["num_iter"]), loss=kind,
                         random_start_delta=g["delta0"] if rs else None,
                         clamp_input=clamp, track_best=best)
    assert (np.abs(x_adv - g["x_adv"]) > 1e-6).mean() <= 0.02
    assert (np.abs(x_adv - g["x"]) <= float(g["eps"]) + 1e-6).all()
