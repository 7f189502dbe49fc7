"""SURVEY.md 8f rank 4 as specified: ONE CUDA graph per APGD iteration (step -> forward -> fused loss ->
backward -> bookkeeping) with the per-iteration scalars in a device control block
(graphs.GraphedAttack, robseg_apgd_step_ctl / robseg_apgd_bookkeep_ctl)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods(pkg):
    from importlib import import_module

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    names = dict(ops=".ops", lib="._lib", attacker=".semseg.attacker", consumers=".consumers", graphs=".graphs")
    return type("M", (), {k: import_module("robseg_b200" + v) for k, v in names.items()})


def test_ctl_kernels_equal_the_host_driven_ones(mods):
    """robseg_apgd_step_ctl / robseg_apgd_bookkeep_ctl, driven only by the device control block, reproduce
    robseg_apgd_step_fused / robseg_apgd_bookkeep called with host scalars -- bit for bit, over a
    40-iteration schedule with several step-size checks and restarts."""
    ops, att = mods.ops, mods.attacker
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(3)
    B, shp, HW, n_iter, eps = 6, (3, 12, 16), 12 * 16, 40, 8 / 255
    rnd = lambda *s: torch.rand(*s, device=dev, generator=g)  # noqa: E731
    x = rnd(B, *shp)
    checks = att.apgd_schedule(n_iter)

    grad0 = rnd(B, *shp) - 0.5

    def init():
        st = dict(x_adv=(x + 0.01).clamp(0, 1).clone(), grad=grad0.clone(), acc=torch.full((B,), 0.9, device=dev),
                  loss_best=torch.zeros(B, device=dev), loss_best_last=torch.zeros(B, device=dev),
                  reduced_last=torch.ones(B, device=dev), step=torch.full((B,), 2 * eps, device=dev),
                  loss_steps=torch.zeros(n_iter, B, device=dev), flags=torch.zeros(3, B, dtype=torch.int32, device=dev),
                  done=torch.zeros(1, dtype=torch.int32, device=dev))
        st["x_old"] = st["x_adv"].clone()
        for k in ("x_best", "x_best_adv"):
            st[k] = st["x_adv"].clone()
        st["grad_best"] = st["grad"].clone()
        return st

    # a fixed script of per-iteration "model outputs": gradient, track loss, correct counts
    gs = torch.Generator(device=dev).manual_seed(9)
    script = [(torch.rand(B, *shp, device=dev, generator=gs) - 0.5, torch.rand(B, device=dev, generator=gs),
               torch.randint(0, HW, (B,), device=dev, generator=gs, dtype=torch.int32)) for _ in range(n_iter)]
    valid = torch.full((B,), HW, dtype=torch.int32, device=dev)

    a_, b_ = init(), init()
    x_new = torch.empty_like(x)
    ctl = ops.set_ctl(ops.make_ctl(n_iter, dev), n_iter, eps, checks)
    for i, (grad_i, track_i, corr_i) in enumerate(script):
        # host-driven pair (buffers rotate)
        ops.apgd_step_fused(x, a_["x_adv"], a_["x_old"], a_["grad"], a_["step"], eps, 0.75 if i else 1.0, x_new,
                            a_["flags"], a_["x_best_adv"], a_["x_best"], a_["grad_best"])
        a_["x_old"], a_["x_adv"], x_new = a_["x_adv"], x_new, a_["x_old"]
        a_["grad"] = grad_i.clone()
        ops.apgd_bookkeep(corr_i, valid, track_i, a_["acc"], a_["loss_best"], a_["loss_best_last"], a_["reduced_last"],
                          a_["step"], a_["loss_steps"], i, checks.get(i, 0), HW, False, a_["flags"], a_["done"])
        # device-driven pair (in place)
        ops.apgd_step_ctl(x, b_["x_adv"], b_["x_old"], b_["grad"], b_["step"], ctl, b_["flags"], b_["x_best_adv"],
                          b_["x_best"], b_["grad_best"])
        b_["grad"].copy_(grad_i)
        ops.apgd_bookkeep_ctl(corr_i, valid, track_i, b_["acc"], b_["loss_best"], b_["loss_best_last"],
                              b_["reduced_last"], b_["step"], b_["loss_steps"], ctl, HW, False, b_["flags"], b_["done"])
        for k in ("x_adv", "x_old", "x_best", "x_best_adv", "grad_best", "acc", "loss_best", "step", "flags", "loss_steps"):
            assert torch.equal(a_[k], b_[k]), (i, k)
    assert int(ctl[0]) == n_iter and float(a_["step"].min()) < 2 * eps  # some rows were halved / restarted


@pytest.mark.parametrize("loss,n_iter,early", [("mask-ce-avg", 10, True), ("mask-ce-bal", 25, False), ("js-avg", 10, True)])
def test_graphed_iteration_attack_matches_eager(mods, loss, n_iter, early):
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda")
    C, S = 7, 32
    model = mods.consumers.TinySegNet(C, seed=4).to(dev).eval()
    g = torch.Generator().manual_seed(12)
    x = torch.rand(4, 3, S, S, generator=g).to(dev)
    with torch.no_grad():
        y = model(x).argmax(1)
    y[0, :3] = -1
    w = (0.5 + torch.rand(C, generator=g)).to(dev)
    gm = mods.graphs.GraphedModel(model, x)
    kw = dict(norm="Linf", eps=8 / 255, n_iter=n_iter, loss=loss, track_loss="ce-avg", use_rs=True, early_stop=early,
              num_classes=C, return_pred=True)
    n0 = mods.lib.launches
    torch.manual_seed(5)
    xa, la, aa, pa = mods.attacker.apgd_largereps(model, x, y, w, **kw)
    n_eager = mods.lib.launches - n0
    torch.manual_seed(5)
    xb, lb, ab, pb = mods.attacker.apgd_largereps(gm, x, y, w, **kw)
    assert len(gm._attacks) == 1 and (loss, "ce-avg") in next(iter(gm._attacks.values())).graphs  # one runner, 3 graphs
    # cuDNN may choose another data-gradient algorithm under capture: sign(grad) flips only where |grad| ~ 0
    assert float(((xa - xb).abs() > 1e-6).float().mean()) <= 0.02
    assert float((aa - ab).abs().max()) <= 3 / (S * S) + 1e-7
    np.testing.assert_allclose(la.cpu().numpy(), lb.cpu().numpy(), rtol=2e-2)
    assert float((pa != pb).float().mean()) <= 0.01
    torch.manual_seed(5)
    again = mods.attacker.apgd_largereps(gm, x, y, w, **kw)  # replays are deterministic
    for p, q in zip(again, (xb, lb, ab, pb)):
        assert torch.equal(p, q)
    assert all(p.requires_grad for p in model.parameters()) and n_eager > 0
