"""CPU tests of the host-side logic: the C-ABI library loads and exports every declared
symbol (no compute calls), the worst-case aggregation host code is bit-exact against the
reference's golden run and against the oracle, the data-independent APGD schedule, the
image-shard plumbing and its world-size-2 gloo all-reduce, and the no-CPU-fallback rule."""
import os
import random
import re
import statistics
import subprocess
import sys

import numpy as np
import pytest
import torch

import robseg_oracle as O

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module")
def mods(pkg):
    from importlib import import_module

    names = dict(lib="._lib", ops=".ops", attacker=".semseg.attacker", worse=".tools.worse_only",
                 dist=".dist", consumers=".consumers", val=".semseg.val", metrics=".semseg.metrics")
    return type("M", (), {k: import_module("robseg_b200" + v) for k, v in names.items()})


def test_library_exports_every_declared_symbol(mods):
    import __graft_entry__ as ge

    if not os.path.isfile(mods.lib.LIB_PATH):
        ge.build()
    header = open(os.path.join(ROOT, "include", "robseg_b200.h")).read()
    declared = set(re.findall(r"\b(robseg_[a-z0-9_]+)\s*\(", header))
    declared -= {"robseg_stream_t"}
    assert declared == set(mods.lib.SIGNATURES), declared ^ set(mods.lib.SIGNATURES)
    lib = mods.lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.robseg_version() == mods.lib.ABI_VERSION
    # per-tile partials (one float4 per 32 pixels, upper bound) + 8 copies of the [B,3,C] int64 class counters
    assert lib.robseg_loss_workspace_bytes(16, 150, 512 * 512, 0) == 16 * (512 * 512 // 32) * 16 + 8 * 16 * 3 * 150 * 8
    syms = subprocess.run(["nm", "-D", "--defined-only", mods.lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (robseg_[a-z0-9_]+)", syms))
    assert declared <= exported


def test_library_is_sm100a_with_tma(mods):
    sass = subprocess.run(["cuobjdump", "-sass", mods.lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "UTMALDG" in sass  # the loss kernel's tensor-map bulk loads


def test_no_cpu_fallback(mods):
    z = torch.randn(1, 5, 4, 4)
    y = torch.randint(0, 5, (1, 4, 4))
    with pytest.raises(RuntimeError):
        mods.ops.loss_fwd_bwd(z, y, "ce")
    with pytest.raises(RuntimeError):
        mods.attacker.apgd_train(mods.consumers.TinySegNet(5).eval(), torch.rand(1, 3, 4, 4), y, "Linf", 0.03)
    with pytest.raises(NotImplementedError):
        mods.attacker.apgd_train(mods.consumers.TinySegNet(5).eval(), torch.rand(1, 3, 4, 4), y, "L2", 0.03)
    # the product never imports the oracle
    for dirpath, _, files in os.walk(os.path.join(ROOT, "robust-segmentation_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "robseg_oracle" not in src and "oracle" not in src.replace("oracle/", ""), f


def test_interpolate_dispatcher_on_cpu_is_the_stock_function(mods):
    """ops.interpolate only takes CUDA fp32 bilinear/align_corners=False calls; on the CPU (and for
    every other mode) it must be exactly torch's function, and the scoped rebinding must undo itself."""
    F = torch.nn.functional
    stock = F.interpolate
    x = torch.randn(1, 2, 4, 5)
    with mods.ops.patched_interpolate():
        assert F.interpolate is mods.ops.interpolate
        a = F.interpolate(x, size=(8, 10), mode="bilinear", align_corners=False)
        b = F.interpolate(x, scale_factor=2, mode="nearest")
    assert F.interpolate is stock
    assert torch.equal(a, stock(x, size=(8, 10), mode="bilinear", align_corners=False))
    assert torch.equal(b, stock(x, scale_factor=2, mode="nearest"))
    with pytest.raises(RuntimeError):  # the kernel wrapper itself refuses CPU tensors
        mods.ops.upsample_bilinear(x, (8, 10))


def test_consumers_cpu_shapes_and_fast_flag_is_inert_on_cpu(mods):
    torch.manual_seed(0)
    m = mods.consumers.upernet_convnext("T", 7).eval()
    x = torch.rand(1, 3, 64, 64)
    with torch.no_grad():
        a = m(x)
        m.fast_upsample = "all"  # CPU input: every up-sampling stays F.interpolate
        b = m(x)
    assert a.shape == (1, 7, 64, 64) and torch.equal(a, b)
    s = mods.consumers.segmenter_vit("S", 7, 32, fast_upsample=True).eval()
    with torch.no_grad():
        assert s(torch.rand(2, 3, 32, 32)).shape == (2, 7, 32, 32)
    with pytest.raises(ValueError):
        s(torch.rand(1, 3, 48, 48))


def test_upsample_walk_kernel_ownership_bound():
    """upsample_bwd_walk_kernel assigns every output to the cell of its first tap and keeps <= RMAX =
    ceil(W/w)+1 outputs per cell and row in registers; the walk also assumes the owner never skips
    a cell.  Both follow from ATen's source-index formula for any ratio >= 1: checked here in float32
    (two-step rounding and the fused variant the device may use)."""
    f32 = np.float32
    for w in list(range(1, 40)) + [59, 119, 128, 237]:
        for W in range(w, min(10 * w, 1200) + 1):
            sx = f32(w) / f32(W)
            X = np.arange(W, dtype=np.float32)
            srcs = ((sx * (X + f32(0.5)) - f32(0.5)).astype(np.float32),
                    (np.float64(sx) * (X.astype(np.float64) + 0.5) - 0.5).astype(np.float32))
            for src in srcs:
                own = np.where(src < 0, -1, np.minimum(src.astype(np.int64), w - 1))
                d = np.diff(own)
                assert d.min(initial=0) >= 0 and d.max(initial=0) <= 1, (w, W)
                assert np.bincount(own + 1, minlength=w + 1).max() <= -(-W // w) + 1, (w, W)
                assert own[-1] == w - 1


def test_bench_helpers():
    import bench

    assert bench.stage_iters(10) == [3, 3, 4] and bench.stage_iters(300) == [90, 90, 120]
    ns = lambda **k: type("A", (), {**dict(stock_upsample=False, logit_upsample_only=False), **k})()
    assert bench.upsample_mode(ns()) == "all"
    assert bench.upsample_mode(ns(logit_upsample_only=True)) is True
    assert bench.upsample_mode(ns(stock_upsample=True)) is False


def test_bench_roofline_assembly_and_traffic_provenance(monkeypatch):
    """bench.sea_roofline: `achieved` over the loss kernel's OWN launches (the brackets the library records,
    robseg_profile_next_kernel), the call-level brackets beside it; bench.load_traffic: the committed ncu capture is
    passed on only for the kernel sources it was taken on."""
    import bench

    nb = 5066719232
    by = {"loss_grad": [54 * nb, 54 * 0.8715, 54], "loss_grad_counts": [36 * nb, 36 * 0.881, 36], "apgd_step": [1, 1.0, 30]}
    byk = {"loss_grad": [54 * nb, 54 * 0.8566, 54], "loss_grad_counts": [36 * nb, 36 * 0.8595, 36]}
    r = bench.sea_roofline(by, byk, 16, 150, 512, (6454.6, "measured"), {"sm_mhz": 1965.0})
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["launches_timed"] == 90
    assert abs(r["avg_launch_ms"] - (54 * 0.8566 + 36 * 0.8595) / 90) < 1e-4
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-4 and 0.91 < r["frac"] < 0.92
    assert r["algorithmic_bytes_per_launch"] == nb
    assert r["call_bracket"]["launches_timed"] == 54 and r["call_bracket_with_class_counters"]["launches_timed"] == 36
    assert r["call_bracket"]["frac"] < r["kernel_uncounted"]["frac"]
    # no kernel-level records (e.g. an older library): falls back to the call brackets and says so
    r0 = bench.sea_roofline(by, {}, 16, 150, 512, (6454.6, "measured"), None)
    assert r0["launches_timed"] == 90 and r0["timed_with"] == "events around the C call" and r0["frac"] < r["frac"]
    # fused up-sampling route (SegMenter): the ex2-pipe fraction is reported beside the (by design small) HBM one
    up = {"loss_up_grad": [13303808 * 10, 10 * 0.18, 10]}
    ru = bench.sea_roofline(up, up, 4, 150, 512, (6454.6, "measured"), {"sm_mhz": 1965.0})
    assert "ex2_pipe" in ru and ru["launches_timed"] == 10 and 0 < ru["ex2_pipe"]["frac"] < 1
    # traffic: current capture -> value; other sources -> null + the stale value
    v, cap = bench.load_traffic("sea_c150")
    assert cap["kernel_source_sha"] == bench.kernel_source_sha() and cap["matches_tree"] and v > 4.9e9
    assert cap["report"].endswith(".ncu-rep")
    monkeypatch.setattr(bench, "kernel_source_sha", lambda: "0" * 16)
    v2, cap2 = bench.load_traffic("sea_c150")
    assert v2 is None and cap2["matches_tree"] is False and cap2["stale_value"] == v
    assert bench.load_traffic("no_such_key") == (None, None)


def test_graph_capture_keeps_the_garbage_collector_out(pkg, monkeypatch):
    """graphs._capturing: garbage is collected BEFORE the capture starts and the cyclic collector stays off until it has
    ended (a CUDAGraph destroyed mid-capture releases its memory pool, which invalidates the capture), thread-local
    capture mode, and the collector's state is restored whatever happens inside."""
    import contextlib
    import gc
    from importlib import import_module

    import torch

    graphs = import_module("robseg_b200.graphs")
    seen = {}

    class Cycle:
        def __init__(self):
            self.me = self

        def __del__(self):
            seen["collected_before_capture"] = "inside" not in seen

    @contextlib.contextmanager
    def fake_graph(graph, pool=None, capture_error_mode="global"):
        seen["mode"], seen["pool"] = capture_error_mode, pool
        seen["inside"] = gc.isenabled()
        yield

    monkeypatch.setattr(torch.cuda, "graph", fake_graph)
    Cycle()  # garbage in a reference cycle, like a dropped GraphedModel
    assert gc.isenabled()
    with graphs._capturing(object(), pool="P"):
        assert not gc.isenabled()
    assert gc.isenabled()
    assert seen == {"collected_before_capture": True, "mode": "thread_local", "pool": "P", "inside": False}
    with pytest.raises(ZeroDivisionError):
        with graphs._capturing(object()):
            1 / 0
    assert gc.isenabled()
    gc.disable()  # a caller that runs with the collector off keeps it off
    try:
        with graphs._capturing(object()):
            pass
        assert not gc.isenabled()
    finally:
        gc.enable()


def test_exact_mean_matches_statistics_mean(mods):
    rng = np.random.default_rng(0)
    for _ in range(300):
        n = int(rng.integers(1, 200))
        v = rng.random(n) * 10.0 ** rng.integers(-12, 3, n)
        if rng.random() < 0.3:
            v[rng.integers(0, n)] = 0.0
        assert mods.worse.exact_mean(v) == statistics.mean(v.tolist())


def test_exact_mean_hard_cases(mods):
    """Cancellation, mixed signs and magnitudes, halfway cases: the C++ super-accumulator must agree
    with statistics.mean (exact rational arithmetic) bit for bit."""
    rng = np.random.default_rng(1)
    cases = [[1e16, 1.0, -1e16], [0.1] * 10, [1.0, 1e-30, -1.0], [2.0 ** -1060, 2.0 ** -1061, 2.0 ** -1062],
             [1.7976931348623157e308, 1.7976931348623157e308, -1.7976931348623157e308], [0.0, -0.0, 0.0],
             [1.0 + 2.0 ** -52, 1.0, 1.0], [3.0, 3.0 + 2.0 ** -51], [-5.5, 2.25, 1e-300], [1 / 3, 1 / 7, 1 / 11] * 50]
    for _ in range(200):
        n = int(rng.integers(1, 300))
        v = rng.standard_normal(n) * 10.0 ** rng.integers(-20, 20, n)
        cases.append(v.tolist())
    for v in cases:
        assert mods.worse.exact_mean(v) == statistics.mean(v), v[:4]
    with pytest.raises(ValueError):
        mods.worse.exact_mean([])
    with pytest.raises(RuntimeError):
        mods.worse.exact_mean([1.0, float("inf")])


def test_greedy_worst_miou_vs_oracle_larger_problem(mods):
    rng = np.random.default_rng(11)
    A, N, C = 3, 60, 13
    tgt = rng.integers(1, 4000, (1, N, C))
    inter = np.minimum(rng.integers(0, 4000, (A, N, C)), tgt)
    union = tgt + rng.integers(0, 3000, (A, N, C))
    r1, r2 = random.Random(225), random.Random(225)
    f1, s1 = mods.worse.greedy_worst_miou(inter, union, rng=r1)
    f2, s2 = O.sea_worst_miou(inter, union, rng=r2)
    assert f1 == f2 and s1 == s2
    assert r1.random() == r2.random()  # same number of shuffles drawn


def test_greedy_worst_miou_bit_exact(mods, golden):
    g = golden("sea")
    random.seed(225)
    final, sel = mods.worse.greedy_worst_miou(g["cons_ints"], g["cons_unions"])
    assert final == float(g["final_miou"])
    # and against the oracle's pure-python replay on a different random problem
    rng = np.random.default_rng(3)
    A, N, C = 3, 12, 7
    tgt = rng.integers(1, 50, (1, N, C))
    inter = np.minimum(rng.integers(0, 50, (A, N, C)), tgt)
    union = tgt + rng.integers(0, 30, (A, N, C))
    r1, r2 = random.Random(5), random.Random(5)
    f1, s1 = mods.worse.greedy_worst_miou(inter, union, rng=r1)
    f2, s2 = O.sea_worst_miou(inter, union, rng=r2)
    assert f1 == f2 and s1 == s2


def test_schedule_is_data_independent(mods):
    for n in (1, 3, 4, 10, 90, 120, 300):
        assert mods.attacker.apgd_schedule(n) == dict(O.apgd_schedule(n))
    assert mods.attacker.apgd_schedule(120)[25] == 26


def test_shard_ranges_cover(mods):
    for n, w in ((2000, 8), (17, 4), (3, 8), (16, 1)):
        spans = [mods.dist.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_signature_parity_with_reference_surface(mods):
    """Names, positional order and defaults of SURVEY.md section 8b."""
    import inspect

    a = mods.attacker
    sig = inspect.signature(a.apgd_largereps)
    assert list(sig.parameters)[:4] == ["model", "x", "y", "weights"]
    assert sig.parameters["eps"].default == 8.0 / 255.0 and sig.parameters["n_iter"].default == 10
    assert sig.parameters["loss"].default == "ce" and sig.parameters["num_classes"].default == 21
    sig = inspect.signature(a.apgd_train)
    # the reference's 18 parameters in order; extensions (return_pred) may only follow them
    assert list(sig.parameters)[:18] == ["model", "x", "y", "norm", "eps", "n_iter", "use_rs", "loss", "verbose",
                                         "is_train", "early_stop", "track_loss", "logger", "y_target",
                                         "ignore_index", "x_init", "num_classes", "weights"]
    ext = list(sig.parameters)[18:]
    assert ext == ["return_pred", "return_counts", "gpuu"]
    assert all(sig.parameters[k].default is False for k in ext[:2])
    assert sig.parameters["gpuu"].default is None  # passed by tools/train_rob_seg.py:314, ignored
    assert set(a.criterion_dict) == {"ce", "ce-avg", "mask-ce-avg", "mask-ce-bal", "js-avg"}
    assert list(inspect.signature(a.compute_iou_acc).parameters) == [
        "pred", "target", "n_cls", "verbose", "ignore_index", "device"]
    p = mods.val.Pgd_Attack(epsilon=2 / 255)  # the trainer's spelling
    assert p.epsilon == 2 / 255 and mods.val.Pgd_Attack(eps=1 / 255).epsilon == 1 / 255
    assert mods.val.Pgd_Attack_1(epsilon=3 / 255).epsilon == 3 / 255
    m = inspect.signature(mods.metrics.Metrics.__init__)
    assert list(m.parameters)[1:] == ["num_classes", "ignore_label", "device"]
    sea = inspect.signature(mods.worse.evalSEA.__init__)
    assert list(sea.parameters)[1:9] == ["val_data", "l_outs", "eps", "n_cls", "addendum", "saveDir",
                                         "saveDict", "modelName"]


_GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
import __graft_entry__ as ge
ge.load_package()
from importlib import import_module
D = import_module("robseg_b200.dist")
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
A, N, C = 3, 11, 5
g = torch.Generator().manual_seed(0)
inter = torch.randint(0, 100, (A, N, C), generator=g)
tgt = inter + torch.randint(0, 50, (A, N, C), generator=g)
prd = inter + torch.randint(0, 50, (A, N, C), generator=g)
hist = torch.randint(0, 9, (A, C, C), generator=g)
lo, hi = D.shard_range(N, rank, world)
gi, gt, gp, gh = D.allreduce_counters(N, lo, inter[:, lo:hi], tgt[:, lo:hi], prd[:, lo:hi], hist)
assert torch.equal(gi, inter) and torch.equal(gt, tgt) and torch.equal(gp, prd)
assert torch.equal(gh, hist * world)
dist.destroy_process_group()
open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ok%d" % rank), "w").write("ok")
"""


def test_allreduce_counters_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER.format(root=ROOT))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29631", str(script)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()  # (stdout of the ranks interleaves)


def test_dropin_rebinds_reference_names(mods, tmp_path, monkeypatch):
    """dropin.install() against a stand-in checkout: the names SURVEY 8b lists are rebound to the
    B200 implementations, everything else in the reference package is left alone."""
    from importlib import import_module

    root = tmp_path / "Robust-Segmentation"
    (root / "semseg").mkdir(parents=True)
    (root / "tools").mkdir()
    (root / "semseg" / "__init__.py").write_text("MODELS = 'theirs'\n")
    (root / "semseg" / "attacker.py").write_text("def apgd_largereps(*a, **k):\n    return 'reference'\n")
    (root / "semseg" / "val.py").write_text("class Pgd_Attack: pass\nclass Pgd_Attack_1: pass\n"
                                            "def evaluate(): return 'reference'\nKEEP = 1\n")
    (root / "semseg" / "metrics.py").write_text("class Metrics: pass\n")
    (root / "semseg" / "losses.py").write_text("class CrossEntropy: pass\ndef get_loss(): pass\n")
    (root / "tools" / "__init__.py").write_text("")
    (root / "tools" / "worse_only.py").write_text("class evalSEA: pass\n")
    for name in [n for n in sys.modules if n == "semseg" or n.startswith("semseg.") or n == "tools" or n.startswith("tools.")]:
        monkeypatch.delitem(sys.modules, name)
    monkeypatch.syspath_prepend(str(root))
    dropin = import_module("robseg_b200.dropin")
    dropin.install(str(root))
    import semseg
    import semseg.attacker as ref_attacker
    import semseg.val as ref_val
    import tools.worse_only as ref_sea

    assert semseg.MODELS == "theirs" and ref_val.KEEP == 1
    assert ref_attacker is mods.attacker and semseg.attacker is mods.attacker
    assert ref_val.Pgd_Attack is mods.val.Pgd_Attack and ref_val.evaluate is mods.val.evaluate
    assert import_module("semseg.metrics").Metrics is mods.metrics.Metrics
    assert ref_sea.evalSEA is mods.worse.evalSEA
    for name in [n for n in sys.modules if n == "semseg" or n.startswith("semseg.") or n == "tools" or n.startswith("tools.")]:
        monkeypatch.delitem(sys.modules, name)


def _purge_reference_modules(monkeypatch):
    for name in [n for n in sys.modules if n == "semseg" or n.startswith("semseg.") or n == "tools" or n.startswith("tools.")]:
        monkeypatch.delitem(sys.modules, name)


def test_dropin_on_the_real_reference_copy(mods, monkeypatch):
    """dropin.install / uninstall / run_infer_main against the unmodified copy of the reference under
    baseline/_ref (no compute: import-level behaviour only).  tools.infer's own evaluate /
    eval_performance / attacker / evalSEA are rebound and restored; run_infer_main really executes the
    file's __main__ block (argparse sees --help and exits)."""
    from importlib import import_module

    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "semseg")):
        pytest.skip("baseline/_ref not present")
    _purge_reference_modules(monkeypatch)
    monkeypatch.syspath_prepend(ref)
    dropin = import_module("robseg_b200.dropin")
    dropin.uninstall()
    dropin.shim_missing_deps()
    import tools.infer as TI

    theirs = {n: getattr(TI, n) for n in ("attacker", "evaluate", "eval_performance", "evalSEA", "check_imgs")}
    assert not theirs["attacker"].__name__.startswith("robseg_b200")
    try:
        dropin.install(ref, accelerate_models=True)
        infer = import_module("robseg_b200.tools.infer")
        assert TI.attacker is mods.attacker and sys.modules["semseg.attacker"] is mods.attacker
        assert TI.evaluate is infer.evaluate and TI.eval_performance is infer.eval_performance
        assert TI.evalSEA is mods.worse.evalSEA
        assert TI.UperNetForSemanticSegmentation.__wrapped__.__name__ == "UperNetForSemanticSegmentation"
        assert import_module("semseg.val").Pgd_Attack is mods.val.Pgd_Attack
        assert import_module("semseg.models").UperNetForSemanticSegmentation is TI.UperNetForSemanticSegmentation.__wrapped__
        with pytest.raises(SystemExit):  # the reference's argparse, reached through ITS main block
            dropin.run_infer_main(["--help"])
        # the PIR-AT trainer (tools/train_rob_seg.py:19,22,33) binds the B200 attack / losses at import time
        TR = import_module("tools.train_rob_seg")
        assert TR.Pgd_Attack is mods.val.Pgd_Attack and TR.evaluate is mods.val.evaluate
        assert TR.attacker is mods.attacker and TR.get_loss is import_module("robseg_b200.semseg.losses").get_loss
        # run_train_main reaches the reference's own Trainer (its __init__ reads cfg["TRAIN"] first) and puts the
        # overridden names back whatever happens
        theirs_ds, sentinel = TR.get_segmentation_dataset, object()
        with pytest.raises(KeyError):
            dropin.run_train_main({}, overrides={"get_segmentation_dataset": sentinel, "brand_new_name": 1})
        assert TR.get_segmentation_dataset is theirs_ds and not hasattr(TR, "brand_new_name")
    finally:
        dropin.uninstall()
    for n, v in theirs.items():
        assert getattr(TI, n) is v, n
    assert sys.modules["semseg.attacker"] is theirs["attacker"]
    with pytest.raises(RuntimeError):
        dropin.run_infer_main(["--help"])  # not installed any more
    with pytest.raises(RuntimeError):
        dropin.run_train_main({})
    _purge_reference_modules(monkeypatch)


def test_launch_train_installs_the_dropin_in_every_spawned_process(pkg):
    """dropin.launch_train: torch.multiprocessing.spawn starts FRESH interpreters, so the rebinding has to be installed
    inside each of them before the reference's Trainer is built (the reference's own launcher after install() would run
    the reference's attack in its children).  One spawned process on the CPU: it must get as far as the reference's
    Trainer.__init__ (which reads cfg["TRAIN"] first) with the B200 names bound -- the worker reports what it saw through
    the exception it dies with."""
    from importlib import import_module

    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "semseg")):
        pytest.skip("baseline/_ref not present")
    dropin = import_module("robseg_b200.dropin")
    import torch.multiprocessing as mp

    with pytest.raises(mp.ProcessRaisedException) as ei:
        dropin.launch_train({"probe": "dropin-names"}, ref, world_size=1)
    msg = str(ei.value)
    assert "KeyError" in msg and "TRAIN" in msg, msg[-800:]        # reached THEIR Trainer.__init__
    assert "tools/train_rob_seg.py" in msg and "run_train_main" in msg  # ... through the installed drop-in
    with pytest.raises(RuntimeError):
        dropin.launch_train({}, ref, world_size=0)


def test_run_sea_shards_the_loader_without_touching_foreign_batches(pkg):
    """a rank must only decode its own batches."""
    from importlib import import_module

    sea = import_module("robseg_b200.tools.sea")

    class Data(torch.utils.data.Dataset):
        def __init__(self):
            self.touched = []

        def __len__(self):
            return 10

        def __getitem__(self, i):
            self.touched.append(i)
            return torch.full((3, 2, 2), float(i)), torch.zeros(2, 2, dtype=torch.int64), f"n{i}"

    for rank, want_imgs, want_lo in ((0, list(range(0, 6)), 0), (1, list(range(6, 10)), 2)):
        d = Data()
        loader = torch.utils.data.DataLoader(d, batch_size=3, shuffle=False)
        sizes, lo_b, own = sea._own_batches(loader, rank, 2, -1)
        got = [int(v[0][k, 0, 0, 0]) for v in own for k in range(v[0].shape[0])]
        assert sizes == [3, 3, 3, 1] and lo_b == want_lo and got == want_imgs and sorted(d.touched) == want_imgs
    # n_batches cap and a plain list of batches
    d = Data()
    sizes, lo_b, own = sea._own_batches(torch.utils.data.DataLoader(d, batch_size=3), 0, 1, 2)
    assert sizes == [3, 3] and len(list(own)) == 2
    batches = [(torch.zeros(2, 3, 2, 2), torch.zeros(2, 2, 2)) for _ in range(5)]
    sizes, lo_b, own = sea._own_batches(batches, 1, 2, -1)
    assert sizes == [2] * 5 and lo_b == 3 and len(own) == 2


def test_margin_type_losses_follow_their_definitions(mods):
    """dlr_loss / dlr_loss_targeted / margin_loss (name parity with semseg/attacker.py:123-141,176-184):
    checked element by element against the published definitions written out with Python sorts."""
    A = mods.attacker
    g = torch.Generator().manual_seed(3)
    x = torch.randn(11, 7, generator=g, dtype=torch.float64)
    y = torch.randint(0, 7, (11,), generator=g)
    t = torch.randint(0, 7, (11,), generator=g)
    y[0] = int(x[0].argmax())  # label = top class: the runner-up enters the numerator
    d, dt = A.dlr_loss(x, y), A.dlr_loss_targeted(x, y, t)
    for i in range(11):
        row = x[i].tolist()
        s = sorted(row)
        other = s[-2] if row.index(s[-1]) == int(y[i]) else s[-1]
        assert abs(float(d[i]) - (-(row[y[i]] - other) / (s[-1] - s[-3] + 1e-12))) < 1e-12
        assert abs(float(dt[i]) - (-(row[y[i]] - row[t[i]]) / (s[-1] - 0.5 * (s[-3] + s[-4]) + 1e-12))) < 1e-12
    z = torch.randn(2, 5, 3, 4, generator=g)
    lab = torch.randint(0, 5, (2, 3, 4), generator=g)
    m = A.margin_loss(z, lab)
    assert m.shape == lab.shape
    for b, i, j in ((0, 0, 0), (1, 2, 3), (0, 1, 2)):
        col = z[b, :, i, j].tolist()
        k = int(lab[b, i, j])
        assert abs(float(m[b, i, j]) - (max(v for c, v in enumerate(col) if c != k) - col[k])) < 1e-5
    z.requires_grad_(True)
    A.margin_loss(z, lab).sum().backward()  # differentiable, like the reference's
    assert z.grad is not None and float(z.grad.abs().sum()) > 0
