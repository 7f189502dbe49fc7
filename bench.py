"""bench.py -- SEA attack throughput on B200 (BASELINE.json metric) + kernel roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One *step* = the full Segmentation Ensemble Attack on one batch: the three SEA losses
(mask-ce-bal, mask-ce-avg, js-avg), each ``apgd_largereps`` with n_iter=10 (stages 3/3/4 at
2eps/1.5eps/eps), on UperNet-ConvNeXt-T_CVST (random init), 150 classes, 16 synthetic 512x512
images per GPU, eps=8/255 (BASELINE.json configs[1]), followed by the per-batch adversarial
bookkeeping: argmax of the adversarial points, exact int64 per-image counters, (N>1: the one
int64 all-reduce), worst-case aACC.  ``value`` = attacked image-iterations per second over all
ranks with the batch resident in HBM; ``e2e`` = the same through the public API with the batch
coming from pinned host memory and x_adv / acc going back to the host every step.
``roofline`` = the dominant robseg kernel of the step (loss_tma_kernel): algorithmic bytes per launch / its own average
launch duration, from CUDA events the library records immediately before and after that launch on its stream
(robseg_profile_next_kernel) for every loss+gradient launch of the timed steps; the wider brackets around the whole C
call are reported beside it, ``traffic`` is the DRAM read+write of an ncu capture tied to the kernel-source sha.

``--impl reference`` times the CPU implementation of the same path on the box's host cores:
the unmodified reference attacker when a copy travels under baseline/_ref, else the oracle port.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOSSES = ["mask-ce-bal", "mask-ce-avg", "js-avg"]
METRIC = "SEA attacked image-iterations/sec (UperNet-ConvNeXt-T, 150 cls, 512x512)"
UNIT = "image-iterations/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--classes", type=int, default=150)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--n-iter", type=int, default=10)
    ap.add_argument("--eps", type=float, default=8.0)
    ap.add_argument("--variant", default=None, help="ConvNeXt T|S (upernet) or ViT S|B|L (segmenter)")
    ap.add_argument("--model", default="upernet", choices=["upernet", "segmenter"],
                    help="consumer: configs[1] UperNet-ConvNeXt (default) or configs[2] Segmenter-ViT")
    ap.add_argument("--workload", default="sea", choices=["sea", "pirat"],
                    help="sea: configs[1]/[2] SEA step (default); pirat: configs[3] PIR-AT training step under DDP")
    ap.add_argument("--own-consumer", action="store_true",
                    help="drive consumers.py's look-alike networks instead of the reference's model classes "
                         "(default: the reference's classes from baseline/_ref when that copy travelled)")
    ap.add_argument("--no-ref-on-gpu", action="store_true",
                    help="skip the unmodified-reference-on-this-GPU comparator (config.reference_on_gpu, N=1 only)")
    ap.add_argument("--micro", action="store_true", help="config-5 kernel microbench instead of the SEA step")
    ap.add_argument("--micro-batch", type=int, default=64)
    ap.add_argument("--micro-dtype", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--reforward", action="store_true",
                    help="re-forward every adversarial batch for its argmax map like tools/infer.py "
                         "(default: the attack returns it, SURVEY 8f-2)")
    ap.add_argument("--pred-maps", action="store_true",
                    help="score the adversarial points from int64 argmax maps + robseg_pixel_hist (round-1 flow) "
                         "instead of the counters fused into the loss kernel (robseg_loss_fwd_bwd_counts)")
    ap.add_argument("--stock-upsample", action="store_true",
                    help="keep F.interpolate for every bilinear up-sampling of the consumer (default: robseg kernels)")
    ap.add_argument("--logit-upsample-only", action="store_true",
                    help="robseg kernels for the final logit up-sampling only, F.interpolate inside the decode head")
    ap.add_argument("--fuse-loss", dest="fuse_loss", action="store_true", default=None,
                    help="the model hands the attack its logits BEFORE the final bilinear up-sampling and the loss "
                         "kernel interpolates on the fly (robseg_loss_upsampled_fwd_bwd, SURVEY 8f-1)")
    ap.add_argument("--no-fuse-loss", dest="fuse_loss", action="store_false")
    ap.add_argument("--graph", action="store_true",
                    help="replay the consumer's forward / input-gradient backward as CUDA graphs (SURVEY 8f-4)")
    ap.add_argument("--debug-stack", type=int, default=0, help="dump python stacks to stderr every N seconds")
    args = ap.parse_args()
    if args.variant is None:
        args.variant = ("S" if args.workload == "pirat" else "T") if args.model == "upernet" else "S"
    if args.fuse_loss is None:  # dropin.accelerate's "auto": on for SegMenter (x16), opt-in for UperNet (x4)
        args.fuse_loss = args.model == "segmenter"
    return args


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        top = sorted(sm)[len(sm) // 2:] if sm else []  # samples under load = upper half
        return {"sm_mhz": statistics.median(top) if top else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------ ours
def make_batch(B, C, S, seed, device=None, pin=False):
    import torch

    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, 3, S, S, generator=g)
    y = torch.randint(0, C, (B, S, S), generator=g)
    if pin:
        return x.pin_memory(), y.pin_memory()
    return x.to(device), y.to(device)


_PINNED = {}


def sea_step(mods, model, x, y, w, args, world, e2e_host=None):
    """Full SEA on one batch + the per-batch bookkeeping.  Returns worst-case accuracy [B]."""
    import torch

    att, ops, dist_mod = mods["attacker"], mods["ops"], mods["dist"]
    if e2e_host is not None:  # host buffers -> device inside the timed region
        x = e2e_host[0].to(x.device, non_blocking=True)
        y = e2e_host[1].to(y.device, non_blocking=True)
    B, C = x.shape[0], args.classes
    preds, counts = [], []
    x_advs = []
    for loss in LOSSES:
        if args.reforward:  # the reference's flow: re-forward every adversarial batch (tools/infer.py:82-90)
            x_adv, _, acc = att.apgd_largereps(
                model, x, y, w, norm="Linf", eps=args.eps / 255.0, n_iter=args.n_iter, loss=loss,
                track_loss="ce-avg", use_rs=True, early_stop=True, num_classes=C)
            with torch.no_grad():
                out = model(x_adv)
            pred = ops.loss_fwd_bwd(out, y, "argmax", want_grad=False, want_pred=True, want_stats=False).pred
        elif args.pred_maps:  # the attack hands back the argmax map of its adversarial point
            x_adv, _, acc, pred = att.apgd_largereps(
                model, x, y, w, norm="Linf", eps=args.eps / 255.0, n_iter=args.n_iter, loss=loss,
                track_loss="ce-avg", use_rs=True, early_stop=True, num_classes=C, return_pred=True)
        else:  # ... or directly the per-image class counters of that point, taken in the loss kernel's argmax pass
            x_adv, _, acc, cnt = att.apgd_largereps(
                model, x, y, w, norm="Linf", eps=args.eps / 255.0, n_iter=args.n_iter, loss=loss,
                track_loss="ce-avg", use_rs=True, early_stop=True, num_classes=C, return_counts=True)
            counts.append(cnt)
            pred = None
        preds.append(pred)
        x_advs.append(x_adv)
    if counts:
        inter, tgt, prd = (torch.stack(counts)[:, :, k].contiguous() for k in range(3))  # each [A,B,C]
    else:
        cnt = ops.pixel_hist(torch.stack(preds).flatten(0, 1), y, C)
        inter, tgt, prd = (cnt[k].view(len(LOSSES), B, C) for k in ("inter", "tgt", "prd"))
    rank = int(os.environ.get("RANK", 0))
    gi, gt, gp, _ = dist_mod.allreduce_counters(B * world, rank * B, inter, tgt, prd)
    acc_an, worst = ops.sea_worst_acc(gi, gt)
    if e2e_host is not None:  # results back to pinned host memory (tools/infer.py:151 keeps every x_adv)
        if "out" not in _PINNED:  # (run_ours allocates these before the timed region: pinning 150 MB takes ~0.1 s)
            _PINNED["out"] = [torch.empty(xa.shape, dtype=xa.dtype).pin_memory() for xa in x_advs]
            _PINNED["worst"] = torch.empty(worst.shape, dtype=worst.dtype).pin_memory()
        for h, xa in zip(_PINNED["out"], x_advs):
            h.copy_(xa, non_blocking=True)
        _PINNED["worst"].copy_(worst, non_blocking=True)
        return worst, (gi, gt, gp), _PINNED
    return worst, (gi, gt, gp), None


REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def import_reference(mods):
    """Make the copy of the reference under baseline/_ref importable (timm / autoattack stubs from
    dropin.shim_missing_deps) WITHOUT installing the drop-in: `semseg.*` stays the reference's."""
    if not os.path.isdir(os.path.join(REF_DIR, "semseg")):
        return False
    mods["dropin"].shim_missing_deps()
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    cwd = os.getcwd()
    os.chdir(REF_DIR)
    try:
        import semseg  # noqa: F401
    finally:
        os.chdir(cwd)
    return True


def build_consumer(args, mods, dev, accelerated=True):
    """The network around the hot path.  Default: the REFERENCE's own model class (unmodified copy under
    baseline/_ref), random init, with its bilinear up-samplings routed to the robseg kernels by
    dropin.accelerate.  Fallback / --own-consumer: consumers.py."""
    import torch

    mode = upsample_mode(args) if accelerated else False
    if not args.own_consumer:
        try:
            if import_reference(mods):
                dropin = mods["dropin"]
                torch.manual_seed(0)
                model = dropin.reference_model(args.model, args.variant, args.classes, args.size).to(dev).eval()
                fuse = bool(getattr(args, "fuse_loss", False)) and bool(mode)
                if mode == "all":
                    dropin.accelerate(model, head=True, fuse_loss=fuse)
                elif mode:
                    (dropin.fast_logit_upsample if args.model == "upernet" else dropin.fast_interpolate)(
                        model, fuse_loss=fuse)
                return model, ("reference class semseg.models.%s (unmodified copy under baseline/_ref), random init"
                               % type(model).__name__)
        except Exception as e:
            print(f"[bench] reference model unusable ({e!r}); using consumers.py", file=sys.stderr)
    torch.manual_seed(0)
    if args.model == "segmenter":
        model = mods["consumers"].segmenter_vit(args.variant, args.classes, args.size, fast_upsample=bool(mode))
    else:
        model = mods["consumers"].upernet_convnext(args.variant, args.classes, fast_upsample=mode)
    return model.to(dev).eval(), "consumers.py look-alike (%s), random init" % type(model).__name__


def variant_step(args, mods, dev, x, y, w, world, iters_per_step, fuse=False, graph=False):
    """The same step in another configuration, one warm-up (two with --graph: capture) and one timed step, reported
    beside the default so the decision rests on driver-visible numbers: `fuse` = the x4 logit up-sampling fused into
    the loss kernel (--fuse-loss), `graph` = one CUDA graph per APGD iteration (--graph)."""
    import torch

    try:
        saved = args.fuse_loss
        args.fuse_loss = bool(fuse)
        model, _ = build_consumer(args, mods, dev)
        args.fuse_loss = saved
        for p in model.parameters():
            p.requires_grad_(True)
        if fuse and not hasattr(model, "forward_lowres"):
            return {"unavailable": "consumer offers no forward_lowres"}
        if graph:
            model = mods["graphs"].GraphedModel(model, x)
        for _ in range(2 if graph else 1):
            torch.manual_seed(1234)
            sea_step(mods, model, x, y, w, args, world)
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats(dev)
        launches0 = mods["lib"].launches
        if not graph:  # (launches inside a replayed graph are not seen by the wrappers' events)
            mods["ops"].profile_start()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.manual_seed(1234)
        t0.record()
        sea_step(mods, model, x, y, w, args, world)
        t1.record()
        torch.cuda.synchronize()
        prof = mods["ops"].profile_stop() if not graph else []
        ms = t0.elapsed_time(t1)
        by = {}
        for name, _, t in prof:
            by[name] = by.get(name, 0.0) + t
        res = {"value": round(iters_per_step / (ms / 1e3), 3), "unit": UNIT, "ms_per_step": round(ms, 1),
               "peak_mem_GiB": round(torch.cuda.max_memory_allocated(dev) / 2**30, 1), "steps": 1,
               "warmup": 2 if graph else 1, "host_launch_calls": mods["lib"].launches - launches0}
        if by:
            res["attack_side_ms_per_step"] = round(sum(by.values()), 3)
            res["kernels_ms_per_step"] = {k: round(v, 3) for k, v in by.items()}
        what = []
        if graph:
            what.append("one CUDA graph per APGD iteration (graphs.GraphedAttack: step + forward + loss + input-gradient "
                        "backward + bookkeeping), the early-stop flag polled one iteration late")
        if fuse:
            what.append("robseg_loss_upsampled_fwd_bwd: the [B,C,512,512] logits / dlogits never exist")
        res["what"] = "; ".join(what)
        del model
        torch.cuda.empty_cache()
        return res
    except Exception as e:
        torch.cuda.empty_cache()
        return {"unavailable": repr(e)[:200]}


def reference_on_gpu(args, mods, dev, x, y, w):
    """The honest same-hardware comparator: the UNMODIFIED reference attacker (baseline/_ref
    semseg/attacker.py, stock ATen op chains) driving the reference's own model (stock F.interpolate)
    on this GPU, same SEA step, re-forward of every adversarial batch as tools/infer.py:82-90 does.
    One warm-up step, one timed step."""
    import torch

    try:
        if not import_reference(mods):
            return {"unavailable": "baseline/_ref did not travel"}
        import semseg.attacker as RA

        if RA.__name__.startswith("robseg_b200"):
            return {"unavailable": "drop-in is installed in this process"}
        saved = (args.own_consumer,)
        args.own_consumer = False
        model, desc = build_consumer(args, mods, dev, accelerated=False)
        args.own_consumer = saved[0]
        for p in model.parameters():
            p.requires_grad_(True)
        w_cpu = w.cpu()  # tools/infer.py:297-301 hands the attack a CPU weight tensor
        B = x.shape[0]

        def step():
            for loss in LOSSES:
                x_adv, _, acc = RA.apgd_largereps(model, x.clone(), y, w_cpu, norm="Linf", eps=args.eps / 255.0,
                                                  n_iter=args.n_iter, loss=loss, track_loss="ce-avg", use_rs=True,
                                                  early_stop=True, num_classes=args.classes)
                with torch.no_grad():
                    model(x_adv).max(1)[1]

        torch.cuda.reset_peak_memory_stats(dev)
        torch.manual_seed(1234)
        step()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.manual_seed(1234)
        t0.record()
        step()
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1)
        peak = torch.cuda.max_memory_allocated(dev) / 2**30
        del model
        torch.cuda.empty_cache()
        return {"value": round(B * args.n_iter * len(LOSSES) / (ms / 1e3), 3), "unit": UNIT,
                "ms_per_step": round(ms, 1), "peak_mem_GiB": round(peak, 1), "steps": 1, "warmup": 1,
                "what": "unmodified reference semseg/attacker.py::apgd_largereps + " + desc +
                        " with stock F.interpolate, same step incl. the re-forward of x_adv, this GPU"}
    except Exception as e:  # a comparator must never take the bench down
        torch.cuda.empty_cache()
        return {"unavailable": repr(e)[:200]}


def run_ours(args):
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge

    ge.load_package()
    from importlib import import_module

    mods = {k: import_module("robseg_b200." + v) for k, v in dict(
        attacker="semseg.attacker", ops="ops", dist="dist", lib="_lib", consumers="consumers",
        worse="tools.worse_only", graphs="graphs", dropin="dropin", val="semseg.val").items()}
    mods["lib"].load()  # fails loudly if the CUDA extension is missing
    assert torch.cuda.is_available(), "bench.py --impl ours needs a GPU (no CPU fallback)"
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.micro:
        return run_micro(args, mods, dev, rank, world)
    if args.workload == "pirat":
        return run_pirat(args, mods, dev, rank, world, local)

    model, consumer_desc = build_consumer(args, mods, dev)
    for p in model.parameters():
        p.requires_grad_(True)  # as in the reference: parameters keep requires_grad
    B, C, S = args.batch, args.classes, args.size
    if args.graph:
        model = mods["graphs"].GraphedModel(model, torch.rand(B, 3, S, S, device=dev))
    w = (0.5 + torch.rand(C, generator=torch.Generator().manual_seed(1))).to(dev)  # class-balance weights
    x, y = make_batch(B, C, S, 100 + rank, dev)
    hx, hy = make_batch(B, C, S, 100 + rank, pin=True)
    iters_per_step = B * args.n_iter * len(LOSSES)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n_steps, e2e):
        barrier()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        keep = None
        for _ in range(n_steps):
            torch.manual_seed(1234)
            worst, counters, host = sea_step(mods, model, x, y, w, args, world, (hx, hy) if e2e else None)
            keep = (worst, counters, host)
        t1.record()
        barrier()
        ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms), keep

    for i in range(args.warmup):
        torch.manual_seed(1234)
        # the last warm-up step goes through the host-buffer path once: it allocates (pins) the result buffers
        # of the e2e arm, a one-off ~0.1 s of cudaHostAlloc that is not part of a step
        sea_step(mods, model, x, y, w, args, world, (hx, hy) if i == args.warmup - 1 else None)
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = mods["lib"].launches
    mods["ops"].profile_start()
    ms, keep = timed(args.steps, e2e=False)
    prof = mods["ops"].profile_stop()
    prof_k = mods["ops"].profile_kernels()
    launches = mods["lib"].launches - launches0
    ms_e2e, keep_e2e = timed(args.steps, e2e=True)
    clocks = sampler.stop() if rank == 0 else None
    peak_gib = torch.cuda.max_memory_allocated(dev) / 2**30

    # epilogue (once per run, outside the steps): evalSEA's sequential greedy worst-case mIoU
    t_ep = time.time()
    gi, gt, gp = keep[1]
    miou, _ = mods["worse"].greedy_worst_miou(gi.cpu().numpy(), (gt + gp - gi).cpu().numpy())
    t_ep = (time.time() - t_ep) * 1e3

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = world * iters_per_step * args.steps / (ms / 1e3)
    value_e2e = world * iters_per_step * args.steps / (ms_e2e / 1e3)
    peaks = load_peaks()
    by = {}
    for name, nbytes, t in prof:
        d = by.setdefault(name, [0, 0.0, 0])
        d[0] += nbytes
        d[1] += t
        d[2] += 1
    byk = {}  # the library's own brackets around the main kernel of every loss call (robseg_profile_next_kernel)
    for name, nbytes, t in prof_k:
        d = byk.setdefault(name, [0, 0.0, 0])
        d[0] += nbytes
        d[1] += t
        d[2] += 1
    ours_ms = sum(v[1] for v in by.values())
    h2d = B * 3 * S * S * 4 + B * S * S * 8
    d2h = len(LOSSES) * B * 3 * S * S * 4 + B * world * 4
    line = {
        "metric": METRIC if args.model == "upernet" else METRIC.replace("UperNet-ConvNeXt-T", f"Segmenter-ViT-{args.variant}"),
        "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": ("configs[1]" if args.model == "upernet" else "configs[2]") +
                        f": full SEA (mask-ce-bal, mask-ce-avg, js-avg) x apgd_largereps n_iter={args.n_iter} "
                        f"({'/'.join(map(str, stage_iters(args.n_iter)))} @ 2eps/1.5eps/eps), " +
                        (f"UperNet-ConvNeXt-{args.variant}_CVST" if args.model == "upernet" else
                         f"Segmenter-ViT-{args.variant}/16 + mask transformer") +
                        f" random init, {C} classes, {S}x{S}, batch {B} per GPU, eps {args.eps:g}/255",
            "image_iterations_per_step": iters_per_step * world,
            "model_fwd_per_step": len(LOSSES) * (args.n_iter + 3 + (1 if args.reforward else 0)),
            "model_bwd_per_step": len(LOSSES) * args.n_iter,
            "adversarial_argmax": "re-forward of x_adv (tools/infer.py:82-90)" if args.reforward else
            "returned by the attack (return_pred=True, SURVEY 8f-2; --reforward restores the re-forward)" if args.pred_maps
            else "never materialised: the attack returns the per-image class counters of its adversarial point, taken in "
                 "the loss kernel's argmax pass (return_counts=True; --pred-maps / --reforward restore the earlier flows)",
            "consumer_model": consumer_desc,
            "peak_mem_GiB": round(peak_gib, 1),
            "final_logit_upsampling": ("fused into the loss kernel (robseg_loss_upsampled_fwd_bwd): the [B,C,H,W] "
                                       "logits / dlogits never exist" if args.fuse_loss and hasattr(
                                           getattr(model, "model", model), "forward_lowres")
                                       else "separate robseg up-sampling kernels + loss_tma_kernel"),
            "consumer": "stock PyTorch fp32 (cuDNN conv TF32 default, matmul fp32)" +
                        (", forward / input-gradient backward replayed as CUDA graphs" if args.graph else ""),
            "bilinear_upsample": {False: "F.interpolate (stock) everywhere",
                                  True: "robseg_upsample_bilinear_fwd/_bwd for the final logit up-sampling (SURVEY 8f-1)",
                                  "all": "robseg_upsample_bilinear_fwd/_bwd for the final logit up-sampling (SURVEY 8f-1) "
                                         "and the decode head's pyramid up-samplings; --stock-upsample restores F.interpolate"
                                  }[upsample_mode(args)],
            "l2_note": "inputs larger than L2: logits/dlogits 2x%.2f GB per launch" % (B * C * S * S * 4 / 1e9),
            "parallelism": f"image-sharded dp{world}, one int64 all-reduce per step",
            "epilogue": "evalSEA greedy worst-case mIoU once per run outside the steps: %.1f ms (mIoU %.4f)" % (t_ep, miou),
            "attack_side_ms_per_step": round(ours_ms / args.steps, 3),
            "attack_side_frac_of_step": round(ours_ms / ms, 4),
            "kernels_ms_per_step": {k: round(v[1] / args.steps, 3) for k, v in by.items()},
        },
        "clocks": clocks,
        "e2e": {"value": round(value_e2e, 3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": round(ms_e2e / args.steps, 3)},
        "gpu_launches": launches,
        "roofline": sea_roofline(by, byk, B, C, S, peaks, clocks),
    }
    if world == 1 and args.model == "upernet" and not args.fuse_loss and not args.graph and upsample_mode(args):
        del model
        torch.cuda.empty_cache()
        line["config"]["fused_x4_variant"] = variant_step(args, mods, dev, x, y, w, world, iters_per_step, fuse=True)
        line["config"]["graph_variant"] = variant_step(args, mods, dev, x, y, w, world, iters_per_step, graph=True)
        line["config"]["graph_fused_x4_variant"] = variant_step(args, mods, dev, x, y, w, world, iters_per_step,
                                                                fuse=True, graph=True)
    if world == 1 and not args.no_ref_on_gpu:
        model = keep = keep_e2e = None
        torch.cuda.empty_cache()
        line["config"]["reference_on_gpu"] = reference_on_gpu(args, mods, dev, x, y, w)
        r = line["config"]["reference_on_gpu"].get("value")
        if r:
            line["config"]["reference_on_gpu"]["ours_over_reference_same_gpu"] = round(value / r, 3)
    if world == 1:
        line["config"]["loss_kernel_c151"] = loss_kernel_probe(mods, dev, B, 151, S)
    if not args.no_cpu_baseline and world == 1 and args.model == "upernet":
        line["cpu_baseline"] = cpu_baseline(args, budget_s=25.0)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def sea_roofline(by, byk, B, C, S, peaks, clocks):
    """`roofline` of the dominant robseg kernel of a SEA step, from the per-launch CUDA events of the timed region.
    Default step: loss_tma_kernel (HBM-bound, algorithmic bytes = logits in + gradient out + labels); `achieved` is over
    the kernel's OWN launch duration -- events the library records immediately before and after that one launch
    (robseg_profile_next_kernel), every loss+gradient launch of the timed steps, counted or not -- and the wider bracket
    around the whole C call (+ the per-image finalize kernel, + the counter zeroing / fold kernels of counted launches)
    is reported beside it.  With the final up-sampling fused into the loss (SegMenter default, --fuse-loss) the
    [B,C,H,W] tensors never exist: that kernel moves 1/R^2 of the bytes and is bound by the ex2 pipe (2 ex2 per
    up-sampled logit, 16 lanes / clk / SM), so its HBM fraction is small BY DESIGN; the ex2-pipe fraction is reported
    beside it."""
    def agg(d, *names):
        tot = [0, 0.0, 0]
        for n in names:
            v = d.get(n)
            if v:
                tot = [tot[0] + v[0], tot[1] + v[1], tot[2] + v[2]]
        return tot if tot[2] else None

    def rate(v):
        return v[0] / (v[1] / 1e3) / 1e9

    if "loss_grad" in by or "loss_grad_counts" in by:
        call = agg(by, "loss_grad", "loss_grad_counts")
        kern = agg(byk, "loss_grad", "loss_grad_counts") or call
        achieved = rate(kern)
        traffic, capture = load_traffic("sea_c%d" % C)
        roof = {"bound": "hbm", "kernel": "loss_tma_kernel<float,VEC=2,G=1> (fused loss+dlogits, C=%d)" % C,
                "achieved": round(achieved, 1), "peak": peaks[0], "unit": "GB/s", "frac": round(achieved / peaks[0], 4),
                "traffic": traffic, "traffic_capture": capture, "peak_source": peaks[1], "launches_timed": kern[2],
                "avg_launch_ms": round(kern[1] / max(kern[2], 1), 4),
                "algorithmic_bytes_per_launch": kern[0] // max(kern[2], 1),
                "timed_with": ("events recorded by the library around the kernel launch itself (robseg_profile_next_kernel)"
                               if kern is not call else "events around the C call"),
                "frac_of_spec_8TBps": round(achieved / 8000.0, 4)}
        # the brackets around the whole C call: + loss_finalize_kernel, and for the last stage's launches (which also
        # take the per-image class counters) + counts_zero_kernel and counts_fold_kernel
        for key, names, what in (
                ("call_bracket", ("loss_grad",), "loss_tma_kernel + loss_finalize_kernel"),
                ("call_bracket_with_class_counters", ("loss_grad_counts",),
                 "counts_zero_kernel + loss_tma_kernel + counts_fold_kernel + loss_finalize_kernel")):
            v = agg(by, *names)
            if v:
                roof[key] = {"launches_timed": v[2], "avg_bracket_ms": round(v[1] / v[2], 4),
                             "achieved": round(rate(v), 1), "frac": round(rate(v) / peaks[0], 4), "bracket": what}
        for key, names in (("kernel_uncounted", ("loss_grad",)), ("kernel_with_class_counters", ("loss_grad_counts",))):
            v = agg(byk, *names)
            if v:
                roof[key] = {"launches_timed": v[2], "avg_launch_ms": round(v[1] / v[2], 4),
                             "frac": round(rate(v) / peaks[0], 4)}
        return roof
    lg = (agg(byk, "loss_up_grad", "loss_up_grad_counts") or agg(by, "loss_up_grad", "loss_up_grad_counts")
          or [0, 1e-9, 1])
    achieved = lg[0] / (lg[1] / 1e3) / 1e9
    mhz = (clocks or {}).get("sm_mhz") or 1965.0
    ex2_peak = 16 * 148 * mhz * 1e6  # MUFU.EX2 lanes per second (B300_MICROARCH: 16 / clk / SM)
    ex2_rate = 2.0 * B * C * S * S * lg[2] / (lg[1] / 1e3)  # pass 2 + pass 3: two ex2 per up-sampled logit
    return {"bound": "hbm", "kernel": "loss_up_kernel<R> (loss taken through the bilinear up-sampling, C=%d)" % C,
            "achieved": round(achieved, 1), "peak": peaks[0], "unit": "GB/s", "frac": round(achieved / peaks[0], 4),
            "traffic": load_traffic("sea_up_c%d" % C)[0], "peak_source": peaks[1], "launches_timed": lg[2],
            "avg_launch_ms": round(lg[1] / max(lg[2], 1), 4), "algorithmic_bytes_per_launch": lg[0] // max(lg[2], 1),
            "note": "not HBM-bound by construction (the [B,C,H,W] logits / gradient never exist): limited by the ex2 pipe",
            "ex2_pipe": {"achieved_Gex2_per_s": round(ex2_rate / 1e9, 1), "peak_Gex2_per_s": round(ex2_peak / 1e9, 1),
                         "frac": round(ex2_rate / ex2_peak, 4)},
            "unfused_equivalent_GBps": round((2 * B * C * S * S * 4 + 8 * B * S * S) * lg[2] / (lg[1] / 1e3) / 1e9, 1)}


def loss_kernel_probe(mods, dev, B, C, S, reps=5):
    """Fused loss+dlogits launch alone at [B,C,S,S] fp32 (inputs 2 x B*C*S*S*4 bytes, far beyond L2):
    median CUDA-event time over `reps` launches after 3 warm-ups -> GB/s against the measured peak.
    Used for the class count the default step does not run (151 = the reference's ADE20K config,
    configs/ade20k_convnext.yaml:14)."""
    import torch

    ops = mods["ops"]
    g = torch.Generator(device=dev).manual_seed(7)
    z = 3 * torch.randn(B, C, S, S, device=dev, generator=g)
    y = torch.randint(0, C, (B, S, S), device=dev, generator=g)
    y = torch.where(torch.rand(B, S, S, device=dev, generator=g) < 0.5, z.argmax(1), y)
    dbuf = torch.empty_like(z)
    ts, tc = [], []
    for i in range(3 + reps):
        # the wrapper's own events bracket just the C call (kernel + per-image finaliser): outer events would
        # also count the host's argument marshalling while the GPU sits idle
        ops.profile_start()
        ops.loss_fwd_bwd(z, y, "mask-ce-avg", None, dlogits_out=dbuf)
        torch.cuda.synchronize()
        t = sum(ms_ for _, _, ms_ in ops.profile_stop())
        tk = sum(ms_ for _, _, ms_ in ops.profile_kernels())  # the loss kernel's own launch
        if i >= 3:
            ts.append(tk or t)
            tc.append(t)
    ms = statistics.median(ts)
    nbytes = 2 * z.numel() * 4 + 8 * y.numel()
    peak = load_peaks()[0]
    del z, y, dbuf
    torch.cuda.empty_cache()
    mc = statistics.median(tc)
    return {"shape": [B, C, S, S], "ms": round(ms, 4), "GBps": round(nbytes / ms / 1e6, 1),
            "frac": round(nbytes / ms / 1e6 / peak, 4), "bytes": nbytes,
            "call_bracket_ms": round(mc, 4), "call_bracket_frac": round(nbytes / mc / 1e6 / peak, 4)}


# ------------------------------------------------------------------------------ PIR-AT (configs[3])
PIRAT_METRIC = "PIR-AT training images/sec (UperNet-ConvNeXt-S, 2-step PGD inner attack, DDP, 512x512)"


def run_pirat(args, mods, dev, rank, world, local):
    """BASELINE configs[3]: one PIR-AT training step per rank and step -- eval-mode 2-step PGD inner attack
    (semseg/val.py:181-218 semantics, eps 4/255, alpha 1e-2, loss "pgd") through the fused loss / step
    kernels, then the train-mode forward (main + 0.4 aux CE), backward and AdamW step -- on the
    reference's UperNet-ConvNeXt-S under DistributedDataParallel(find_unused_parameters=True), as
    tools/train_rob_seg.py:143-145,293-352 runs it.  Timed in three modes: the drop-in default
    (loss.backward() semantics, SURVEY 9-Q7: every attack step also produces parameter gradients and a
    DDP all-reduce), input_grad_only=True, and the unmodified reference attack on the same GPUs."""
    import torch
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP

    B, C, S = args.batch, args.classes, args.size
    model, consumer_desc = build_consumer(args, mods, dev)
    model.train()
    for p in model.parameters():
        p.requires_grad_(True)
    n_param = sum(p.numel() for p in model.parameters())
    net = DDP(model, device_ids=[local], find_unused_parameters=True) if world > 1 else model
    opt = torch.optim.AdamW(net.parameters(), lr=1e-4, weight_decay=0.05)
    g = torch.Generator().manual_seed(100 + rank)
    hx = torch.rand(B, 3, S, S, generator=g).pin_memory()
    hy = torch.randint(0, C, (B, S, S), generator=g).pin_memory()
    x, y = hx.to(dev), hy.to(dev)
    val = mods["val"]
    n_attack = 2
    attacks = {
        "dropin": val.Pgd_Attack_1(epsilon=4 / 255, alpha=1e-2, num_iter=n_attack, los="pgd"),
        "input_grad_only": val.Pgd_Attack_1(epsilon=4 / 255, alpha=1e-2, num_iter=n_attack, los="pgd",
                                            input_grad_only=True),
    }
    if not args.no_ref_on_gpu:
        try:
            if import_reference(mods):
                import semseg.val as RV

                if not RV.Pgd_Attack_1.__module__.startswith("robseg_b200"):
                    # Pgd_Attack(epsilon=...) raises TypeError in the reference; Pgd_Attack_1
                    # matches the trainer's keywords and the scalar "pgd" loss
                    attacks["reference_on_gpu"] = RV.Pgd_Attack_1(epsilon=4 / 255, alpha=1e-2, num_iter=n_attack,
                                                                  los="pgd")
        except Exception as e:
            print(f"[bench] reference attack unusable ({e!r})", file=sys.stderr)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(attack, e2e):
        xi, yi = (hx.to(dev, non_blocking=True), hy.to(dev, non_blocking=True)) if e2e else (x, y)
        opt.zero_grad(set_to_none=True)
        net.eval()
        adv = attack.adv_attack(net, xi, yi)[0]  # tools/train_rob_seg.py:333-336
        net.train()
        loss, _ = net(adv, yi)
        loss.backward()
        opt.step()
        return loss.item() if e2e else loss

    def timed(attack, n, e2e=False):
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(n):
            last = step(attack, e2e)
        t1.record()
        barrier()
        ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms), last

    res = {}
    sampler = ClockSampler(local)
    clocks = None
    for name, attack in attacks.items():
        for _ in range(args.warmup):
            step(attack, False)
        if name == "dropin":
            if rank == 0:
                sampler.start()
            launches0 = mods["lib"].launches
            mods["ops"].profile_start()
            ms, last = timed(attack, args.steps)
            prof = mods["ops"].profile_stop()
            prof_k = mods["ops"].profile_kernels()
            launches = mods["lib"].launches - launches0
            ms_e2e, _ = timed(attack, args.steps, e2e=True)
            clocks = sampler.stop() if rank == 0 else None
            res[name] = ms
            assert torch.isfinite(last).all()
        else:
            res[name], _ = timed(attack, args.steps)
    # replicas must still agree after all those steps
    chk = torch.stack([p.detach().double().sum() for p in model.parameters()]).sum().reshape(1)
    if world > 1:
        allc = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        assert all(torch.equal(allc[0], c) for c in allc), "DDP replicas diverged"
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = load_peaks()
    by = {}
    for name, nbytes, t in prof:
        d = by.setdefault(name, [0, 0.0, 0])
        d[0] += nbytes
        d[1] += t
        d[2] += 1
    lg = by.get("loss_grad", [0, 1e-9, 1])
    lk = [0, 0.0, 0]  # the loss kernel's own launches (robseg_profile_next_kernel)
    for name, nbytes, t in prof_k:
        if name == "loss_grad":
            lk = [lk[0] + nbytes, lk[1] + t, lk[2] + 1]
    if lk[2]:
        lg = lk
    achieved = lg[0] / (lg[1] / 1e3) / 1e9
    ips = lambda m: round(world * B * args.steps / (m / 1e3), 3)  # noqa: E731
    line = {
        "metric": PIRAT_METRIC, "value": ips(res["dropin"]), "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(res["dropin"] / args.steps, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": f"configs[3]: PIR-AT step, UperNet-ConvNeXt-{args.variant}_CVST random init, {C} classes, {S}x{S}, "
                        f"batch {B} per GPU, {n_attack}-step PGD inner attack (Pgd_Attack_1 semantics, eps 4/255, alpha 1e-2, "
                        "loss pgd) + train forward/backward + AdamW(lr 1e-4, wd 0.05), "
                        + (f"DistributedDataParallel x{world} (find_unused_parameters=True)" if world > 1 else "single process"),
            "consumer_model": consumer_desc,
            "parallelism": f"dp{world} (stock DDP gradient all-reduce over NCCL; attack kernels are rank-local)",
            "l2_note": "inputs larger than L2: logits/dlogits 2x%.2f GB per loss launch" % (B * C * S * S * 4 / 1e9),
            "modes_images_per_s": {k: ips(v) for k, v in res.items()},
            "modes_ms_per_step": {k: round(v / args.steps, 2) for k, v in res.items()},
            "attack_time_allreduce_bytes_per_step": {
                "dropin (loss.backward() semantics, SURVEY 9-Q7)": n_param * 4 * n_attack if world > 1 else 0,
                "input_grad_only": 0, "training backward": n_param * 4 if world > 1 else 0},
            "parameters": n_param,
            "attack_side_ms_per_step": round(sum(v[1] for v in by.values()) / args.steps, 3),
            "kernels_ms_per_step": {k: round(v[1] / args.steps, 3) for k, v in by.items()},
        },
        "clocks": clocks,
        "e2e": {"value": ips(ms_e2e), "unit": "images/s", "h2d_bytes_per_step": B * 3 * S * S * 4 + B * S * S * 8,
                "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e / args.steps, 3)},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "loss_tma_kernel<float,VEC=2,G=1> (fused CE loss+dlogits, C=%d)" % C,
                     "achieved": round(achieved, 1), "peak": peaks[0], "unit": "GB/s",
                     "frac": round(achieved / peaks[0], 4), "traffic": load_traffic("sea_c%d" % C)[0],
                     "peak_source": peaks[1], "launches_timed": lg[2],
                     "avg_launch_ms": round(lg[1] / max(lg[2], 1), 4),
                     "algorithmic_bytes_per_launch": lg[0] // max(lg[2], 1)},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_source_sha():
    """sha256 over the sources the loss kernels are compiled from (what an ncu traffic capture is valid for)."""
    import hashlib

    h = hashlib.sha256()
    for rel in ("loss_kernel.cu", "common.cuh"):
        with open(os.path.join(ROOT, "robust-segmentation_b200", "csrc", rel), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def load_traffic(key):
    """(DRAM bytes per launch from the committed ncu --set full capture, provenance).  DRAM bytes need ncu, so the
    bench cannot measure them live; the capture records the sha of the kernel sources it was taken on
    (scripts/update_traffic.py) and a capture of other sources is reported as null, not passed on as current."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            d = json.load(f)
    except Exception:
        return None, None
    v = d.get(key)
    if v is None:
        return None, None
    try:
        cur = kernel_source_sha()
    except OSError:  # sources not shipped with the library: provenance cannot be checked
        cur = None
    sha = d.get("_kernel_source_sha")
    cap = {"report": d.get("_reports", {}).get(key), "kernel_source_sha": sha, "matches_tree": sha == cur}
    if sha != cur:
        cap["stale_value"] = v
        return None, cap
    return v, cap


# ------------------------------------------------------------------------------ microbench
def stage_iters(n_iter):
    """apgd_largereps' three-stage split (semseg/attacker.py:693-694)."""
    a = int(0.3 * n_iter)
    return [a, a, n_iter - 2 * a]


def upsample_mode(args):
    """consumers.UperNetConvNeXt.fast_upsample value for the command line."""
    return False if args.stock_upsample else (True if args.logit_upsample_only else "all")


def run_micro(args, mods, dev, rank, world):
    """BASELINE config 5: loss+dlogits, APGD step and histogram kernels alone at
    [micro_batch,150,512,512]; one JSON line with per-kernel GB/s vs the HBM roofline."""
    import torch

    ops = mods["ops"]
    B, C, S = args.micro_batch, args.classes, args.size
    dt = torch.float32 if args.micro_dtype == "fp32" else torch.bfloat16
    g = torch.Generator(device=dev).manual_seed(rank)
    z = (3 * torch.randn(B, C, S, S, device=dev, generator=g)).to(dt)
    y = torch.randint(0, C, (B, S, S), device=dev, generator=g)
    y = torch.where(torch.rand(B, S, S, device=dev, generator=g) < 0.5, z.argmax(1), y)
    w = 0.5 + torch.rand(C, device=dev, generator=g)
    dbuf = torch.empty_like(z)
    x = torch.rand(B, 3, S, S, device=dev, generator=g)
    xa, xo, gr, xn = (torch.rand_like(x) for _ in range(4))
    step = torch.full((B,), 16 / 255, device=dev)
    peaks = load_peaks()
    res = {}
    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def time_it(name, fn, nbytes, inner=None):
        """Median device time of fn().  inner: time only the events the wrapper records around
        its own C call (ops.profile_start) -- for kernels shorter than the host-side work of a
        call (output allocation, argument marshalling), which outer events would measure instead."""
        for _ in range(max(args.warmup, 3)):
            fn()
        torch.cuda.synchronize()
        if world > 1:  # all GPUs run the same kernel at the same time
            import torch.distributed as dist

            dist.barrier()
        ts = []
        for _ in range(max(args.steps, 5)):
            if nbytes < 2 * l2_flush.numel():  # working set could stay in the 126 MB L2: evict it
                l2_flush.zero_()
            # keep the stream busy (~0.5 ms spin kernel) while the host enqueues the call: the events below then bracket
            # back-to-back device execution, as inside the attack loop where the GPU queue never drains; without it a
            # ~0.2 ms kernel is charged 15-45 us of host-side launch work (tensor-map encode, attribute set, 2-3 launches)
            torch.cuda._sleep(1_000_000)
            if inner:
                ops.profile_start()
                fn()
                torch.cuda.synchronize()
                ts.append(sum(ms for n, _, ms in ops.profile_stop() if n == inner))
                continue
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = statistics.median(ts)
        res[name] = {"ms": round(ms, 4), "GBps": round(nbytes / ms / 1e6, 1), "frac": round(nbytes / ms / 1e6 / peaks[0], 4),
                     "bytes": nbytes}
        if name.startswith(("loss_", "argmax")):
            # the main kernel alone (events recorded by the library around that one launch, robseg_profile_next_kernel):
            # the bracket above also holds the finalize kernel and, for counted launches, the counter zeroing / fold
            tk = []
            for _ in range(5):
                if nbytes < 2 * l2_flush.numel():
                    l2_flush.zero_()
                torch.cuda._sleep(1_000_000)
                ops.profile_start()
                fn()
                torch.cuda.synchronize()
                ops.profile_stop()
                tk.append(sum(t for _, _, t in ops.profile_kernels()))
            mk = statistics.median(tk)
            if mk > 0:
                res[name].update({"kernel_ms": round(mk, 4), "kernel_GBps": round(nbytes / mk / 1e6, 1),
                                  "kernel_frac": round(nbytes / mk / 1e6 / peaks[0], 4)})

    es = z.element_size()
    for kind in LOSSES + ["ce-avg"]:
        time_it("loss_grad/" + kind, lambda: ops.loss_fwd_bwd(z, y, kind, w, dlogits_out=dbuf),
                2 * z.numel() * es + 8 * y.numel())
    time_it("loss_only/mask-ce-avg", lambda: ops.loss_fwd_bwd(z, y, "mask-ce-avg", w, want_grad=False),
            z.numel() * es + 8 * y.numel())
    time_it("argmax", lambda: ops.loss_fwd_bwd(z, y, "argmax", want_grad=False, want_pred=True, want_stats=False),
            z.numel() * es + 16 * y.numel())
    # the per-image class counters taken in the same pass (robseg_loss_fwd_bwd_counts): uniformly random labels are
    # the worst case for the warp-aggregated reductions (one per lane), coherent maps the normal one (one per warp row)
    time_it("loss_grad+counts/mask-ce-avg uniform-random", lambda: ops.loss_fwd_bwd(z, y, "mask-ce-avg", w, dlogits_out=dbuf,
                                                                                      want_counts=True),
            2 * z.numel() * es + 8 * y.numel())
    time_it("argmax+counts uniform-random", lambda: ops.loss_fwd_bwd(z, y, "argmax", want_grad=False, want_stats=False,
                                                                      want_counts=True), z.numel() * es + 8 * y.numel())
    time_it("apgd_step", lambda: ops.apgd_step(x, xa, xo, gr, step, 8 / 255, 0.75, xn), 20 * x.numel(), "apgd_step")
    pred = z.argmax(1)
    time_it("pixel_hist/counts uniform-random", lambda: ops.pixel_hist(pred, y, C), 16 * y.numel(), "pixel_hist")
    time_it("pixel_hist/full uniform-random", lambda: ops.pixel_hist(pred, y, C, want_hist=True, want_counts=False), 16 * y.numel(),
            "pixel_hist")
    ys = torch.where(torch.rand(B, S, S, device=dev, generator=g) < 0.8, torch.full_like(y, 3), y)  # 80 % one class
    ps = torch.where(torch.rand(B, S, S, device=dev, generator=g) < 0.7, ys, pred)
    time_it("pixel_hist/counts skewed-80pct", lambda: ops.pixel_hist(ps, ys, C), 16 * y.numel(), "pixel_hist")
    nb = (S + 63) // 64
    blk = torch.randint(0, C, (B, nb, nb), device=dev, generator=g)  # 64x64 constant regions
    yc = blk.repeat_interleave(64, 1).repeat_interleave(64, 2)[:, :S, :S].contiguous()
    pblk = torch.where(torch.rand(blk.shape, device=dev, generator=g) < 0.7, blk, torch.randint_like(blk, C))
    pc = pblk.repeat_interleave(64, 1).repeat_interleave(64, 2)[:, :S, :S].contiguous()  # region-wise right / wrong
    time_it("pixel_hist/counts piecewise-constant-64x64", lambda: ops.pixel_hist(pc, yc, C), 16 * y.numel(), "pixel_hist")
    time_it("loss_grad+counts/mask-ce-avg piecewise-constant-64x64 labels, random predictions",
            lambda: ops.loss_fwd_bwd(z, yc, "mask-ce-avg", w, dlogits_out=dbuf, want_counts=True),
            2 * z.numel() * es + 8 * y.numel())
    # the realistic case: labels AND predictions spatially coherent (argmax = pc: region-wise right / wrong)
    z.scatter_add_(1, pc.unsqueeze(1), torch.full((B, 1, S, S), 30.0, device=dev, dtype=dt))
    time_it("loss_grad/mask-ce-avg coherent labels and predictions",
            lambda: ops.loss_fwd_bwd(z, yc, "mask-ce-avg", w, dlogits_out=dbuf), 2 * z.numel() * es + 8 * y.numel())
    time_it("loss_grad+counts/mask-ce-avg coherent labels and predictions",
            lambda: ops.loss_fwd_bwd(z, yc, "mask-ce-avg", w, dlogits_out=dbuf, want_counts=True),
            2 * z.numel() * es + 8 * y.numel())
    z.scatter_add_(1, pc.unsqueeze(1), torch.full((B, 1, S, S), -30.0, device=dev, dtype=dt))
    time_it("pixel_hist/full piecewise-constant-64x64", lambda: ops.pixel_hist(pc, yc, C, want_hist=True, want_counts=False),
            16 * y.numel(), "pixel_hist")
    if args.micro_dtype == "fp32":
        import torch.nn.functional as F

        Bu = min(B, 16)
        low = torch.randn(Bu, C, S // 4, S // 4, device=dev, generator=g)
        gup = torch.randn(Bu, C, S, S, device=dev, generator=g)
        nb = 4 * (low.numel() + gup.numel())
        time_it("upsample_fwd x4 (ours)", lambda: ops._upsample_fwd(low, S, S), nb, "upsample_fwd")
        time_it("upsample_bwd x4 (ours)", lambda: ops._upsample_bwd(gup, S // 4, S // 4), nb, "upsample_bwd")
        time_it("upsample_fwd x4 (ATen)", lambda: F.interpolate(low, size=(S, S), mode="bilinear", align_corners=False), nb)
        lr = low.clone().requires_grad_()
        up = F.interpolate(lr, size=(S, S), mode="bilinear", align_corners=False)
        time_it("upsample_bwd x4 (ATen)", lambda: torch.autograd.grad(up, [lr], grad_outputs=gup, retain_graph=True), nb)
        del up, lr, low, gup
    # the stock ATen op chains of the reference on the same device (SURVEY 8d "also report"):
    # what the fused kernels replace, timed on identical inputs
    if args.micro_dtype == "fp32" and B <= 16:
        import torch.nn.functional as F

        def aten_maskce():  # semseg/attacker.py:143-152,237-240 + autograd.grad
            zz = z.detach().requires_grad_()
            mask = (zz.max(1)[1] == y) * (y != -1)
            l = (mask.float().detach() * F.cross_entropy(zz, y, reduction="none", ignore_index=-1))
            torch.autograd.grad(l.view(B, -1).mean(-1).sum(), [zz])

        def aten_js():  # semseg/attacker.py:187-234
            zz = z.detach().requires_grad_()
            p_ = F.softmax(zz, 1)
            q_ = F.one_hot(y.view(B, -1), C).permute(0, 2, 1).view(p_.shape).float()
            m_ = (p_ + q_) / 2
            l = ((F.kl_div(m_.log(), p_, reduction="none") + F.kl_div(m_.log(), q_, reduction="none")) / 2).sum(1)
            torch.autograd.grad(l.view(B, -1).mean(-1).sum(), [zz])

        def aten_track_and_acc():  # :353-361,370-373,485-490: second log-softmax + two more max(1)
            F.cross_entropy(z, y, reduction="none", ignore_index=-1).view(B, -1).mean(-1)
            (z.max(1)[1] == y).float().view(B, -1).mean(-1)
            z.max(1)[1]

        def aten_step():  # :388-410
            eps, a = 8 / 255, 0.75
            st = step.view(-1, 1, 1, 1)
            g2 = xa - xo
            x1 = xa + st * torch.sign(gr)
            x1 = torch.clamp(torch.min(torch.max(x1, x - eps), x + eps), 0.0, 1.0)
            torch.clamp(torch.min(torch.max(xa + (x1 - xa) * a + g2 * (1 - a), x - eps), x + eps), 0.0, 1.0)

        def aten_iou_acc():  # compute_iou_acc, :9-52: 2*C masked reductions + 3 host syncs
            acc_cls, n_pxl = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
            int_cls, uni = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
            hit = pred == y
            for cl in range(C):
                ind = y == cl
                acc_cls[cl] += (hit * ind).float().sum()
                n_pxl[cl] += ind.float().sum()
            (acc_cls[n_pxl > 0] / n_pxl[n_pxl > 0]).mean().cpu()
            (acc_cls.sum() / n_pxl.sum()).cpu()
            for cl in range(C):
                ind = y == cl
                s_ = hit[ind].float().sum()
                int_cls[cl] += s_
                uni[cl] += ind.float().sum() + (pred == cl).float().sum() - s_
            (int_cls[uni > 0] / uni[uni > 0]).mean().cpu()

        nb = 2 * z.numel() * es + 8 * y.numel()
        time_it("ATen chain: mask-ce-avg loss+grad", aten_maskce, nb)
        time_it("ATen chain: js-avg loss+grad", aten_js, nb)
        time_it("ATen chain: track CE + accuracy + argmax", aten_track_and_acc, 0)
        time_it("ATen chain: APGD step", aten_step, 20 * x.numel())
        time_it("ATen chain: compute_iou_acc", aten_iou_acc, 16 * y.numel())
    # N > 1 (BASELINE configs[4]: "swept at 1/2/4/8 GPUs"): every rank ran the same kernels on its own GPU at
    # the same time; the aggregate is the sum of the per-rank rates, the slowest rank is reported beside it
    if world > 1:
        import torch.distributed as dist

        names = sorted(res)
        mine = torch.tensor([res[n]["ms"] for n in names], dtype=torch.float64, device=dev)
        allms = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allms, mine)
        allms = torch.stack(allms).cpu()
        for j, n in enumerate(names):
            nb = res[n]["bytes"]
            per_rank = [nb / float(ms) / 1e6 if float(ms) > 0 else 0.0 for ms in allms[:, j]]
            res[n].update({"ms_max_over_ranks": round(float(allms[:, j].max()), 4),
                           "GBps_aggregate": round(sum(per_rank), 1), "GBps_slowest_rank": round(min(per_rank), 1),
                           "frac_slowest_rank": round(min(per_rank) / peaks[0], 4)})
    if rank == 0:
        k = res["loss_grad/mask-ce-avg"]
        agg = k.get("GBps_aggregate", k["GBps"])
        print(json.dumps({
            "metric": "attack-kernel microbench (loss+dlogits GB/s, all GPUs)", "value": agg, "unit": "GB/s",
            "n_gpus": world, "steps": max(args.steps, 5), "warmup": max(args.warmup, 3),
            "ms_per_step": k.get("ms_max_over_ranks", k["ms"]),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.micro_dtype,
            "data": "synthetic", "config": {"workload": f"configs[4]: logits {B}x{C}x{S}x{S} {args.micro_dtype} per GPU, "
                                                        f"{world} GPU(s) running concurrently",
                                            "kernels": res, "l2_note": "inputs larger than L2 (smaller ones: 256 MB flush between launches)",
                                            "timing": "CUDA events on the launching stream around each call, median of >= 5; a "
                                                      "0.5 ms spin kernel precedes every timed call so the events bracket "
                                                      "back-to-back device execution, not host launch latency"},
            "roofline": {"bound": "hbm", "achieved": k.get("GBps_slowest_rank", k["GBps"]), "peak": peaks[0], "unit": "GB/s",
                         "frac": k.get("frac_slowest_rank", k["frac"]),
                         "traffic": load_traffic("micro_c%d_%s" % (C, args.micro_dtype))[0], "peak_source": peaks[1]},
        }), flush=True)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


# ------------------------------------------------------------------------------ CPU arms
def _cpu_sample(args, loss, seed, prefer_reference=True, n_iter=None):
    """One bounded CPU sample of the same workload: apgd_largereps with the GPU arm's n_iter (10 ->
    stages 3/3/4: 13 forwards + 10 backwards, the same model passes per counted image-iteration as
    the GPU arm) on ONE 512x512 image, C classes, UperNet-ConvNeXt-T on the host cores.
    Returns (seconds, image_iterations, kind)."""
    import numpy as np
    import torch

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    C, S, n_iter, B = args.classes, args.size, n_iter or args.n_iter, 1
    x, y = make_batch(B, C, S, seed)
    w = 0.5 + torch.rand(C, generator=torch.Generator().manual_seed(1))
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if prefer_reference and os.path.isdir(os.path.join(ref_dir, "semseg")):
        try:
            import ref_shims

            ref_shims.install()
            if ref_dir not in sys.path:
                sys.path.insert(0, ref_dir)
            cwd = os.getcwd()
            os.chdir(ref_dir)
            try:
                import semseg.attacker as RA
                from semseg.models import UperNetForSemanticSegmentation
            finally:
                os.chdir(cwd)
            key = ("ref", C)
            if key not in _CPU_MODELS:
                torch.manual_seed(0)
                _CPU_MODELS[key] = UperNetForSemanticSegmentation("ConvNeXt-T_CVST", C, None).eval()
            model = _CPU_MODELS[key]
            t0 = time.time()
            RA.apgd_largereps(model, x.clone(), y, w, norm="Linf", eps=args.eps / 255.0, n_iter=n_iter, loss=loss,
                              track_loss="ce-avg", use_rs=True, early_stop=True, num_classes=C)
            return time.time() - t0, B * n_iter, "reference"
        except Exception as e:  # fall through to the oracle port
            print(f"[bench] reference copy unusable ({e!r}); using the oracle port", file=sys.stderr)
    import __graft_entry__ as ge
    import robseg_oracle as O

    ge.load_package()
    from importlib import import_module

    key = ("port", C)
    if key not in _CPU_MODELS:
        torch.manual_seed(0)
        _CPU_MODELS[key] = import_module("robseg_b200.consumers").upernet_convnext(args.variant, C).eval()
    model = O.TorchModelAdapter(_CPU_MODELS[key])
    noise = [(2 * torch.rand_like(x) - 1).numpy() for _ in range(3)]
    t0 = time.time()
    O.apgd_largereps(model, x.numpy(), y.numpy(), w.numpy(), eps=args.eps / 255.0, n_iter=n_iter, loss=loss,
                     early_stop=True, use_rs=True, rand_ts=noise)
    _ = np
    return time.time() - t0, B * n_iter, "port"


_CPU_MODELS = {}


def cpu_baseline(args, budget_s=25.0):
    import torch

    t, n, kind = 0.0, 0, None
    for i, loss in enumerate(LOSSES):
        dt, it, kind = _cpu_sample(args, loss, 100 + i)
        t += dt
        n += it
        if t > budget_s:
            break
    return {"value": round(n / t, 4), "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
            "host_cpus": os.cpu_count(),
            "sample": f"{i + 1} of the 3 SEA losses x apgd_largereps(n_iter={args.n_iter} -> "
                      f"{'/'.join(map(str, stage_iters(args.n_iter)))}, same schedule as the GPU arm) on 1 image "
                      f"{args.size}x{args.size}, {args.classes} classes, UperNet-ConvNeXt-T on host cores, {t:.1f} s"}


def run_reference(args):
    import torch

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    # all the host threads the process may use (torchrun pins OMP_NUM_THREADS=1 by default)
    try:
        n_thr = len(os.sched_getaffinity(0))
    except AttributeError:
        n_thr = os.cpu_count() or 1
    torch.set_num_threads(max(1, n_thr))
    if args.workload == "pirat":
        return run_reference_pirat(args, rank, world)
    for i in range(args.warmup):  # untimed: a short schedule is enough to touch every code path
        _cpu_sample(args, LOSSES[i % 3], 50 + i, n_iter=3)
    t, n, kind = 0.0, 0, None
    for i in range(args.steps):
        dt, it, kind = _cpu_sample(args, LOSSES[i % 3], 100 + i)
        t += dt
        n += it
    value = n / t
    st = stage_iters(args.n_iter)
    sample = (f"each step = one SEA loss (rotating {LOSSES}) x apgd_largereps(n_iter={args.n_iter} -> stages "
              f"{'/'.join(map(str, st))}, {args.n_iter + 3} fwd + {args.n_iter} bwd: the GPU arm's schedule) "
              f"on 1 image {args.size}x{args.size}, {args.classes} classes, UperNet-ConvNeXt-T, host cores")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(t / max(args.steps, 1) * 1e3, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"configs[1] on the host CPU, bounded sample: {sample}",
                   "model_fwd_per_image_iteration": (args.n_iter + 3) / args.n_iter,
                   "model_bwd_per_image_iteration": 1.0},
        "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
                         "host_cpus": os.cpu_count(), "sample": sample},
        "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


def run_reference_pirat(args, rank, world):
    """configs[3] on the host CPU: the reference's own Pgd_Attack_1 (semseg/val.py:181-218) + train step on
    ONE 512x512 image per step, UperNet-ConvNeXt-S, all host threads."""
    import torch

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_shims

    ref_shims.install()
    sys.path.insert(0, REF_DIR)
    cwd = os.getcwd()
    os.chdir(REF_DIR)
    try:
        import semseg.val as RV
        from semseg.models import UperNetForSemanticSegmentation
    finally:
        os.chdir(cwd)
    C, S = args.classes, args.size
    torch.manual_seed(0)
    model = UperNetForSemanticSegmentation(f"ConvNeXt-{args.variant}_CVST", C, None)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0.05)
    attack = RV.Pgd_Attack_1(epsilon=4 / 255, alpha=1e-2, num_iter=2, los="pgd")
    real_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self  # val.py:192 hard-codes .cuda()
    g = torch.Generator().manual_seed(100)
    t = 0.0
    try:
        for i in range(args.warmup + args.steps):
            x, y = torch.rand(1, 3, S, S, generator=g), torch.randint(0, C, (1, S, S), generator=g)
            t0 = time.time()
            opt.zero_grad(set_to_none=True)
            model.eval()
            adv = attack.adv_attack(model, x, y)[0]
            model.train()
            loss, _ = model(adv, y)
            loss.backward()
            opt.step()
            if i >= args.warmup:
                t += time.time() - t0
    finally:
        torch.Tensor.cuda = real_cuda
    value = args.steps / t
    sample = (f"each step = the reference's Pgd_Attack_1 (2 steps, loss pgd) + train forward/backward + AdamW on 1 image "
              f"{S}x{S}, {C} classes, UperNet-ConvNeXt-{args.variant}, host cores")
    print(json.dumps({
        "impl": "reference", "metric": PIRAT_METRIC, "value": round(value, 4), "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(t / max(args.steps, 1) * 1e3, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"configs[3] on the host CPU, bounded sample: {sample}"},
        "cpu_baseline": {"value": round(value, 4), "unit": "images/s", "cores": torch.get_num_threads(),
                         "kind": "reference", "host_cpus": os.cpu_count(), "sample": sample},
        "e2e": {"value": round(value, 4), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


def _protect_stdout():
    """Libraries (NCCL's version banner) write to fd 1; the driver wants ONE JSON line there.  Point
    fd 1 at stderr for the run and hand back a file object on the real stdout for the JSON line."""
    real = os.fdopen(os.dup(1), "w")
    sys.stdout.flush()
    os.dup2(2, 1)
    return real


if __name__ == "__main__":
    a = parse()
    _real_stdout = _protect_stdout()
    _print = print

    def print(*args, **kw):  # noqa: A001  (only the final JSON lines go through print(..., flush=True))
        if kw.get("file") is None:
            kw["file"] = _real_stdout
        _print(*args, **kw)
    if a.debug_stack:
        import faulthandler

        faulthandler.dump_traceback_later(a.debug_stack, repeat=True, file=sys.stderr)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
