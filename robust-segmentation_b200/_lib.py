"""ctypes binding of librobseg_b200.so (C ABI declared in include/robseg_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails the
caller gets a RuntimeError.  Build it with ``python -c "import __graft_entry__ as g; g.build()"``
or ``make -C robust-segmentation_b200/csrc``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librobseg_b200.so")
ABI_VERSION = 1
MAX_ROW_JOBS = 8
CTL_ITER, CTL_NITER, CTL_EPS, CTL_SCHED, CTL_MAX_ITER = 0, 1, 2, 8, 4096

F32, BF16 = 0, 1
LOSS_CE, LOSS_MASK_CE, LOSS_MASK_CE_BAL, LOSS_JS, LOSS_ARGMAX = 0, 1, 2, 3, 4

_p, _i, _i64, _f, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t


class RowJob(C.Structure):
    _fields_ = [("dst", _p), ("src", _p), ("flags", _p), ("unless", _p), ("row_bytes", _i64)]


# name -> (restype, argtypes); one entry per symbol of include/robseg_b200.h
SIGNATURES = {
    "robseg_version": (_i, []),
    "robseg_last_error": (C.c_char_p, []),
    "robseg_profile_next_kernel": (_i, [_p, _p]),
    "robseg_loss_workspace_bytes": (_sz, [_i, _i, _i64, _i]),
    "robseg_loss_fwd_bwd": (_i, [_p, _i, _p, _p, _i, _i, _i, _i, _i64, _p, _p, _p, _p, _p, _p,
                                 _p, _p, _p, _p, _sz, _p]),
    "robseg_loss_fwd_bwd_counts": (_i, [_p, _i, _p, _p, _i, _i, _i, _i, _i64, _p, _p, _p, _p, _p, _p,
                                        _p, _p, _p, _p, _p, _sz, _p]),
    "robseg_loss_upsampled_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "robseg_loss_upsampled_fwd_bwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p,
                                           _p, _p, _sz, _p]),
    "robseg_loss_upsampled_fwd_bwd_counts": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p,
                                                  _p, _p, _p, _p, _sz, _p]),
    "robseg_apgd_step": (_i, [_p, _p, _p, _p, _p, _f, _f, _f, _i, _i64, _p, _p]),
    "robseg_apgd_step_fused": (_i, [_p, _p, _p, _p, _p, _f, _f, _f, _i, _i64, _p, _p, _p, _p, _p, _p]),
    "robseg_apgd_step_ctl": (_i, [_p, _p, _p, _p, _p, _p, _i, _i64, _p, _p, _p, _p, _p]),
    "robseg_apgd_bookkeep_ctl": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i64, _i, _p, _p, _p]),
    "robseg_project_linf": (_i, [_p, _p, _p, _f, _i64, _p, _p]),
    "robseg_pgd_step": (_i, [_p, _p, _p, _f, _f, _i, _i, _i64, _p, _p]),
    "robseg_apgd_bookkeep": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i64, _i,
                                  _p, _p, _p, _p]),
    "robseg_row_select": (_i, [C.POINTER(RowJob), _i, _i, _p]),
    "robseg_pixel_hist": (_i, [_p, _p, _i, _i, _i64, _i, _i, _p, _p, _p, _p, _p, _p]),
    "robseg_sea_worst_acc": (_i, [_p, _p, _i, _i, _i, _p, _p, _p]),
    "robseg_upsample_bilinear_fwd": (_i, [_p, _i64, _i, _i, _p, _i, _i, _p]),
    "robseg_upsample_bilinear_bwd": (_i, [_p, _i64, _i, _i, _p, _i, _i, _p]),
    "robseg_upsample_bilinear_bwd_strided": (_i, [_p, _i64, _i, _i64, _i64, _i, _i, _p, _i, _i, _p]),
    "robseg_exact_mean_host": (_i, [_p, _i64, _p]),
    "robseg_sea_greedy_round_host": (_i, [_p, _p, _i, _i, _i, _p, _p, _p, _p, _p]),
}

_lib = None
launches = 0  # kernels launched through this binding (bench.py reports it)


def load():
    """Load (once) and return the ctypes handle; RuntimeError if it cannot be loaded."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA extension has not been built "
            "(run __graft_entry__.build()); robseg-b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError -> missing export
        fn.restype, fn.argtypes = res, args
    if lib.robseg_version() != ABI_VERSION:
        raise RuntimeError(f"librobseg_b200.so ABI {lib.robseg_version()} != {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().robseg_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def count(n):
    global launches
    launches += n
