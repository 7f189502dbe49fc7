"""robseg-b200: the attack-side hot path of Robust-Segmentation's SEA / PIR-AT attacks on
NVIDIA B200 (sm_100a).

Layout
  csrc/      hand-written CUDA kernels + the C ABI (include/robseg_b200.h) -> librobseg_b200.so
  _lib.py    ctypes binding of that ABI (fails loudly when the library is missing)
  ops.py     tensor-level wrappers + ``torch.ops.robseg.*`` custom-op registrations
  semseg/    host-side mirror of the reference call surface: attacker.py, losses.py,
             metrics.py, val.py  (same names, argument order, defaults, return tuples)
  tools/     worse_only.py (evalSEA) and the infer.py bookkeeping mirrors
  dist.py    image-sharded SEA across ranks + the single int64 all-reduce
  dropin.py  swap the mirrors into a checkout of the reference

The directory name contains a hyphen, so it is imported under the alias ``robseg_b200``
(see ``__graft_entry__.load_package``) or by putting this directory's parent on sys.path
and using importlib; inside the package only relative imports are used.
"""
__version__ = "0.1.0"
