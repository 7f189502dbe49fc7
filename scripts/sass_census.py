"""SASS mnemonic census of librobseg_b200.so (CPU side, cuobjdump): which kernels carry TMA / mbarrier /
MUFU / vector-store instructions.  python scripts/sass_census.py > profiles/r02_sass_census.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "robust-segmentation_b200", "librobseg_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()  # noqa: E731
WATCH = ["UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDGSTS", "MUFU.EX2", "MUFU.LG2", "FMNMX3", "HMNMX2", "HMUL2", "STG.E.EF", "STG.E",
         "LDS.128", "LDS.64", "ATOMS", "REDG", "RED.E", "SHFL", "MATCH", "UTC", "HMMA", "LDTM"]
counts, order, total = collections.defaultdict(collections.Counter), [], collections.Counter()
fn = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        order.append(fn)
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if fn and m:
        op = m.group(1)
        total[fn] += 1
        for w in WATCH:
            if op.startswith(w):
                counts[fn][w] += 1
                break
print("# SASS census of librobseg_b200.so (sm_100a), `cuobjdump -sass`, round 2\n")
print("Static instruction counts per kernel for the mnemonics that identify the design: `UTMALDG` = TMA tensor load "
      "(`cp.async.bulk.tensor`), `SYNCS` = mbarrier, `LDGSTS` = `cp.async`, `MUFU.EX2` = `ex2.approx`, `FMNMX3` = 3-input max, "
      "`STG.E.EF` = streaming (`st.global.cs`) stores, `HMUL2` / `HMNMX2` = the packed bf16 multiply / max of the bf16 loss kernel, `ATOMS` = shared-memory atomics, `REDG` = global reductions without return value (the per-image class counters of the loss kernels).  No `UTMASTG`: the gradient is "
      "written from registers with `st.global.cs` (one 8-16 B store per lane and channel: already full coalesced lines, a TMA "
      "store would need the gradient staged in shared memory the kernel has no room for).  No tensor-core instructions "
      "(`UTC*MMA`, `HMMA`) anywhere: the path has no dense contraction.\n")
print("| kernel | SASS instr | " + " | ".join(WATCH[:16]) + " |")
print("|---|---|" + "---|" * 16)
for fn in order:
    name = demangle(fn)
    name = re.sub(r"robseg::", "", name)
    name = re.sub(r"\(.*$", "", name)
    print(f"| `{name}` | {total[fn]} | " + " | ".join(str(counts[fn].get(w, "")) for w in WATCH[:16]) + " |")
tc = sum(counts[f][w] for f in order for w in ("UTC", "HMMA", "LDTM"))
print(f"\ntensor-core / TMEM instructions in the library: {tc}")
# excerpt: the producer warp's TMA issue of the SEA loss kernel (mbarrier expect-tx, UTMALDG) and a consumer wait
target = next((f for f in order if demangle(f).startswith("void robseg::loss_tma_kernel<float, 2, 1, 0>")), None)
if target:
    body, on = [], False
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            on = m.group(1) == target
            continue
        if on and re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", line):
            body.append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", line.rstrip()))
    idx = [i for i, l in enumerate(body) if "UTMALDG" in l]
    print(f"\n## Excerpt: `{demangle(target).split('(')[0]}` ({len(body)} SASS instructions)\n")
    print("The producer warp arms the stage's mbarrier with the expected byte count (`SYNCS.ARRIVE.TRANS64` after `SYNCS.EXCH` "
          "initialised it) and issues ONE 3-D tensor load per stage (`UTMALDG.3D`: a [C x 64-pixel] box of one image, issued by an "
          "elected lane); `SYNCS.PHASECHK.TRANS64.TRYWAIT` is the mbarrier wait -- the producer's for a released stage before it "
          "re-fills it (the first one below), the consumers' for the stage's bytes before their first shared-memory read.\n")
    print("```")
    for i in idx[:1]:
        print("\n".join(body[max(0, i - 10):i + 4]))
    waits = [i for i, l in enumerate(body) if "TRYWAIT" in l]
    if waits:
        print("        ...")
        print("\n".join(body[max(0, waits[0] - 2):waits[0] + 3]))
    print("```")
