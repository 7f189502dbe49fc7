"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) as a markdown share table.
  python scripts/launch_list_md.py launches.csv "title / command" > profiles/xx.md"""
import collections
import csv
import re
import sys

path, title = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.defaultdict(float)
cnt = collections.Counter()
for r in rows[1:]:
    ns = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r[ui], 1e-3)
    name = re.sub(r"\(.*", "", r[ki])[:100]
    tot[name] += ns
    cnt[name] += 1
total = sum(tot.values())
print(f"# {title}\n")
print(f"total {total/1e3:.1f} ms over {sum(cnt.values())} launches (per-launch times under ncu are cold-cache and serialised: read SHARES)\n")
print("| share | time (us) | launches | kernel |\n|---|---|---|---|")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:16]:
    print(f"| {100*v/total:.2f}% | {v:.0f} | {cnt[k]} | `{k}` |")
ours = {k: v for k, v in tot.items() if "robseg" in k}
print("\n## robseg-b200 kernels in the same window\n")
print("| share | time (us) | launches | avg (us) | kernel |\n|---|---|---|---|---|")
for k, v in sorted(ours.items(), key=lambda kv: -kv[1]):
    print(f"| {100*v/total:.3f}% | {v:.0f} | {cnt[k]} | {v/cnt[k]:.1f} | `{k}` |")
print(f"\nAll robseg-b200 kernels together: {100*sum(ours.values())/total:.2f}% of the GPU time in the window.")
ups = sum(v for k, v in tot.items() if "upsample_bilinear2d" in k)
print(f"ATen `upsample_bilinear2d*` kernels left in the window: {100*ups/total:.2f}%.")
