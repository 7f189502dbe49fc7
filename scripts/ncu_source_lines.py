"""Executed warp instructions per source line of an ncu report captured with --import-source on:
    python scripts/ncu_source_lines.py report.ncu-rep [top_n]"""
import csv, subprocess, sys
rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, hdr, agg = None, None, []
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur, hdr = r[1], None
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit():
        d = dict(zip(hdr, r))
        try:
            agg.append((int(d["Instructions Executed"]), cur.split("/")[-1], int(r[0]), r[1][:120], d.get("# Samples", "")))
        except (KeyError, ValueError):
            pass
tot = sum(a[0] for a in agg)
print("total warp instructions", tot)
for n, f, l, src, smp in sorted(agg, reverse=True)[:top]:
    print(f"{100 * n / tot:5.1f}% {n:>11d} samples {smp:>6s} {f}:{l}  {src}")
