"""Write profiles/roofline_traffic.json from ncu --set full captures of the loss kernel (CPU side):

    python scripts/update_traffic.py sea_c150=gpurun_out/r2v_loss_c150.ncu-rep sea_c151=gpurun_out/r2v_loss_c151.ncu-rep

Each value is dram__bytes_read.sum + dram__bytes_write.sum of the (single) kernel in the report, per launch.  The
file also records the sha of the kernel sources of THIS tree (bench.kernel_source_sha), so bench.py reports a
capture taken on other sources as stale (`traffic: null`) instead of passing it on as current."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def dram_bytes(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, row = rows[0], rows[1], rows[2]
    tot = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(m)
        tot += float(row[i].replace(",", "")) * UNIT[units[i]]
    return int(round(tot)), row[hdr.index("Kernel Name")]


def main():
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    with open(path) as f:
        d = json.load(f)
    d.setdefault("_reports", {})
    for arg in sys.argv[1:]:
        key, rep = arg.split("=", 1)
        d[key], kern = dram_bytes(rep)
        d["_reports"][key] = os.path.relpath(rep, ROOT)
        print(key, d[key], kern[:80])
    d["_kernel_source_sha"] = bench.kernel_source_sha()
    with open(path, "w") as f:
        json.dump(d, f, indent=1)
        f.write("\n")


if __name__ == "__main__":
    main()
