"""Summarise an ncu report (CPU side): python scripts/ncu_summary.py report.ncu-rep [pattern ...]"""
import csv, subprocess, sys
rep = sys.argv[1]
pats = sys.argv[2:] or ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct",
    "gpu__dram_throughput", "sm__throughput.avg.pct", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct",
    "sm__warps_active.avg.pct", "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "warp_issue_stalled",
    "lts__t_bytes.sum ", "sm__cycles_elapsed.avg ", "pipe_xu", "inst_executed_pipe_lsu", "inst_executed_pipe_alu", "inst_executed_pipe_fma",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared", "smsp__inst_executed.sum", "lts__t_sectors_op_write", "lts__t_sectors_op_read", "achieved_occupancy",
    "l1tex__data_bank_conflicts", "smsp__cycles_active.avg", "sm__cycles_active.avg", "dram__cycles_active"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:90], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
    for i, h in enumerate(hdr):
        if any(p.strip() in h for p in pats) and r[i] not in ("", "0", "n/a"):
            print(f"  {h[:100]:100s} {r[i]:>16s} {units[i]}")
