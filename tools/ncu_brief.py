#!/usr/bin/env python
"""Brief per-kernel table (time, occupancy, issue, top stalls) of an .ncu-rep:  python tools/ncu_brief.py rep.ncu-rep"""
import csv, subprocess, sys, io
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
def col(r, name):
    try: return r[hdr.index(name)]
    except ValueError: return ""
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    print("== %s  grid %s block %s" % (col(r, "Kernel Name")[:70], col(r, "launch__grid_size"), col(r, "launch__block_size")))
    for k in ("gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
              "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
              "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
              "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum"):
        v = col(r, k)
        if v: print("   %-62s %s" % (k, v))
    st = sorted(((float(col(r, s).replace(",", "") or 0)), s) for s in stall)[::-1][:6]
    print("   stalls/issue: " + ", ".join("%s %.2f" % (s.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v) for v, s in st))
