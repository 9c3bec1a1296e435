#!/usr/bin/env python
"""Pretty-print the per-kernel part of a bench.py JSON line (stdin or file)."""
import json
import sys
d = json.loads((open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin).read().strip().splitlines()[-1])
v = d["config"]["views_per_gpu_per_step"]
print("ms_per_view %.4f  value %.1f M/s  e2e %.1f M/s  launches %s" % (d["ms_per_view"], d["value"] / 1e6, d["e2e"]["value"] / 1e6, d.get("gpu_launches")))
tot = 0.0
for k, x in (d.get("kernels") or {}).items():
    per_view = x["total_ms"] / v
    if k != "binning_stage":          # a derived entry (sum of the binning kernels), not a kernel
        tot += per_view
    print("  %-28s x%-3d avg %.4f ms  per-view %.4f ms  %s" % (k, x["launches"] // v, x["avg_ms"], per_view,
          ("%.0f GB/s (%.0f%%)" % (x["alg_GBps"], 100 * x["frac_of_hbm_peak"])) if "alg_GBps" in x else ""))
print("  kernel sum per view %.4f ms" % tot)
print("  roofline", d.get("roofline"))
print("  clocks", d.get("clocks"))
