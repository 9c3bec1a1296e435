#!/usr/bin/env python
"""Per-source-line stall samples / executed instructions of one kernel:  python tools/ncu_lines.py rep.ncu-rep <kernel regex> [top]"""
import csv, subprocess, sys, io
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None
out = []
for r in rows:
    if len(r) > 6 and r[0] == "Line No" and "# Samples" in r:
        hdr = r; js = hdr.index("# Samples"); je = hdr.index("Instructions Executed"); continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        try: out.append((int(r[js]), int(r[je]), int(r[0]), r[1][:130]))
        except ValueError: pass
tot = sum(o[0] for o in out); tote = sum(o[1] for o in out)
print("total samples %d, warp instructions %d" % (tot, tote))
for s, e, ln, src in sorted(out, reverse=True)[:top]:
    print("%6d (%4.1f%%) %9d  L%-4d %s" % (s, 100.0 * s / max(tot, 1), e, ln, src))
