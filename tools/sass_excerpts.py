#!/usr/bin/env python
"""profiles/r02_sass_excerpts.txt: opcode census and inner-loop excerpts of libgsr_b200.so (`cuobjdump -sass`)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "gaussian-splatting_deformable_b200", "libgsr_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
OPS = ("FFMA2", "FMUL2", "FADD2", "UTCHMMA", "UTMALDG", "UBLKCP", "LDTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "LDG.E.ENL2.256", "STG.E.ENL2.256",
       "REDG", "ATOMS", "MUFU.EX2", "MUFU.RCP", "MATCH.ANY", "HMMA", "LDGSTS")


def dem(n):
    d = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    d = re.sub(r"^void ", "", d)
    depth, cut = 0, len(d)
    for i, ch in enumerate(d):                # cut the argument list, keep template arguments
        if ch == "<": depth += 1
        elif ch == ">": depth -= 1
        elif ch == "(" and depth == 0: cut = i; break
    return d[:cut]


out = ["# SASS evidence, libgsr_b200.so (sm_100a, nvcc 12.9), `cuobjdump -sass`; regenerate with tools/sass_excerpts.py", "",
       "## Per-kernel counts of Blackwell-specific / notable opcodes (static instruction counts)", ""]
table = []
for f in funcs:
    name = dem(f.split("\n", 1)[0].strip())
    ins = re.findall(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f)
    c = collections.Counter()
    for i in ins:
        for o in OPS:
            if i.startswith(o): c[o] += 1
    table.append((name, len(ins), c, f))
for name, n, c, _ in sorted(table, key=lambda t: t[0]):
    out.append("%-70s %5d instr  %s" % (name[:70], n, "  ".join("%s=%d" % kv for kv in sorted(c.items()))))


def excerpt(pattern, key, before, after, title):
    for name, n, c, f in table:
        if pattern in name:
            lines = [l for l in f.split("\n") if re.search(r"/\*[0-9a-f]{4}\*/", l)]
            idx = [i for i, l in enumerate(lines) if key in l]
            if not idx: continue
            i0, i1 = max(0, idx[0] - before), min(len(lines), idx[0] + after)
            out.extend(["", "## " + title, "# " + name, ""])
            out.extend(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/", "", l).rstrip() for l in lines[i0:i1])
            return
    out.extend(["", "## " + title, "# (not found: %s / %s)" % (pattern, key)])


excerpt("mlp_gemm_kernel<256>", "UTCHMMA", 10, 36, "deformation-network GEMM: MMA issue (tcgen05.mma kind::tf32 -> UTCHMMA, tcgen05.commit -> UTCBAR)")
excerpt("mlp_gemm_kernel<256>", "UTMALDG", 6, 14, "deformation-network GEMM: TMA producer (cp.async.bulk.tensor.2d -> UTMALDG.2D)")
excerpt("mlp_gemm_kernel<256>", "LDTM", 4, 10, "deformation-network GEMM: epilogue TMEM load (tcgen05.ld -> LDTM)")
excerpt("blend_fwd_v2_kernel<1, 7", "FFMA2", 10, 60, "blend forward inner loop (packed FP32x2)")
excerpt("blend_bwd_v2_kernel<1, 0, true, true>", "FFMA2", 10, 70, "blend backward inner loop (packed FP32x2)")
excerpt("preprocess_fwd_kernel", "LDG.E.ENL2.256", 4, 12, "preprocess forward: 256-bit SH loads")
excerpt("blend_fwd_kernel<2, true>", "UBLKCP", 6, 10, "first-generation blend with TMA staging (cp.async.bulk -> UBLKCP)")
open(os.path.join(ROOT, "profiles", "r02_sass_excerpts.txt"), "w").write("\n".join(out) + "\n")
print("wrote profiles/r02_sass_excerpts.txt (%d lines)" % len(out))
