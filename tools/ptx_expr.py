#!/usr/bin/env python
"""Tiny PTX symbolic evaluator: prints the float expression tree behind every
st.global in a kernel (linear pass; PTX virtual registers are near-SSA).

Used to pin the ROUNDED OPERATION ORDER (which mul/add pairs nvcc contracted
into fma) of the reference kernels, so the sm_100a kernels can reproduce it
with explicit __fmaf_rn/__fmul_rn/__fadd_rn and stay bit-exact.

usage: ptx_expr.py file.ptx kernel_substring [max_len]
"""
import re, sys

IDX = None
def short(e):
    e = re.sub(r"add_s64\(mul_s64\(add_s64\(mul_s64\(add_s64\(mul_s32\(%ctaid.z, %nctaid.y\), %ctaid.y\), %nctaid.x\), %ctaid.x\), mul_s32\(mul_s32\(%ntid.x, %ntid.y\), %ntid.z\)\), mad_s32\(mad_s32\(%tid.z, %ntid.y, %tid.y\), %ntid.x, %tid.x\)\)", "IDX", e)
    return e

def main():
    path, kname = sys.argv[1], sys.argv[2]
    maxlen = int(sys.argv[3]) if len(sys.argv) > 3 else 4000
    txt = open(path).read()
    m = re.search(r"\.entry\s+(\S*%s\S*)\(" % re.escape(kname), txt)
    start = m.start()
    body = txt[txt.index("{", start):]
    env = {}
    nload = [0]
    def val(x):
        x = x.strip()
        if x.startswith("0f"):
            import struct
            return repr(struct.unpack(">f", bytes.fromhex(x[2:]))[0])
        if x.startswith("0d"):
            import struct
            return repr(struct.unpack(">d", bytes.fromhex(x[2:]))[0]) + "d"
        return env.get(x, x)
    for line in body.splitlines():
        line = line.strip()
        if line.startswith("}") and not line.startswith("};"):
            pass
        pm = re.match(r"(@!?%p\d+\s+)?([a-z0-9_.:]+)\s+(.*);", line)
        if not pm:
            if line.startswith("$L") or line.startswith("BB"):
                print("---", line)
            continue
        pred, op, rest = pm.group(1) or "", pm.group(2), pm.group(3)
        args = [a.strip() for a in re.split(r",\s*(?![^\[]*\])", rest)]
        base = op.split(".")[0]
        if op.startswith("ld.param"):
            env[args[0]] = "p" + re.sub(r".*_param_(\d+)\]", r"\1", args[1])
        elif op.startswith("ld."):
            addr = args[1].strip("[]")
            am = re.match(r"(%\w+)(\+(-?\d+))?", addr)
            a0 = val(am.group(1)); off = am.group(3) or "0"
            dsts = args[0].strip("{}").split(",")
            for i, d in enumerate(dsts):
                env[d.strip()] = "LD(%s+%s#%d)" % (a0, off, i)
        elif op.startswith("st.global") or op.startswith("st.shared"):
            addr = args[0]
            am = re.match(r"\[(%\w+)(\+(-?\d+))?\]", addr)
            a0 = val(am.group(1)); off = am.group(3) or "0"
            srcs = [s.strip() for s in ",".join(args[1:]).strip("{}").split(",")]
            for i, s in enumerate(srcs):
                e = val(s)
                print("%sST %s [%s+%s#%d] = %s" % (pred, op, short(a0)[:60], off, i, short(e)[:maxlen]))
        elif base in ("fma", "mul", "add", "sub", "div", "max", "min", "mad"):
            env[args[0]] = "%s(%s)" % (base + ("64" if "f64" in op else ("" if "f32" in op else "_" + op.split(".")[-1])), ", ".join(val(a) for a in args[1:]))
        elif base in ("neg", "abs", "sqrt", "rcp", "ex2", "lg2", "rsqrt", "sin", "cos"):
            env[args[0]] = "%s(%s)" % (base, val(args[1]))
        elif base in ("cvt", "cvta", "mov"):
            tag = op if base == "cvt" and ("f32" in op or "f64" in op) else None
            env[args[0]] = ("%s(%s)" % (op, val(args[1]))) if tag else val(args[1])
        elif base in ("setp", "selp", "and", "or", "xor", "not", "shl", "shr", "bra", "ret", "bar", "call", "cvt", "mul24", "rem", "popc", "clz", "bfe", "exit", "shfl", "vote"):
            if base == "selp":
                env[args[0]] = "selp(%s, %s, %s)" % (val(args[1]), val(args[2]), val(args[3]))
            elif base == "setp":
                env[args[0]] = "setp.%s(%s)" % (op, ", ".join(val(a) for a in args[1:]))
                print("SETP %s%s = %s" % (pred, args[0], short(env[args[0]])[:maxlen]))
            elif base in ("bra", "ret", "exit"):
                print("CTRL", pred, op, rest)
            else:
                env[args[0]] = "%s(%s)" % (op, ", ".join(val(a) for a in args[1:]))
        else:
            if args:
                env[args[0]] = "%s(%s)" % (op, ", ".join(val(a) for a in args[1:]))
if __name__ == "__main__":
    main()
