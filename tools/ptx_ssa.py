#!/usr/bin/env python
"""Print a kernel's floating-point dataflow from PTX as a compact SSA listing
(one line per float op, loads shown with base parameter + byte offset).
Companion of ptx_expr.py; used to read off the reference's rounded op order.

usage: ptx_ssa.py file.ptx kernel_substring
"""
import re, struct, sys

def lit(x):
    if x.startswith("0f"):
        return repr(struct.unpack(">f", bytes.fromhex(x[2:]))[0]) + "f"
    if x.startswith("0d"):
        return repr(struct.unpack(">d", bytes.fromhex(x[2:]))[0]) + "d"
    return x

def main():
    path, kname = sys.argv[1], sys.argv[2]
    txt = open(path).read()
    m = re.search(r"\.entry\s+(\S*%s\S*)\(" % re.escape(kname), txt)
    body = txt[txt.index("{", m.start()):]
    end = body.index("\n}")
    body = body[:end]
    base = {}   # %rd -> (param, note)
    for line in body.splitlines():
        line = line.strip()
        if not line or line.startswith("//") or line.startswith("."):
            continue
        if re.match(r"^\$?L?_*BB\w+:", line) or line.endswith(":"):
            print(line); continue
        pm = re.match(r"(@!?%p\d+\s+)?([a-z0-9_.:]+)\s+(.*);", line)
        if not pm:
            continue
        pred, op, rest = pm.group(1) or "", pm.group(2), pm.group(3)
        args = [a.strip() for a in re.split(r",\s*(?![^\[{]*[\]}])", rest)]
        if op.startswith("ld.param"):
            base[args[0]] = "p" + re.sub(r".*_param_(\d+)\]", r"\1", args[1])
            if ".f32" in op:
                print("%s%s = param %s" % (pred, args[0], base[args[0]]))
            continue
        if op.startswith("cvta") or (op.startswith("mov") and args[1] in base):
            base[args[0]] = base.get(args[1], args[1]); continue
        if op.startswith("add.s64") or op.startswith("mul.wide") or op.startswith("shl.b64") or op.startswith("mad.wide"):
            srcs = [base.get(a) for a in args[1:] if a in base]
            if srcs:
                base[args[0]] = srcs[0] + "+i"
            continue
        isf = (".f32" in op or ".f64" in op)
        if op.startswith("ld.") and isf:
            addr = args[1].strip("[]")
            am = re.match(r"(%\w+)(\+(-?\d+))?", addr)
            print("%s%s = LD %s[%s+%s]" % (pred, args[0], op.split(".")[-1], base.get(am.group(1), am.group(1)), am.group(3) or "0"))
        elif op.startswith("st.") :
            addr = args[0].strip("[]")
            am = re.match(r"(%?\w+)(\+(-?\d+))?", addr)
            b = am.group(1) if am else addr
            print("%sST %s [%s+%s] <- %s" % (pred, op, base.get(b, b), (am.group(3) if am else None) or "0", ", ".join(lit(a) for a in args[1:])))
        elif isf or op.startswith("setp") or op.startswith("selp") or op.startswith("bra") or op.startswith("cvt"):
            if op.startswith("bra"):
                print("%sbra %s" % (pred, rest)); continue
            print("%s%s = %s(%s)" % (pred, args[0], op, ", ".join(lit(a) for a in args[1:])))

if __name__ == "__main__":
    main()
