#!/usr/bin/env python
"""Static SASS opcode histogram of one kernel of libndzip_b200.so, grouped by issue pipe.

    python scripts/sass_hist.py <kernel-name-regex> [--range 0xa10:0x4560] [--so path] [--dump file]

alu pipe: LOP3 SHF PRMT IADD3 VIADD LEA SEL ISETP ... ; fma pipe: IMAD* (B300_MICROARCH.md "Pipe rates");
each pipe issues one warp instruction per two cycles per scheduler, so max(2*alu, 2*fma, total) bounds a loop body.
"""
import argparse
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ALU = ("LOP3", "SHF", "PRMT", "IADD3", "VIADD", "LEA", "SEL", "ISETP", "PLOP3", "R2P", "P2R", "MOV", "FLO", "POPC", "IABS", "VIMNMX", "IMNMX", "BREV", "SGXT", "BMSK")
FMA = ("IMAD", "FFMA", "FMUL", "FADD")
LSU = ("LDS", "STS", "LDG", "STG", "LDL", "STL", "ATOM", "RED", "LDC", "LDCU", "SYNCS", "UTMA", "UBLKCP", "LDSM", "STSM")


def pipe_of(op):
    base = op.split(".")[0]
    if base in FMA:
        return "fma"
    if base in ALU:
        return "alu"
    if base.startswith(LSU):
        return "lsu"
    if base.startswith("U") and base not in ("UTMALDG", "UBLKCP"):
        return "uniform"
    return "other"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("kernel")
    ap.add_argument("--so", default=os.path.join(ROOT, "ndzip_b200", "libndzip_b200.so"))
    ap.add_argument("--range", default=None, help="hex address range lo:hi inside the kernel")
    ap.add_argument("--dump", default=None)
    ap.add_argument("--top", type=int, default=25)
    args = ap.parse_args()
    sass = subprocess.run(["cuobjdump", "-sass", args.so], capture_output=True, text=True, check=True).stdout
    cur, rows = None, []
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur is None or not re.search(args.kernel, cur):
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            rows.append((cur, int(m.group(1), 16), m.group(2).strip()))
    names = sorted({r[0] for r in rows})
    if len(names) != 1:
        raise SystemExit("kernel regex matches %d functions:\n  %s" % (len(names), "\n  ".join(names)))
    lo, hi = 0, 1 << 30
    if args.range:
        a, b = args.range.split(":")
        lo, hi = int(a, 16), int(b, 16)
    sel = [(a, t) for _, a, t in rows if lo <= a <= hi]
    if args.dump:
        with open(args.dump, "w") as f:
            for a, t in sel:
                f.write("%04x %s\n" % (a, t))
    ops, pipes = collections.Counter(), collections.Counter()
    for _, t in sel:
        tok = t.split()
        op = tok[1] if tok[0].startswith("@") else tok[0]
        ops[op] += 1
        pipes[pipe_of(op)] += 1
    print(names[0])
    print("instructions:", len(sel), dict(pipes))
    for op, n in ops.most_common(args.top):
        print("%5d %-28s %s" % (n, op, pipe_of(op)))


if __name__ == "__main__":
    main()
