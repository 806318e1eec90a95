#!/usr/bin/env python
"""SASS evidence for profiles/: per kernel of libgcnb200.so, the counts of the instructions that prove what the kernel is
(tcgen05 MMA = UTCHMMA / UTCQMMA, TMEM loads = LDTM, TMA tensor loads / stores = UTMALDG / UTMASTG, bulk copies = UBLKCP,
128-bit global gathers = LDG.E.128, system-scope release / acquire of the peer barrier) plus the first lines holding each.

    python tools/sass_excerpt.py > profiles/r2_sass_excerpts.txt      (runs here: cuobjdump needs no GPU)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "geographconv_b200", "libgcnb200.so")
KERNELS = ["gemm_tc2_kernel", "gemm_tc_kernel", "wgrad_tc_kernel", "spmm_panel_kernel", "spmm_bulk_kernel", "spmm_ldg_kernel",
           "slice_push_kernel", "peer_barrier_kernel", "highway_bwd_colsum_kernel", "xent_grad_dense_kernel",
           "adj_hub_setbits", "expand_u16_kernel"]
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDG.E.128", "LDG.E.CONSTANT.128",
        "STG.E.128", "ST.E.128", "STG.E.EF.128", "RED.E", "ATOMG", "MATCH.ANY", "MEMBAR", "LDG.E.STRONG.SYS", "STG.E.STRONG.SYS",
        "FFMA", "MUFU"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)[1:]
    print("# cuobjdump -sass geographconv_b200/libgcnb200.so (sm_100a) -- instruction evidence per kernel\n")
    for body in funcs:
        name = body.split("\n", 1)[0].strip()
        if not any(k in name for k in KERNELS):
            continue
        dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
        lines = [l for l in body.split("\n") if re.search(r"/\*[0-9a-f]{4,6}\*/", l)]
        ops = collections.Counter()
        first = {}
        for l in lines:
            m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
            if not m:
                continue
            op = m.group(1)
            for k in KEYS:
                if op.startswith(k) or (k in op and k in ("STRONG.SYS",)):
                    ops[k] += 1
                    first.setdefault(k, l.strip()[:150])
        print("## %s" % dem[:160])
        print("instructions: %d;  " % len(lines) + ", ".join("%s x%d" % (k, ops[k]) for k in KEYS if ops[k]))
        for k in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "LDG.E.128", "LDG.E.CONSTANT.128", "STG.E.STRONG.SYS",
                  "LDG.E.STRONG.SYS", "MATCH.ANY"):
            if k in first:
                print("    " + first[k])
        print()


if __name__ == "__main__":
    main()
