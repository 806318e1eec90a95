#!/usr/bin/env python
"""A dense product must give every row the same bits whatever the number of rows in the call (row blocks of a
multi-GPU run see 1/P of the rows): run the fused highway kernel and a plain product on all rows and on row prefixes
and compare bit patterns.  `python tools/gemm_rows_check.py [gemm_v]`"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from geographconv_b200 import capi  # noqa: E402

n, hd, ld, cn = 500000, 300, 320, 256
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev)
ctx = capi.Context(0, C.c_void_p(stream.cuda_stream))
if len(sys.argv) > 1:
    ctx.set_option("gemm_v", int(sys.argv[1]))
g = torch.Generator(device="cuda").manual_seed(1)
rnd = lambda r, c, s=1.0: torch.randn(r, c, device=dev, generator=g) * s
S, X = rnd(n, ld), torch.tanh(rnd(n, ld))
S[:, hd:] = 7.0   # padding columns hold junk on purpose: nothing may read them
X[:, hd:] = -3.0
Wh, Wt, Wout = rnd(hd, ld, 0.05), rnd(hd, ld, 0.05), rnd(hd, cn, 0.05)
bh, bt = rnd(1, ld).view(-1), rnd(1, ld).view(-1)
ws = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
ctx.call("gcnb_set_workspace", C.c_void_p(ws.data_ptr()), ws.numel())
p = lambda t: C.c_void_p(t.data_ptr())


def highway(rows, off=0):
    Y, H, T = (torch.zeros(rows, ld, device=dev) for _ in range(3))
    stream.wait_stream(torch.cuda.current_stream())
    o = off * ld * 4
    ctx.call("gcnb_highway_fwd_f32", rows, hd, C.c_void_p(S.data_ptr() + o), ld, C.c_void_p(X.data_ptr() + o), ld, p(Wh), ld,
             p(bh), p(Wt), ld, p(bt), 1, p(Y), ld, p(H), ld, p(T), ld)
    ctx.sync()
    return Y, H, T


def plain(rows, off=0):
    Q = torch.zeros(rows, cn, device=dev)
    stream.wait_stream(torch.cuda.current_stream())
    ctx.call("gcnb_gemm_f32", 0, 0, rows, cn, hd, C.c_void_p(X.data_ptr() + off * ld * 4), ld, p(Wout), cn, p(Q), cn, 0, None, 0)
    ctx.sync()
    return (Q,)


ok = True
for name, fn in (("highway", highway), ("plain N=256", plain)):
    full = fn(n)
    again = fn(n)
    same = all(torch.equal(a.view(torch.int32), b.view(torch.int32)) for a, b in zip(full, again))
    print("%-12s all rows twice: %s" % (name, "bit-identical" if same else "DIFFERENT"))
    ok &= same
    for rows, off in ((250000, 0), (250000, 250000), (62500, 0), (62500, 437500), (1000, 0), (128, 499872)):
        part = fn(rows, off)
        bad = sum(int((a[off:off + rows].view(torch.int32) != b.view(torch.int32)).sum()) for a, b in zip(full, part))
        print("%-12s rows [%d, %d) alone vs inside the full call: %d differing elements" % (name, off, off + rows, bad))
        ok &= bad == 0
print("GEMM_ROWS_CHECK", "PASS" if ok else "FAIL")
