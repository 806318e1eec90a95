#!/usr/bin/env python
"""Every dense product shape of one C3 training step, timed alone through the C ABI for both tcgen05 GEMM kernels
(ctx option gemm_v: 2 = two co-resident CTAs per SM, 1 = one tile per SM) and checked against each other.

    python tools/gemm_bench.py [n_rows]        (default 500000)
"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from geographconv_b200 import capi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
hd, ld, kh, cn, ldcn = 300, 320, 1024, 256, 256
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev)
ctx = capi.Context(0, C.c_void_p(stream.cuda_stream))
g = torch.Generator(device="cuda").manual_seed(0)
rnd = lambda r, c, s=1.0: torch.randn(r, c, device=dev, generator=g) * s
S, X, V = rnd(n, ld), rnd(n, ld), rnd(n, ld)
S[:, hd:] = 0; X[:, hd:] = 0; V[:, hd:] = 0
Xh = rnd(n, kh)
Wh, Wt = rnd(hd, ld, 0.05), rnd(hd, ld, 0.05)
W0h = rnd(kh, ld, 0.05)
Wout = rnd(hd, ldcn, 0.05)
bh, bt = rnd(1, ld).view(-1), rnd(1, ld).view(-1)
Y, H, T, Cm, Q = (torch.zeros(n, ld, device=dev) for _ in range(5))
ws = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ctx.call("gcnb_set_workspace", C.c_void_p(ws.data_ptr()), ws.numel())
stream.wait_stream(torch.cuda.current_stream())
p = lambda t: C.c_void_p(t.data_ptr())

CASES = {
    "highway fused (S.Wh, X.Wt, gate mix)": (lambda: ctx.call(
        "gcnb_highway_fwd_f32", n, hd, p(S), ld, p(X), ld, p(Wh), ld, p(bh), p(Wt), ld, p(bt), 1, p(Y), ld, p(H), ld,
        p(T), ld), 2 * 2.0 * n * hd * hd, (Y, H, T)),
    "plain x.Wout (N=256)": (lambda: ctx.call(
        "gcnb_gemm_f32", 0, 0, n, cn, hd, p(X), ld, p(Wout), ldcn, p(Q), ld, 0, None, 0), 2.0 * n * hd * cn, (Q,)),
    "hot block X_hot.W0 (K=1024)": (lambda: ctx.call(
        "gcnb_gemm_f32", 0, 0, n, hd, kh, p(Xh), kh, p(W0h), ld, p(Cm), ld, 0, None, 0), 2.0 * n * kh * hd, (Cm,)),
    "dgrad accumulate dx += V.Wh^T": (lambda: ctx.call(
        "gcnb_gemm_f32", 0, 1, n, hd, hd, p(V), ld, p(Wh), ld, p(Cm), ld, 1, None, 0), 2.0 * n * hd * hd, (Cm,)),
    "dgrad pair dx += dT.Wt^T + V.Wh^T": (lambda: ctx.call(
        "gcnb_gemm_pair_f32", 1, n, hd, hd, p(S), ld, p(Wt), ld, p(V), ld, p(Wh), ld, p(Cm), ld, 1),
        2 * 2.0 * n * hd * hd, (Cm,)),
    "dgrad G.Wout^T (K=256)": (lambda: ctx.call(
        "gcnb_gemm_f32", 0, 1, n, hd, cn, p(Q), ld, p(Wout), ldcn, p(Y), ld, 0, None, 0), 2.0 * n * hd * cn, (Y,)),
    "wgrad x^T.V": (lambda: ctx.call(
        "gcnb_gemm_f32", 1, 0, hd, hd, n, p(X), ld, p(V), ld, p(T), ld, 0, None, 0), 2.0 * n * hd * hd, ()),
}

results = {}
for name, (fn, flops, outs) in CASES.items():
    row = []
    ref = None
    for v in (1, 2):
        ctx.set_option("gemm_v", v)
        if "accumulate" in name or "pair" in name:
            Cm.zero_()
            stream.wait_stream(torch.cuda.current_stream())
        for _ in range(2):
            fn()
        ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        row.append(ms)
        if outs and "accumulate" not in name and "pair" not in name:
            snap = [o[:4096].clone() for o in outs]
            if ref is None:
                ref = snap
            else:
                err = max(float((a - b).abs().max() / (b.abs().max() + 1e-30)) for a, b in zip(snap, ref))
                row.append(err)
    extra = ("  max rel diff v2 vs v1 %.1e" % row[2]) if len(row) > 2 else ""
    print("%-38s v1 %7.3f ms (%6.1f TF/s fp32-equivalent)   v2 %7.3f ms (%6.1f TF/s)%s"
          % (name, row[0], flops / row[0] / 1e9, row[1], flops / row[1] / 1e9, extra), flush=True)
