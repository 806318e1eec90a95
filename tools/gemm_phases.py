#!/usr/bin/env python
"""Phase timing of the fused highway tcgen05 kernel: per-CTA clock64 stamps (debug hook gcnb_debug_set_tc_buffer)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from geographconv_b200 import capi  # noqa: E402

n, hd, ld = 500000, 300, 320
version = int(sys.argv[1]) if len(sys.argv) > 1 else 1   # tcgen05 GEMM kernel (ctx option gemm_v)
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev)
ctx = capi.Context(0, C.c_void_p(stream.cuda_stream))
ctx.lib.gcnb_debug_set_tc_buffer.argtypes = [C.c_void_p, C.c_void_p]
ctx.set_option("gemm_v", version)
g = torch.Generator(device="cuda").manual_seed(0)
S = torch.randn(n, ld, device=dev, generator=g); X = torch.randn(n, ld, device=dev, generator=g)
Wh = torch.randn(hd, ld, device=dev, generator=g) * 0.05; Wt = torch.randn(hd, ld, device=dev, generator=g) * 0.05
bh = torch.zeros(ld, device=dev); bt = torch.zeros(ld, device=dev)
Y, H, T = (torch.zeros(n, ld, device=dev) for _ in range(3))
ws = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
ctx.call("gcnb_set_workspace", C.c_void_p(ws.data_ptr()), ws.numel())
n_cta = ((n + 127) // 128) * (3 if version >= 2 else 2)
dbg = torch.zeros(n_cta * 16, dtype=torch.int64, device=dev)
stream.wait_stream(torch.cuda.current_stream())
p = lambda t: C.c_void_p(t.data_ptr())


def run():
    ctx.call("gcnb_highway_fwd_f32", n, hd, p(S), ld, p(X), ld, p(Wh), ld, p(bh), p(Wt), ld, p(bt), 1, p(Y), ld, p(H), ld,
             p(T), ld)


for _ in range(3):
    run()
ctx.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(5):
    run()
e1.record(stream)
torch.cuda.synchronize()
print("highway kernel (gemm_v %d): %.3f ms per launch" % (version, e0.elapsed_time(e1) / 5))
ctx.lib.gcnb_debug_set_tc_buffer(ctx.h, p(dbg))
for mode in [1, 2, 3, 4, 7]:
    ctx.set_option("tc_dbg_mode", mode)
    run()
    ctx.sync()
    dd = dbg.cpu().numpy().reshape(-1, 16)
    print("dbg_mode %d (1 = identity activations, 2 = no stores, 4 = no bias loads): epilogue median %.0f clk, math+staging %.0f"
          % (mode, np.median(dd[:, 5] - dd[:, 4]), np.median(dd[:, 14])))
ctx.set_option("tc_dbg_mode", 0)
run()
ctx.sync()
d = dbg.cpu().numpy().reshape(-1, 16)
setup = d[:, 1] - d[:, 0]
first = d[:, 2] - d[:, 1]
main = d[:, 3] - d[:, 2]
wait_acc = d[:, 4] - d[:, 3]
epi = d[:, 5] - d[:, 4]
tot = d[:, 5] - d[:, 0]
if version >= 2:  # three n-tiles per row block: 128, 128 and 48 columns; report the full-width tiles
    full = np.arange(len(d)) % 3 != 2
    d, setup, first, main, wait_acc, epi, tot = (v[full] for v in (d, setup, first, main, wait_acc, epi, tot))
for name, v in [("setup (barriers, TMEM alloc)", setup), ("first TMA stage lands", first), ("main loop (20 k-blocks, converters' view)", main),
                ("wait for last MMA", wait_acc), ("epilogue", epi), ("total", tot),
                ("  producer: waiting for a free stage (sum)", d[:, 8]), ("  MMA warp: waiting for converters (sum)", d[:, 9]),
                ("  converter warp 2: waiting for TMA (sum)", d[:, 10]), ("  converter warp 2: converting (sum)", d[:, 11]),
                ("  epilogue: X tile TMA load wait", d[:, 12]), ("  epilogue: first tmem_ld per chunk (sum)", d[:, 13]),
                ("  epilogue: math + staging writes incl. store wait (sum)", d[:, 14]), ("  epilogue: fence + group barrier (sum)", d[:, 15])]:
    print("%-45s median %8.0f clk  p10 %8.0f  p90 %8.0f" % (name, np.median(v), np.percentile(v, 10), np.percentile(v, 90)))
span = (d[:, 7].max() - d[:, 7].min()) / 1e6
print("CTAs %d, SMs used %d, first-to-last CTA start %.3f ms" % (len(d), len(np.unique(d[:, 6])), span))
