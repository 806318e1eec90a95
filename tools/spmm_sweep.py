#!/usr/bin/env python
"""SpMM micro-benchmark on the GPU box: A_hat.H / X.W0 / X^T.dz at BASELINE shapes, gather-engine variants.

    python tools/spmm_sweep.py [--n 500000] [--deg 32] [--k 300] [--alpha 2.0] [--what a,x,xt]
Prints one line per configuration: ms per launch, B_touch GB/s, fraction of the measured HBM peak.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from geographconv_b200 import capi, synth  # noqa: E402
from geographconv_b200.engine import DeviceCsr, HostCsr  # noqa: E402
from geographconv_b200.partition import ld_of, transpose_csr  # noqa: E402


class Mini:
    def __init__(self):
        self.dev = torch.device("cuda", 0)
        self.stream = torch.cuda.Stream(self.dev)
        self.ctx = capi.Context(0, C.c_void_p(self.stream.cuda_stream))
        self.ws = torch.empty(3 << 30, dtype=torch.uint8, device=self.dev)
        self.ctx.call("gcnb_set_workspace", C.c_void_p(self.ws.data_ptr()), self.ws.numel())


def time_spmm(m, csr, B, ldb, Cbuf, ldc, K, reps=5):
    """``csr`` may be a list of column-range parts of one matrix: the first stores, the others add (accumulate = 2)."""
    from geographconv_b200.capi import GcnbEpilogue
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    parts = csr if isinstance(csr, list) else [csr]
    acc = GcnbEpilogue()
    acc.accumulate = 2
    for i in range(2 + reps):
        if i == 2:
            ev[0].record(m.stream)
        for j, part in enumerate(parts):
            m.ctx.call("gcnb_spmm_csr_f32", C.byref(part.struct), C.c_void_p(B.data_ptr()), ldb,
                       C.c_void_p(Cbuf.data_ptr()), ldc, K, C.byref(acc) if j else None)
    ev[1].record(m.stream)
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=500000)
    ap.add_argument("--deg", type=int, default=32)
    ap.add_argument("--k", type=int, default=300)
    ap.add_argument("--f", type=int, default=50000)
    ap.add_argument("--xnnz", type=int, default=256)
    ap.add_argument("--alpha", type=float, default=None)
    ap.add_argument("--what", default="a")
    ap.add_argument("--chunks", default="256")
    ap.add_argument("--variants", default="0:0,0:2,0:4,0:8,1:0",
                    help="comma list of engine:unroll[:panel[:policy]] (engine 2 = L2-resident column panels)")
    ap.add_argument("--blocks", default="1", help="comma list: cut rows at this many equal column ranges (plan_col_blocks)")
    ap.add_argument("--colsplit", default="", help="comma list: also time the product as this many launches, one per "
                    "equal column range of the matrix (first stores, the rest accumulate), panel engine")
    ap.add_argument("--hot", action="store_true", help="X / X^T: drop the dense hot-column block first, like the engine")
    args = ap.parse_args()
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    m = Mini()
    K, ld = args.k, ld_of(args.k)
    mats = {}
    if "a" in args.what.split(","):
        mats["A_hat.H"] = (synth.synthetic_graph(args.n, args.deg, 77, args.alpha), args.n)
    if "x" in args.what.split(",") or "xt" in args.what.split(","):
        X = synth.synthetic_features(args.n, args.f, args.xnnz, 77)
        if args.hot:
            from geographconv_b200.engine import split_hot_columns
            X = split_hot_columns(X, 0.05, 1024)[2]
        if "x" in args.what.split(","):
            mats["X.W0"] = (X, args.f)
        if "xt" in args.what.split(","):
            mats["X^T.dz"] = (transpose_csr(X), args.n)
    for name, (M, brows) in mats.items():
        B = torch.randn(brows, ld, device=m.dev)
        B[:, K:] = 0
        Cbuf = torch.zeros(M.shape[0], ld, device=m.dev)
        m.stream.wait_stream(torch.cuda.current_stream())
        deg = np.diff(M.indptr)
        print("%s: rows %d nnz %d max-row %d  K=%d" % (name, M.shape[0], M.nnz, int(deg.max()), K), flush=True)
        for chunk, blocks in [(int(c), int(b)) for c in args.chunks.split(",") for b in args.blocks.split(",")]:
            class E:  # what DeviceCsr needs from an engine
                dev, ctx = m.dev, m.ctx
            csr = DeviceCsr(E, HostCsr(M, chunk, blocks), 0)
            if blocks > 1:
                print("  column blocks %d:" % blocks, flush=True)
            csr.struct.engine, csr.struct.unroll = -1, 0
            m.ctx.sync()
            for v in args.variants.split(","):
                f = [int(x) for x in v.split(":")] + [0, 0, 0]
                var, unroll, panel, policy = f[0], f[1], f[2] or 32, f[3]
                m.ctx.set_option("spmm_variant", var)
                m.ctx.set_option("spmm_unroll", unroll)
                m.ctx.set_option("spmm_panel", panel)
                m.ctx.set_option("spmm_panel_policy", policy)
                ms = time_spmm(m, csr, B, ld, Cbuf, ld, K)
                gbs = csr.touched_bytes(K) / ms / 1e6
                extra = " panel %d policy %d" % (panel, policy) if var == 2 else ""
                print("  chunk %5d variant %d unroll %d%s: %8.3f ms  %7.1f GB/s  %.3f of HBM peak  (items %d, long rows %d)"
                      % (chunk, var, unroll, extra, ms, gbs, gbs / peak, csr.n_items, csr.n_long), flush=True)
        if args.colsplit:
            colsplit_runs(args, m, M, B, ld, Cbuf, K, peak)


def colsplit_runs(args, m, M, B, ld, Cbuf, K, peak):
    import scipy.sparse as sp
    for nparts in [int(x) for x in args.colsplit.split(",") if x]:
        class E:
            dev, ctx = m.dev, m.ctx
        coo = M.tocoo()
        n_cols = M.shape[1]
        parts = []
        for b in range(nparts):
            lo, hi = (n_cols * b + nparts - 1) // nparts, (n_cols * (b + 1) + nparts - 1) // nparts
            keep = (coo.col >= lo) & (coo.col < hi)
            Mb = sp.csr_matrix((coo.data[keep], (coo.row[keep], coo.col[keep])), shape=M.shape)
            Mb.sort_indices()
            d = DeviceCsr(E, HostCsr(Mb, 1024), 0)
            d.struct.engine, d.struct.unroll = 2, 0
            parts.append(d)
        m.ctx.sync()
        m.ctx.set_option("spmm_panel", 32)
        m.ctx.set_option("spmm_panel_policy", 1)
        ms = time_spmm(m, parts, B, ld, Cbuf, ld, K)
        touched = M.nnz * 8 + (M.shape[0] + 1) * 4 + M.nnz * K * 4 + M.shape[0] * K * 4
        print("  column split into %d launches (panel 32, evict_last): %8.3f ms  %7.1f GB/s  %.3f of HBM peak"
              % (nparts, ms, touched / ms / 1e6, touched / ms / 1e6 / peak), flush=True)


if __name__ == "__main__":
    main()
