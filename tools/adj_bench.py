#!/usr/bin/env python
"""A_hat construction on the GPU box: csrc/adjacency.cu vs the NumPy/SciPy restatement of gcnmain.py:115-128.

    python tools/adj_bench.py [--n 500000] [--deg 32] [--alpha 2.0]
Prints device time (CUDA events, edges already resident), bytes moved per the algorithmic model and the host time of
the NumPy path on the same edge list; checks bit-equality."""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from geographconv_b200 import adjacency, synth  # noqa: E402
from geographconv_b200.layers import get_dev  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=500000)
    ap.add_argument("--deg", type=int, default=32)
    ap.add_argument("--alpha", type=float, default=None)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    n = args.n
    rng = np.random.RandomState(77)
    m = n * (args.deg - 1) // 2
    if args.alpha is None:
        u = rng.randint(0, n, size=m, dtype=np.int64)
        v = rng.randint(0, n, size=m, dtype=np.int64)
    else:
        w = np.arange(1, n + 1, dtype=np.float64) ** (-1.0 / (args.alpha - 1.0))
        cdf = np.cumsum(w)
        cdf /= cdf[-1]
        perm = rng.permutation(n)
        u = perm[np.minimum(np.searchsorted(cdf, rng.random_sample(m)), n - 1)]
        v = perm[np.minimum(np.searchsorted(cdf, rng.random_sample(m)), n - 1)]
    d = get_dev(0)
    du = d.upload(u.astype(np.int32))
    dv = d.upload(v.astype(np.int32))
    d.ctx.sync()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    out = None
    for i in range(args.reps + 1):
        if i == 1:
            ev[0].record(d.stream)
        out = adjacency.build_on_device(d, du, dv, n)
    ev[1].record(d.stream)
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / args.reps
    nnz = int(out[1].numel())
    raw = 2 * m + n
    # algorithmic bytes: edges read twice (count, scatter), raw buckets written once, read+written by the sort, read by
    # the fill; rowptr / degree arrays; CSR written once
    alg = 2 * (8 * m) + 4 * raw * 4 + 6 * 4 * n + nnz * 8
    t0 = time.perf_counter()
    R = synth.normalized_adjacency_from_edges(u, v, n)
    host_s = time.perf_counter() - t0
    A = adjacency.normalized_adjacency_from_edges(u, v, n)
    same = (np.array_equal(A.indptr, R.indptr) and np.array_equal(A.indices, R.indices) and np.array_equal(A.data, R.data))
    print("A_hat build: n=%d edges=%d nnz=%d alpha=%s max-row=%d | GPU %.3f ms (%.1f GB/s algorithmic, %d bytes) | "
          "NumPy host %.2f s (%.0fx) | bit-identical=%s"
          % (n, m, nnz, args.alpha, int(np.diff(R.indptr).max()), ms, alg / ms / 1e6, alg, host_s, host_s * 1e3 / ms, same),
          flush=True)


if __name__ == "__main__":
    main()
