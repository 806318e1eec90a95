"""Synthetic inputs of the shapes BASELINE.json names.

The real GeoText / TwitterUS dumps are not redistributable (reference README.md:24), so the
bench and the parity tests run on generated data with the same structure:

* ``synthetic_graph``  -- undirected @-mention-like graph -> A_hat = D^-1/2 (Adj - diag + I) D^-1/2
  float32 CSR with int32 indices, rows sorted by column: the matrix gcnmain.py:115-128 builds.
* ``synthetic_features`` -- N x F bag-of-words CSR: Zipf(1.1) term ids, binary TF x smooth IDF,
  rows L2-normalised (what data.py:275-278's TfidfVectorizer(binary=True, norm='l2') yields).
* ``synthetic_dump``   -- the 13-tuple ``dump.pkl`` layout gcnmain.py:153,170 reads.

All NumPy, chunked so the 500k x 256 case stays within a few GB of host memory.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

CONFIGS = {
    # name: N, avg degree, F, nnz/row of X, hidden sizes, classes   (BASELINE.json configs[0..3])
    "C1": dict(n=9_500, deg=8, f=10_000, xnnz=128, hid=[300, 300, 300], classes=129),
    "C2": dict(n=450_000, deg=16, f=50_000, xnnz=256, hid=[300, 300, 300], classes=256),
    "C3": dict(n=500_000, deg=32, f=50_000, xnnz=256, hid=[300, 300, 300], classes=256),
    "C4": dict(n=2_000_000, deg=64, f=50_000, xnnz=256, hid=[512] * 6, classes=256),
    "tiny": dict(n=512, deg=6, f=300, xnnz=24, hid=[40, 40, 40], classes=7),
}


def normalized_adjacency_from_edges(u, v, n, dtype=np.float32):
    """Symmetrise, dedupe, force unit self loops, scale by D^-1/2 on both sides.

    Restates gcnmain.py:117-128 (setdiag(0); setdiag(1); row sums; 1/sqrt; D*adj*D) for an
    unweighted edge list (no edge carries a 'w' attribute, data.py:56,61).
    """
    u = np.asarray(u, dtype=np.int64)
    v = np.asarray(v, dtype=np.int64)
    keep = u != v
    u, v = u[keep], v[keep]
    loops = np.arange(n, dtype=np.int64)
    keys = np.concatenate([u * n + v, v * n + u, loops * n + loops])
    keys = np.unique(keys)  # sorted => CSR order with columns ascending inside each row
    rows = keys // n
    cols = (keys - rows * n).astype(np.int32)
    counts = np.bincount(rows, minlength=n)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    with np.errstate(divide="ignore"):
        dinv = 1.0 / np.sqrt(counts.astype(np.float64))
    dinv[np.isinf(dinv)] = 0
    vals = (dinv[rows] * dinv[cols]).astype(dtype)
    A = sp.csr_matrix((vals, cols, rowptr.astype(np.int32)), shape=(n, n))
    A.has_sorted_indices = True
    return A


def synthetic_graph(n, avg_degree, seed=77, alpha=None, builder=None):
    """A_hat for a random undirected graph with about ``avg_degree`` nonzeros per row.

    ``alpha=None``: endpoints uniform.  Otherwise Chung-Lu endpoints with weights
    w_i ~ i^(-1/(alpha-1)) (power-law degree exponent ``alpha``; BASELINE.json configs[4]).
    ``builder(u, v, n)``: what turns the edge list into A_hat -- default the NumPy restatement of
    gcnmain.py:115-128 below; ``geographconv_b200.adjacency.normalized_adjacency_from_edges`` builds
    the bit-identical matrix on the GPU (seconds instead of minutes at N = 2M, degree 64).
    """
    rng = np.random.RandomState(seed)
    m = max(int(n * max(avg_degree - 1, 0) / 2), 0)
    if alpha is None:
        u = rng.randint(0, n, size=m, dtype=np.int64)
        v = rng.randint(0, n, size=m, dtype=np.int64)
    else:
        w = np.arange(1, n + 1, dtype=np.float64) ** (-1.0 / (alpha - 1.0))
        cdf = np.cumsum(w)
        cdf /= cdf[-1]
        perm = rng.permutation(n)  # hubs are not the first rows
        u = perm[np.minimum(np.searchsorted(cdf, rng.random_sample(m)), n - 1)]
        v = perm[np.minimum(np.searchsorted(cdf, rng.random_sample(m)), n - 1)]
    return (builder or normalized_adjacency_from_edges)(u, v, n)


def _zipf_draws_for_unique(p, target):
    """Number of with-replacement draws from pmf ``p`` whose expected count of distinct values is ``target``."""
    lo, hi = float(target), float(target) * 64.0
    l1p = np.log1p(-np.minimum(p, 1 - 1e-12))
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        if np.sum(-np.expm1(mid * l1p)) < target:
            lo = mid
        else:
            hi = mid
    return int(round(hi))


def synthetic_features(n, f, nnz_per_row, seed=77, zipf_s=1.1, chunk_rows=32768, row_range=None):
    """N x F float32 CSR bag-of-words: Zipf(1.1) term ids, binary TF x IDF, L2-normalised rows
    (what data.py:275-278's TfidfVectorizer(binary=True, norm='l2') yields), about ``nnz_per_row``
    distinct terms per row.

    Rows are generated in independent chunks (own RNG stream per chunk) and the IDF comes from the
    model's expected document frequency, so a rank can generate just its ``row_range`` = (r0, r1):
    rows outside it are left empty in the returned N x F matrix.
    """
    w = np.arange(1, f + 1, dtype=np.float64) ** (-zipf_s)
    p = w / w.sum()
    cdf = np.cumsum(p)
    cdf[-1] = 1.0
    draws = _zipf_draws_for_unique(p, min(nnz_per_row, 0.5 * f))
    term_of_rank = np.random.RandomState(seed + 1).permutation(f).astype(np.int64)  # frequent terms get arbitrary ids
    df_rank = n * -np.expm1(draws * np.log1p(-np.minimum(p, 1 - 1e-12)))  # expected document frequency by rank
    idf = np.empty(f, dtype=np.float64)
    idf[term_of_rank] = np.log((1.0 + n) / (1.0 + df_rank)) + 1.0  # sklearn smooth_idf
    chunk_rows = int(min(chunk_rows, max((2**31 - 1) // f, 1)))  # row*f + col fits int32
    r_lo, r_hi = (0, n) if row_range is None else (int(row_range[0]), int(row_range[1]))
    counts = np.zeros(n, dtype=np.int64)
    col_chunks, val_chunks = [], []
    for c0 in range(0, n, chunk_rows):
        c1 = min(n, c0 + chunk_rows)
        if c1 <= r_lo or c0 >= r_hi:
            continue
        rr = c1 - c0
        rng = np.random.RandomState((seed * 1000003 + 7919 * (c0 // chunk_rows) + 11) % (2**32))
        ranks = np.minimum(np.searchsorted(cdf, rng.random_sample(rr * draws)), f - 1)
        keys = (np.repeat(np.arange(rr, dtype=np.int64), draws) * f + term_of_rank[ranks]).astype(np.int32)
        keys.sort()
        keys = keys[np.concatenate(([True], keys[1:] != keys[:-1]))]  # binary TF: a term counts once per row
        rows = keys // f
        cols = keys - rows * f
        cnt = np.bincount(rows, minlength=rr)
        vals = idf[cols]
        starts = np.concatenate(([0], np.cumsum(cnt)[:-1]))
        sq = np.add.reduceat(vals * vals, np.minimum(starts, max(len(vals) - 1, 0))) if len(vals) else np.zeros(rr)
        sq = np.where(cnt > 0, sq, 1.0)
        vals = (vals / np.sqrt(np.repeat(sq, cnt))).astype(np.float32)
        lo, hi = max(c0, r_lo) - c0, min(c1, r_hi) - c0  # keep only the requested rows of this chunk
        if lo > 0 or hi < rr:
            e0, e1 = starts[lo] if lo < rr else len(cols), (starts[hi] if hi < rr else len(cols))
            cols, vals, cnt = cols[e0:e1], vals[e0:e1], cnt[lo:hi]
        counts[c0 + lo:c0 + hi] = cnt
        col_chunks.append(cols.astype(np.int32))
        val_chunks.append(vals)
    cols = np.concatenate(col_chunks) if col_chunks else np.zeros(0, np.int32)
    vals = np.concatenate(val_chunks) if val_chunks else np.zeros(0, np.float32)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    X = sp.csr_matrix((vals, cols, rowptr.astype(np.int32)), shape=(n, f))
    X.has_sorted_indices = True
    return X


def synthetic_labels(n, n_classes, seed=77):
    rng = np.random.RandomState(seed + 2)
    y = rng.randint(0, n_classes, size=n).astype(np.int32)
    y[:n_classes] = np.arange(n_classes, dtype=np.int32)  # every class occurs => max(Y)+1 == C
    return y


def split_indices(n):
    """60/20/20 contiguous train/dev/test like gcnmain.py:189,211-212."""
    n_tr = int(0.6 * n)
    n_dev = int(0.2 * n)
    idx = np.arange(n, dtype=np.int32)
    return idx[:n_tr], idx[n_tr:n_tr + n_dev], idx[n_tr + n_dev:]


def synthetic_problem(name_or_cfg, seed=77, alpha=None, row_range=None, graph_builder=None):
    """(A_hat, X, Y, train_idx, dev_idx, test_idx, cfg) for one of ``CONFIGS`` or a cfg dict.
    ``row_range`` = (r0, r1): generate only those rows of X (others empty) -- what one rank of a
    row-partitioned run needs.  ``graph_builder``: see ``synthetic_graph``."""
    cfg = dict(CONFIGS[name_or_cfg]) if isinstance(name_or_cfg, str) else dict(name_or_cfg)
    A = synthetic_graph(cfg["n"], cfg["deg"], seed, alpha, builder=graph_builder)
    X = synthetic_features(cfg["n"], cfg["f"], cfg["xnnz"], seed, row_range=row_range)
    Y = synthetic_labels(cfg["n"], cfg["classes"], seed)
    tr, dev, te = split_indices(cfg["n"])
    return A, X, Y, tr, dev, te, cfg


def synthetic_dump(name_or_cfg, seed=77):
    """The 13-tuple gcnmain.preprocess_data returns / dump.pkl stores (gcnmain.py:153)."""
    A, X, Y, tr, dev, te, cfg = synthetic_problem(name_or_cfg, seed)
    rng = np.random.RandomState(seed + 3)
    C = cfg["classes"]
    lat = rng.uniform(25, 49, size=C)
    lon = rng.uniform(-124, -67, size=C)
    classLatMedian = {str(c): float(lat[c]) for c in range(C)}
    classLonMedian = {str(c): float(lon[c]) for c in range(C)}
    users = ["u%d" % i for i in range(cfg["n"])]
    userLocation = {users[i]: "%f,%f" % (lat[Y[i]] + rng.normal(0, 0.5), lon[Y[i]] + rng.normal(0, 0.5))
                    for i in range(cfg["n"])}
    U_train = [users[i] for i in tr]
    U_dev = [users[i] for i in dev]
    U_test = [users[i] for i in te]
    return (A, X[tr], Y[tr], X[dev], Y[dev], X[te], Y[te], U_train, U_dev, U_test,
            classLatMedian, classLonMedian, userLocation)
