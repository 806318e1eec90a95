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


def synthetic_graph(n, avg_degree, seed=77, alpha=None):
    """A_hat for a random undirected graph with about ``avg_degree`` nonzeros per row.

    ``alpha=None``: endpoints uniform.  Otherwise Chung-Lu endpoints with weights
    w_i ~ i^(-1/(alpha-1)) (power-law degree exponent ``alpha``; BASELINE.json configs[4]).
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
    return normalized_adjacency_from_edges(u, v, n)


def synthetic_features(n, f, nnz_per_row, seed=77, zipf_s=1.1, chunk_rows=65536):
    """N x F float32 CSR bag-of-words with Zipf term ids, binary-TF x IDF, L2-normalised rows."""
    rng = np.random.RandomState(seed + 1)
    w = np.arange(1, f + 1, dtype=np.float64) ** (-zipf_s)
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    term_of_rank = rng.permutation(f).astype(np.int64)  # frequent terms get arbitrary ids
    col_chunks, cnt_chunks = [], []
    for r0 in range(0, n, chunk_rows):
        r1 = min(n, r0 + chunk_rows)
        rr = r1 - r0
        ranks = np.minimum(np.searchsorted(cdf, rng.random_sample(rr * nnz_per_row)), f - 1)
        cols = term_of_rank[ranks]
        keys = np.repeat(np.arange(rr, dtype=np.int64), nnz_per_row) * f + cols
        keys = np.unique(keys)  # binary TF: a term counts once per row
        rows = keys // f
        col_chunks.append((keys - rows * f).astype(np.int32))
        cnt_chunks.append(np.bincount(rows, minlength=rr))
    cols = np.concatenate(col_chunks)
    counts = np.concatenate(cnt_chunks)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    df = np.bincount(cols, minlength=f).astype(np.float64)
    idf = np.log((1.0 + n) / (1.0 + df)) + 1.0  # sklearn smooth_idf
    vals = idf[cols]
    sq = np.add.reduceat(vals * vals, np.minimum(rowptr[:-1], max(len(vals) - 1, 0))) if len(vals) else np.zeros(n)
    sq = np.where(counts > 0, sq, 1.0)
    vals = (vals / np.sqrt(np.repeat(sq, counts))).astype(np.float32)
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


def synthetic_problem(name_or_cfg, seed=77, alpha=None):
    """(A_hat, X, Y, train_idx, dev_idx, test_idx, cfg) for one of ``CONFIGS`` or a cfg dict."""
    cfg = dict(CONFIGS[name_or_cfg]) if isinstance(name_or_cfg, str) else dict(name_or_cfg)
    A = synthetic_graph(cfg["n"], cfg["deg"], seed, alpha)
    X = synthetic_features(cfg["n"], cfg["f"], cfg["xnnz"], seed)
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
