"""A_hat = D^-1/2 (Adj - diag + I) D^-1/2 built on the GPU.

The reference builds the normalised adjacency on the host with networkx + SciPy right before the hot
path (gcnmain.py:115-128): ``adj = nx.adjacency_matrix(graph)``; ``setdiag(0)``; ``setdiag(1)``; row sums;
``1/sqrt``; ``D * adj * D``; ``astype(float32)``.  Here the same matrix comes out of ``libgcnb200.so``
(csrc/adjacency.cu: count, scan, scatter, per-row bitonic sort, compact + scale) from an undirected edge list,
bit-identical to the SciPy result: indices exactly, values because the float64 arithmetic and the final rounding to
float32 are the reference's.  The graph is unweighted, like the reference's (no edge carries a 'w' attribute,
data.py:56,61).

``normalize_adjacency(adj)`` is the drop-in for gcnmain.py:117-128; ``normalized_adjacency_from_edges`` takes the
edge list directly (what ``DataLoader.get_graph`` holds before networkx materialises the matrix).
There is no CPU fallback: without the CUDA library / a GPU these functions raise.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.sparse as sp
import torch

from . import capi
from .layers import get_dev


def build_on_device(d, u, v, n):
    """Device edge arrays (int32 tensors) -> (rowptr, colidx, val) device tensors of A_hat.  ``d`` is a layers._Dev."""
    n = int(n)
    ne = int(u.numel())
    lib = d.ctx.lib
    need = int(lib.gcnb_adj_workspace_bytes(ne, n))
    work = torch.empty(max(need, 256), dtype=torch.uint8, device=d.dev)
    rowptr = torch.empty(n + 1, dtype=torch.int32, device=d.dev)
    nnz = C.c_int64(0)
    d.fence()
    p = lambda t: C.c_void_p(t.data_ptr())
    d.ctx.call("gcnb_adj_build_rows", p(u) if ne else None, p(v) if ne else None, ne, n, p(work), work.numel(), p(rowptr),
               C.byref(nnz))
    colidx = torch.empty(max(nnz.value, 1), dtype=torch.int32, device=d.dev)
    val = torch.empty(max(nnz.value, 1), dtype=torch.float32, device=d.dev)
    d.fence()
    d.ctx.call("gcnb_adj_fill_f32", ne, n, p(work), p(rowptr), p(colidx), p(val))
    d.ctx.sync()
    return rowptr, colidx[: nnz.value], val[: nnz.value]


def normalized_adjacency_from_edges(u, v, n, device=None):
    """Undirected edge list (any order, duplicates, reversed pairs and self loops allowed) -> A_hat as a SciPy CSR
    (float32 values, int32 indices, columns ascending inside each row)."""
    d = get_dev(device)
    u = np.ascontiguousarray(u)
    v = np.ascontiguousarray(v)
    if u.shape != v.shape or u.ndim != 1:
        raise ValueError("u and v must be 1-D arrays of the same length")
    n = int(n)
    if n < 0 or n >= 2**31 - 1:
        raise ValueError("number of nodes must fit int32")
    if len(u) and (min(u.min(), v.min()) < 0 or max(u.max(), v.max()) >= n):
        raise ValueError("edge list holds a node id outside [0, %d)" % n)
    du = d.upload(u.astype(np.int32, copy=False))
    dv = d.upload(v.astype(np.int32, copy=False))
    rowptr, colidx, val = build_on_device(d, du, dv, n)
    h_rowptr = np.empty(n + 1, dtype=np.int32)
    nnz = int(colidx.numel())
    h_col = np.empty(nnz, dtype=np.int32)
    h_val = np.empty(nnz, dtype=np.float32)
    d.ctx.call("gcnb_d2h", C.c_void_p(h_rowptr.ctypes.data), C.c_void_p(rowptr.data_ptr()), h_rowptr.nbytes)
    if nnz:
        d.ctx.call("gcnb_d2h", C.c_void_p(h_col.ctypes.data), C.c_void_p(colidx.data_ptr()), h_col.nbytes)
        d.ctx.call("gcnb_d2h", C.c_void_p(h_val.ctypes.data), C.c_void_p(val.data_ptr()), h_val.nbytes)
    d.ctx.sync()
    A = sp.csr_matrix((h_val, h_col, h_rowptr), shape=(n, n))
    A.has_sorted_indices = True
    return A


def normalize_adjacency(adj, dtype="float32", device=None):
    """Drop-in for gcnmain.py:117-128: ``adj`` is the symmetric 0/1 adjacency (any SciPy sparse format) of the
    undirected graph; returns A_hat as float32 CSR.  Edge weights (``nx.adjacency_matrix(..., weight='w')``,
    gcnmain.py:115) go through ``normalize_weighted_adjacency``; asymmetric inputs are rejected (``nx.Graph``)."""
    if dtype != "float32":
        raise ValueError("the B200 path produces float32 (gcnmain.py:167 fixes dtype to float32)")
    adj = sp.coo_matrix(adj)
    if adj.shape[0] != adj.shape[1]:
        raise ValueError("adjacency must be square")
    n = adj.shape[0]
    keep = adj.data != 0
    if not np.all(adj.data[keep] == 1):
        return normalize_weighted_adjacency(adj, device=device)
    u, v = adj.row[keep], adj.col[keep]
    A = normalized_adjacency_from_edges(u, v, n, device=device)
    # an asymmetric input would have been symmetrised by the edge-list builder: refuse instead of guessing
    pat = sp.csr_matrix((np.ones(len(u), dtype=np.int8), (u, v)), shape=(n, n))
    pat.sum_duplicates()
    off_diag = pat.nnz - int(np.count_nonzero(pat.diagonal()))
    if A.nnz != off_diag + n:
        raise ValueError("adjacency is not symmetric")
    return A


def normalize_weighted_adjacency(adj, device=None):
    """gcnmain.py:115-128 for a graph whose edges carry weights: ``adj`` is the symmetric weighted adjacency SciPy matrix
    ``nx.adjacency_matrix(graph, weight='w')`` returns.  The diagonal is forced to 1 on the host (``setdiag(0)``;
    ``setdiag(1)`` are structure edits); row sums, ``1/sqrt`` (inf -> 0) and ``D * adj * D`` run on the GPU in float64
    with one rounding to float32 (``gcnb_adj_normalize_weighted_f64``).  Returns float32 CSR, columns ascending."""
    d = get_dev(device)
    W = sp.csr_matrix(adj, dtype=np.float64, copy=True)
    if W.shape[0] != W.shape[1]:
        raise ValueError("adjacency must be square")
    n = W.shape[0]
    if n >= 2**31 - 1 or W.nnz + n >= 2**31 - 1:
        raise ValueError("adjacency must fit int32 indices")
    W.sum_duplicates()
    D = (W - W.T).tocsr()
    if D.nnz and float(np.abs(D.data).max()) != 0.0:
        raise ValueError("adjacency is not symmetric")
    W.setdiag(1.0)  # setdiag(0) then setdiag(1): every diagonal entry present and equal to 1
    W = W.tocsr()
    W.sort_indices()
    rowptr = np.ascontiguousarray(W.indptr, dtype=np.int32)
    col = np.ascontiguousarray(W.indices, dtype=np.int32)
    dr, dc, dw = d.upload(rowptr), d.upload(col), d.upload(np.ascontiguousarray(W.data, dtype=np.float64))
    dinv = torch.empty(max(n, 1), dtype=torch.float64, device=d.dev)
    val = torch.empty(max(W.nnz, 1), dtype=torch.float32, device=d.dev)
    d.fence()
    p = lambda t: C.c_void_p(t.data_ptr())
    d.ctx.call("gcnb_adj_normalize_weighted_f64", p(dr), p(dc), p(dw), n, p(dinv), p(val))
    h_val = np.empty(W.nnz, dtype=np.float32)
    if W.nnz:
        d.ctx.call("gcnb_d2h", C.c_void_p(h_val.ctypes.data), p(val), h_val.nbytes)
    d.ctx.sync()
    A = sp.csr_matrix((h_val, col, rowptr), shape=(n, n))
    A.has_sorted_indices = True
    return A
