"""Single-layer functional entry points over host arrays (reference gcnmodel.py:29-42,72-157).

Each call uploads its operands, runs the CUDA kernels through the C ABI and returns a host
ndarray.  They exist for the layer-level surface of the reference module
and for kernel parity tests; ``GraphConv`` itself keeps everything resident (engine.py).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.sparse as sp
import torch

from . import capi
from .capi import ACT, GcnbCsr, GcnbEpilogue
from .partition import ld_of

_contexts = {}


class _Dev:
    """Context + allocator helpers for one device (cached)."""

    def __init__(self, device):
        if not torch.cuda.is_available():
            raise capi.GcnbError("geographconv_b200 needs a B200 GPU (no CPU fallback)")
        device = torch.cuda.current_device() if device is None else int(device)
        self.dev = torch.device("cuda", device)
        torch.cuda.set_device(self.dev)
        self.stream = torch.cuda.Stream(self.dev)
        self.ctx = capi.Context(device, C.c_void_p(self.stream.cuda_stream))
        self.ws = None
        self.spmm_chunk = 256

    def fence(self):
        """torch-side fills (current stream) happen-before the kernels on this context's stream."""
        self.stream.wait_stream(torch.cuda.current_stream(self.dev))

    def ensure_ws(self, nbytes):
        nbytes = max(int(nbytes), 1 << 20)
        if self.ws is None or self.ws.numel() < nbytes:
            self.ctx.sync()
            self.ws = torch.empty(nbytes, dtype=torch.uint8, device=self.dev)
            self.ctx.call("gcnb_set_workspace", C.c_void_p(self.ws.data_ptr()), self.ws.numel())

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        t = torch.empty(arr.size, dtype={np.dtype("float32"): torch.float32, np.dtype("int32"): torch.int32,
                                         np.dtype("float64"): torch.float64}[arr.dtype], device=self.dev)
        if arr.size:
            self.ctx.call("gcnb_h2d", C.c_void_p(t.data_ptr()), C.c_void_p(arr.ctypes.data), arr.nbytes)
        return t

    def dense(self, M):
        """Host (r x c) -> device (r x ld) zero padded; returns (tensor, ld)."""
        M = np.asarray(M, dtype=np.float32)
        ld = ld_of(M.shape[1])
        buf = np.zeros((M.shape[0], ld), dtype=np.float32)
        buf[:, : M.shape[1]] = M
        return self.upload(buf.reshape(-1)), ld

    def vec(self, v):
        v = np.asarray(v, dtype=np.float32)
        buf = np.zeros(ld_of(len(v)), dtype=np.float32)
        buf[: len(v)] = v
        return self.upload(buf)

    def download(self, t, rows, ld, cols):
        host = np.empty((rows, ld), dtype=np.float32)
        if host.size:
            self.ctx.call("gcnb_d2h", C.c_void_p(host.ctypes.data), C.c_void_p(t.data_ptr()), host.nbytes)
        self.ctx.sync()
        return np.ascontiguousarray(host[:, :cols])


def get_dev(device=None):
    key = torch.cuda.current_device() if device is None else int(device)
    if key not in _contexts:
        _contexts[key] = _Dev(key)
    return _contexts[key]


class CsrOnDevice:
    def __init__(self, d, M, tag=capi.TAG_SPMM_A, chunk=None):
        M = sp.csr_matrix(M)
        M.sort_indices()
        rowptr = np.ascontiguousarray(M.indptr, dtype=np.int32)
        items, long_rows, n_slots = capi.csr_plan(rowptr, int(chunk or d.spmm_chunk))
        self.keep = [d.upload(rowptr), d.upload(np.ascontiguousarray(M.indices, dtype=np.int32)),
                     d.upload(np.ascontiguousarray(M.data, dtype=np.float32)), d.upload(items.reshape(-1)),
                     d.upload(long_rows.reshape(-1))]
        s = GcnbCsr()
        s.n_rows, s.n_cols, s.nnz = M.shape[0], M.shape[1], int(M.nnz)
        s.rowptr, s.colidx, s.val, s.items = (t.data_ptr() for t in self.keep[:4])
        s.n_items = len(items)
        s.long_rows = self.keep[4].data_ptr() if len(long_rows) else None
        s.n_long, s.n_slots, s.tag = len(long_rows), int(n_slots), int(tag)
        s.engine, s.unroll = -1, 0  # the context options decide (tests sweep them)
        self.struct = s
        self.shape = M.shape


def spmm(A, B, bias=None, act="linear", softmax=False, dropout_p=0.0, seed=0, row0=0, accumulate_into=None,
         want_logits=False, chunk=None, device=None, variant=None, panel=None, unroll=None, accumulate_mode=None):
    """epilogue(A.B) for CSR ``A`` and dense ``B`` -- ``theano.sparse.structured_dot`` plus the
    fused bias / activation / dropout / softmax (gcnmodel.py:39-42,130-136,153-157)."""
    d = get_dev(device)
    if variant is not None:
        d.ctx.set_option("spmm_variant", variant)
    if panel is not None:
        d.ctx.set_option("spmm_panel", panel)
    if unroll is not None:
        d.ctx.set_option("spmm_unroll", unroll)
    B = np.asarray(B, dtype=np.float32)
    K = B.shape[1]
    csr = CsrOnDevice(d, A, chunk=chunk)
    dB, ldb = d.dense(B)
    rows = csr.shape[0]
    ldc = ld_of(K)
    if accumulate_into is not None:
        dC, _ = d.dense(accumulate_into)
    else:
        dC = torch.zeros(max(rows, 1) * ldc, dtype=torch.float32, device=d.dev)
    dL = torch.zeros(max(rows, 1) * ldc, dtype=torch.float32, device=d.dev) if want_logits else None
    d.ensure_ws(d.ctx.lib.gcnb_spmm_workspace_bytes(C.byref(csr.struct), K))
    epi = GcnbEpilogue()
    db = d.vec(bias) if bias is not None else None
    epi.bias = db.data_ptr() if db is not None else None
    epi.act, epi.softmax, epi.accumulate = ACT[act], int(bool(softmax)), int(accumulate_into is not None)
    if accumulate_mode is not None:  # 2: C = epilogue(C + A.B), the pre-activation accumulate
        epi.accumulate = int(accumulate_mode)
    epi.dropout_p, epi.seed, epi.row0 = float(dropout_p), int(seed), int(row0)
    epi.logits = dL.data_ptr() if dL is not None else None
    d.fence()
    d.ctx.call("gcnb_spmm_csr_f32", C.byref(csr.struct), C.c_void_p(dB.data_ptr()), ldb, C.c_void_p(dC.data_ptr()), ldc,
               K, C.byref(epi))
    out = d.download(dC, rows, ldc, K)
    if variant is not None:
        d.ctx.set_option("spmm_variant", 0)
    if panel is not None:
        d.ctx.set_option("spmm_panel", 32)
    if unroll is not None:
        d.ctx.set_option("spmm_unroll", 0)
    if want_logits:
        return out, d.download(dL, rows, ldc, K)
    return out


def gemm(A, B, transA=False, transB=False, bias=None, act="linear", accumulate_into=None, device=None, tc=None):
    """act(op(A).op(B) + bias) -- ``T.dot`` and its gradients (gcnmodel.py:126,149,285)."""
    d = get_dev(device)
    if tc is not None:
        d.ctx.set_option("gemm_tc", int(tc))
    A = np.asarray(A, dtype=np.float32)
    B = np.asarray(B, dtype=np.float32)
    M, K = (A.shape[1], A.shape[0]) if transA else A.shape
    N = B.shape[0] if transB else B.shape[1]
    dA, lda = d.dense(A)
    dB, ldb = d.dense(B)
    ldc = ld_of(N)
    if accumulate_into is not None:
        dC, _ = d.dense(accumulate_into)
    else:
        dC = torch.zeros(max(M, 1) * ldc, dtype=torch.float32, device=d.dev)
    db = d.vec(bias) if bias is not None else None
    d.ensure_ws(d.ctx.lib.gcnb_gemm_workspace_bytes(int(transA), M, N, K))
    d.fence()
    d.ctx.call("gcnb_gemm_f32", int(transA), int(transB), M, N, K, C.c_void_p(dA.data_ptr()), lda,
               C.c_void_p(dB.data_ptr()), ldb, C.c_void_p(dC.data_ptr()), ldc, int(accumulate_into is not None),
               C.c_void_p(db.data_ptr()) if db is not None else None, ACT[act])
    out = d.download(dC, M, ldc, N)
    if tc is not None:
        d.ctx.set_option("gemm_tc", 1)
    return out


def highway(S, X, Wh, bh, Wt, bt, act="tanh", device=None, tc=None):
    """Fused highway layer on S = A.X (gcnmodel.py:266,281-288): returns (Y, H, T)."""
    d = get_dev(device)
    if tc is not None:
        d.ctx.set_option("gemm_tc", int(tc))
    S = np.asarray(S, dtype=np.float32)
    n, hd = S.shape
    dS, ld = d.dense(S)
    dX, _ = d.dense(X)
    dWh, ldw = d.dense(Wh)
    dWt, _ = d.dense(Wt)
    dbh, dbt = d.vec(bh), d.vec(bt)
    outs = [torch.zeros(max(n, 1) * ld, dtype=torch.float32, device=d.dev) for _ in range(3)]
    d.ensure_ws(d.ctx.lib.gcnb_highway_workspace_bytes(n, hd))
    p = lambda t: C.c_void_p(t.data_ptr())
    d.fence()
    d.ctx.call("gcnb_highway_fwd_f32", n, hd, p(dS), ld, p(dX), ld, p(dWh), ldw, p(dbh), p(dWt), ldw, p(dbt),
               ACT[act], p(outs[0]), ld, p(outs[1]), ld, p(outs[2]), ld)
    res = tuple(d.download(t, n, ld, hd) for t in outs)
    if tc is not None:
        d.ctx.set_option("gemm_tc", 1)
    return res


def sparse_dense(X, W, b, act="tanh", device=None):
    """SparseInputDenseLayer: act(X.W + b) (gcnmodel.py:39-42)."""
    return spmm(X, W, bias=b, act=act, device=device)


def graph_conv_dense(A, x, W, b, act="tanh", device=None):
    """ConvolutionDenseLayer2/3: act(A.(x.W) + b), softmax when act == 'softmax' (gcnmodel.py:126-157).
    ``W`` None means ``x`` is already the projected operand; ``A`` None skips the convolution
    (gcnmodel.py:129)."""
    q = gemm(x, W, device=device) if W is not None else np.asarray(x, dtype=np.float32)
    if A is None:
        if act == "softmax":
            # the reference skips the convolution when A is falsy and still returns softmax(x.W + b)
            # (gcnmodel.py:152-157): the same fused bias + row-softmax epilogue behind an identity graph
            eye = sp.identity(q.shape[0], dtype=np.float32, format="csr")
            return spmm(eye, q, bias=b, softmax=True, device=device)
        return gemm(q, np.eye(q.shape[1], dtype=np.float32), bias=b, act=act, device=device)
    if act == "softmax":
        return spmm(A, q, bias=b, softmax=True, device=device)
    return spmm(A, q, bias=b, act=act, device=device)


def gate_mix(t, h1, h2, device=None):
    """t*h1 + (1-t)*h2 (MultiplicativeGatingLayer, gcnmodel.py:266) on the device."""
    d = get_dev(device)
    t = np.asarray(t, dtype=np.float32)
    n, k = t.shape
    dT, ld = d.dense(t)
    dH, _ = d.dense(h1)
    dX, _ = d.dense(h2)
    dY = torch.zeros(max(n, 1) * ld, dtype=torch.float32, device=d.dev)
    p = lambda x: C.c_void_p(x.data_ptr())
    d.fence()
    d.ctx.call("gcnb_highway_mix_f32", n, k, p(dH), ld, p(dT), ld, p(dX), ld, p(dY), ld)
    return d.download(dY, n, ld, k)


def dropout_keep_mask(n_rows, k, p, seed, row0=0, device=None):
    """uint8 keep mask (n_rows x k) of the engine's Philox dropout stream (gcnb_dropout_mask_u8)."""
    d = get_dev(device)
    m = torch.empty(max(n_rows * k, 1), dtype=torch.uint8, device=d.dev)
    d.fence()
    d.ctx.call("gcnb_dropout_mask_u8", int(n_rows), int(k), float(p), int(seed) & (2**64 - 1), int(row0),
               C.c_void_p(m.data_ptr()))
    host = np.empty((n_rows, k), dtype=np.uint8)
    if host.size:
        d.ctx.call("gcnb_d2h", C.c_void_p(host.ctypes.data), C.c_void_p(m.data_ptr()), host.nbytes)
    d.ctx.sync()
    return host
