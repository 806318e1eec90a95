"""Device engine of the GCN hot path: owns the HBM buffers and sequences the C-ABI kernels.

This is the host side of ``f_train`` / ``f_val`` / ``f_gates`` (reference gcnmodel.py:396-411):
one full-graph forward (+ backward + Adam) per call, all on one CUDA stream, through
``libgcnb200.so`` only.  PyTorch is the HBM allocator and the NCCL plumbing; there is no
torch.sparse, no torch.matmul and no CPU fallback anywhere on this path.

Data layout in HBM:
* every dense matrix is row-major fp32 with leading dimension round_up(cols, 32) floats, so
  each row starts on a 128-byte line; padding columns are zero and stay zero;
* weights, gradients and the two Adam moments are four flat buffers with one shared layout
  (``partition.ParamLayout``, Lasagne ``get_all_param_values`` order);
* sparse operands are CSR (int32 rowptr / colidx, fp32 val) plus the row-item plan the SpMM
  kernel walks; X is kept in both CSR (forward X.W0) and transposed CSR (backward X^T.dz).

Layer arithmetic, per SURVEY.md 8a:
* first layer      H0 = dropout(act(X.W0 + b0))                        (gcnmodel.py:39-42,357)
* highway layer    S = A.x ; h = act(S.Wh + bh) ; t = sigmoid(x.Wt + bt) ; y = t*h + (1-t)*x
                   (gcnmodel.py:126-136,266,281-288; (A.x).Wh == A.(x.Wh) by associativity,
                   the bias is added after both, as in the reference)
* plain layer      y = act(A.(x.W) + b)                                (gcnmodel.py:372)
* output layer     P = softmax(A.(x.Wout) + bout)                      (gcnmodel.py:149-157)

Row-partitioned runs (world > 1): rank p owns the contiguous row block
[p*n_pad, (p+1)*n_pad) of X and of every activation; weights are replicated and their gradients
all-reduced once per step (NCCL).  A graph convolution needs rows of the dense operand that live on
other ranks; two exchange designs are built in (``GCNB_EXCHANGE``):

* ``slice`` (default): A_hat is replicated (nnz*8 bytes), rank q multiplies ALL rows of A_hat by ITS
  column slice of the operand; the transposes before and after the product are stores into the
  peers' memory over NVLink (csrc/peer.cu).  N*K*4*(P-1)/P^2 bytes per rank each way.
* ``gather``: rank p keeps its row block of A_hat and all-gathers the whole N x K operand (NCCL).
  N*K*4*(P-1)/P bytes into every rank -- the round-1 design, kept for comparison and as the path
  when CUDA IPC is unavailable.

Both sum every row in CSR order, so the forward pass is bit-identical to the single-GPU one.
"""
from __future__ import annotations

import copy
import ctypes as C
import os

import numpy as np
import scipy.sparse as sp
import torch

from . import capi
from .capi import ACT, GcnbCsr, GcnbEpilogue
from .partition import (ParamLayout, ld_of, local_index_split, row_blocks, slice_columns, slice_rows, transpose_csr,
                        is_symmetric)

SPMM_CHUNK_DEFAULT = 1024  # nonzeros per row item (rows longer than this are split; see gcnb_csr_plan)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _pinned(arr):
    """Copy a host array into page-locked memory (async H2D needs it); returns the ndarray view."""
    arr = np.ascontiguousarray(arr)
    if arr.size == 0:
        return arr
    t = torch.from_numpy(arr)
    try:
        return t.pin_memory().numpy()
    except RuntimeError:  # no CUDA runtime (CPU-only host): plain memory still works for staging
        return arr


def panel_col_blocks(n_cols, target_bytes=16 << 20, max_blocks=8):
    """Column ranges to cut the rows of a panel-engine operand into so that the live part of one 32-column panel of
    the dense operand (n_cols / blocks rows x 128 B) is about ``target_bytes``.  Measured on C3's X^T.dz (64 MB
    panel): 1 block 9.4 ms, 2 blocks 6.5 ms, 4 blocks 5.3 ms, 8 blocks 5.5 ms (profiles/r1c_spmm_panel_sweep.txt)."""
    return int(min(max_blocks, max(1, -(-int(n_cols) * 128 // int(target_bytes)))))


def plan_col_blocks(rowptr, colidx, n_cols, n_blocks, chunk):
    """Row-item plan whose items never straddle a boundary between ``n_blocks`` equal ranges of column ids.

    The panel engine keeps one 32-column panel of the dense operand in L2; cutting every row at the column-range
    boundaries and running the items range by range (all items of range 0, then range 1, ...) shrinks the live part
    of the panel to n_cols/n_blocks rows: a 64 MB panel (N = 500k) becomes two 32 MB halves that the L2 of either
    die holds.  Same item / long-row format as gcnb_csr_plan (a row with more than one item is a long row whose
    items write partial sums that the fix-up adds in slot order = column order), so results stay deterministic.
    Returns (items[n,4], long_rows[m,3], n_slots)."""
    rowptr = np.asarray(rowptr, dtype=np.int64)
    n_rows = len(rowptr) - 1
    deg = np.diff(rowptr)
    rows_of = np.repeat(np.arange(n_rows, dtype=np.int64), deg)
    key = rows_of * n_cols + np.asarray(colidx, dtype=np.int64)
    base = np.arange(n_rows, dtype=np.int64) * n_cols
    cuts = [rowptr[:-1]]
    for b in range(1, n_blocks):
        cuts.append(np.searchsorted(key, base + (n_cols * b + n_blocks - 1) // n_blocks))
    cuts.append(rowptr[1:])
    seg_len = [cuts[b + 1] - cuts[b] for b in range(n_blocks)]
    pieces = [(l + chunk - 1) // chunk for l in seg_len]
    pieces[0] = np.where(deg == 0, 1, pieces[0])  # an empty row still owns one (empty) item: its epilogue must run
    total = np.sum(pieces, axis=0)
    is_long = total > 1
    slot_base = np.zeros(n_rows, dtype=np.int64)
    slot_base[is_long] = np.cumsum(total[is_long]) - total[is_long]
    n_slots = int(total[is_long].sum())
    long_ids = np.nonzero(is_long)[0]
    long_rows = np.stack([long_ids, slot_base[long_ids], total[long_ids]], axis=1).astype(np.int32) if len(long_ids) \
        else np.zeros((0, 3), dtype=np.int32)
    out = []
    before = np.zeros(n_rows, dtype=np.int64)  # pieces of the row in earlier blocks
    for b in range(n_blocks):
        pb = pieces[b]
        r = np.repeat(np.arange(n_rows, dtype=np.int64), pb)
        q = np.arange(len(r), dtype=np.int64) - np.repeat(np.cumsum(pb) - pb, pb)
        per = (seg_len[b][r] + np.maximum(pb[r], 1) - 1) // np.maximum(pb[r], 1)
        beg = cuts[b][r] + q * per
        end = np.minimum(beg + per, cuts[b + 1][r])
        slot = np.where(is_long[r], slot_base[r] + before[r] + q, -1)
        it = np.stack([r, beg, end, slot], axis=1)
        it = it[np.argsort(-(end - beg), kind="stable")]  # longest first inside a block
        out.append(it)
        before += pb
    items = np.ascontiguousarray(np.concatenate(out, axis=0).astype(np.int32)) if out else np.zeros((0, 4), np.int32)
    return items, long_rows, n_slots


def _fingerprint(M):
    """Cheap content fingerprint of a SciPy sparse matrix (about 4096 strided samples of its value and index arrays):
    an in-place edit of a cached X / A is noticed without hashing gigabytes on every call."""
    M = M if hasattr(M, "indices") and hasattr(M, "data") else M.tocsr()
    if M.nnz == 0:
        return (0.0, 0)
    step = max(1, M.nnz // 4096)
    return (float(np.asarray(M.data[::step], dtype=np.float64).sum()), int(np.asarray(M.indices[::step], dtype=np.int64).sum()))


class HostCsr:
    """Host-side CSR of one SpMM operand: int32 / fp32 arrays in pinned memory plus the row-item plan."""

    def __init__(self, M, chunk, col_blocks=1, row_groups=1):
        M = M.tocsr()
        if not M.has_sorted_indices:
            M = M.copy()
            M.sort_indices()
        if M.shape[0] >= 2**31 or M.shape[1] >= 2**31 or M.nnz >= 2**31:
            raise ValueError("CSR dimensions / nnz per shard must fit int32")
        self.shape = tuple(int(x) for x in M.shape)
        self.nnz = int(M.nnz)
        rowptr = np.ascontiguousarray(M.indptr, dtype=np.int32)
        self.row_bounds = [0, self.shape[0]]  # row groups of the item plan (one group unless ``row_groups`` > 1)
        if col_blocks > 1 and M.nnz:
            items, long_rows, n_slots = plan_col_blocks(rowptr, M.indices, self.shape[1], int(col_blocks), int(chunk))
        else:
            items, long_rows, n_slots = capi.csr_plan(rowptr, int(chunk))
            # Longest items first.  The panel engine walks four to eight items per warp in lock step and only its
            # predicate-free fast path is cheap (8 instructions per gathered float4), so neighbours of equal length pay;
            # the persistent engine pulls items from a counter, so a short tail pays.  Results do not depend on the
            # order (every item owns its output row or its partial-sum slot).
            lens = items[:, 2] - items[:, 1] if len(items) else np.zeros(0, dtype=np.int32)
            # Row groups (X only): the items of consecutive row ranges are kept together, longest first inside a group,
            # so that a product can also be launched group by group while the later rows are still crossing PCIe
            # (Engine.bind's repeated uploads).  One launch over all items computes the same thing.
            if row_groups > 1 and len(long_rows) == 0 and self.shape[0] >= 64 * row_groups:
                per = -(-self.shape[0] // int(row_groups))
                per = -(-per // 128) * 128  # whole 128-row GEMM tiles
                self.row_bounds = list(range(0, self.shape[0], per)) + [self.shape[0]]
                gid = items[:, 0] // per
                items = np.ascontiguousarray(items[np.lexsort((-lens, gid))])
                self.item_bounds = [0] + list(np.cumsum(np.bincount(gid, minlength=len(self.row_bounds) - 1)))
            elif len(lens) and int(lens.max()) > int(lens.min()):
                items = np.ascontiguousarray(items[np.argsort(-lens, kind="stable")])
        if len(self.row_bounds) == 2:
            self.item_bounds = [0, len(items)]
        self.col_blocks = int(col_blocks)
        self.rowptr = _pinned(rowptr)
        self.colidx = _pinned(np.ascontiguousarray(M.indices, dtype=np.int32))
        # column ids below 65536 (the BoW features) cross PCIe as uint16 and are widened on the device
        self.colidx16 = _pinned(self.colidx.astype(np.uint16)) if (self.shape[1] <= 65536 and self.nnz) else None
        self.val = _pinned(np.ascontiguousarray(M.data, dtype=np.float32))
        self.items = _pinned(items.reshape(-1))
        self.long_rows = _pinned(long_rows.reshape(-1))
        self.n_items, self.n_long, self.n_slots = len(items), len(long_rows), int(n_slots)
        self.nbytes = sum(a.nbytes for a in (self.rowptr, self.colidx if self.colidx16 is None else self.colidx16,
                                             self.val, self.items, self.long_rows))  # bytes one upload moves


def split_hot_columns(Xl, min_density, max_cols, df=None, n_total=None):
    """Split CSR ``Xl`` into a dense block of its most frequent columns and the CSR of the rest.

    Bag-of-words columns are Zipf distributed: at C3 the 512 most frequent of 50k terms carry 63% of the
    nonzeros.  Those go to a dense N x Kh fp32 block that the tensor cores multiply (X_hot . W0[hot] forward,
    X_hot^T . dz backward); only the sparse tail pays the per-nonzero 1.2 KB gather.  A column is "hot" when its
    density (document frequency / rows) is at least ``min_density``; Kh is a multiple of 32, at most ``max_cols``.
    Returns (hot_cols int32[Kh] ascending, X_hot CSR n x Kh with hot-local column ids, X_cold CSR) or
    (None, None, Xl).  The hot block stays CSR on the host (a quarter of the dense bytes) and is expanded on the device
    (gcnb_csr_to_dense_f32).

    ``df`` / ``n_total``: document frequencies and row count of the WHOLE matrix when ``Xl`` is one rank's row block.
    Every rank then picks the same hot set, each row's sum is associated exactly as on one GPU and the row-partitioned
    forward is bit-equal to the single-GPU forward.
    """
    n, f = Xl.shape
    if df is None:
        df, n_total = np.bincount(Xl.indices, minlength=f), n
    if n_total == 0 or int(df.sum()) == 0 or max_cols < 32 or min_density <= 0:
        return None, None, Xl
    n_hot = int(np.count_nonzero(df >= min_density * n_total))
    kh = min(n_hot, int(max_cols), 65536) // 32 * 32  # hot-local ids travel as uint16
    if kh < 64:
        return None, None, Xl
    order = np.argsort(-df, kind="stable")
    hot_cols = np.sort(order[:kh]).astype(np.int32)
    hotmap = np.full(f, -1, dtype=np.int32)
    hotmap[hot_cols] = np.arange(kh, dtype=np.int32)
    m = hotmap[Xl.indices]
    is_hot = m >= 0
    rows = np.repeat(np.arange(n, dtype=np.int32), np.diff(Xl.indptr))
    hot_counts = np.bincount(rows[is_hot], minlength=n)
    hot_ptr = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(hot_counts, out=hot_ptr[1:])
    X_hot = sp.csr_matrix((Xl.data[is_hot], m[is_hot], hot_ptr), shape=(n, kh))
    cold = ~is_hot
    counts = np.bincount(rows[cold], minlength=n)
    indptr = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(counts, out=indptr[1:])
    X_cold = sp.csr_matrix((Xl.data[cold], Xl.indices[cold], indptr), shape=(n, f))
    X_cold.has_sorted_indices = bool(Xl.has_sorted_indices)  # a subsequence of sorted rows is sorted; nothing else is
    return hot_cols, X_hot, X_cold


def canonical_csr(M):
    """CSR with ascending column ids and duplicate (row, column) entries summed -- what SciPy's and Theano's products
    compute on a non-canonical matrix (the dense hot-column expansion and the sorted-plan kernels need it explicit).
    Canonical inputs are returned as they are; others are copied, never edited in place (caller-owned)."""
    M = M.tocsr()
    if not M.has_canonical_format:
        M = M.copy()
        M.sum_duplicates()
    return M


class HostGraph:
    """Everything ``Engine.bind`` uploads for one (X, A) pair and one rank: the row block of X and A,
    the transposed block of X for the backward pass, and A^T's block when A is not symmetric.
    Built once per (X, A) (the analogue of the reference's preprocess_data output staying in host
    memory across epochs, gcnmain.py:172-179) and cached by the engine."""

    def __init__(self, X, A, world, rank, chunk, need_backward, assume_symmetric=None, hot_density=0.0,
                 hot_max=0, xt_blocks=0, allreduce=None, full_graph=False, x_groups=1):
        n = X.shape[0]
        self.n = n
        self.n_pad, blocks = row_blocks(n, world)
        self.r0, self.r1 = blocks[rank]
        self.n_loc = self.r1 - self.r0
        self.n_tot = self.n_pad * world if world > 1 else n
        self.need_backward = bool(need_backward)

        def widen(M):  # gathered operands have n_pad*world rows
            if world > 1 and self.n_tot != n:
                return sp.csr_matrix((M.data, M.indices, M.indptr), shape=(M.shape[0], self.n_tot))
            return M

        self.full_graph = bool(full_graph) and world > 1
        if self.full_graph:      # feature-sliced exchange: every rank multiplies all rows of A_hat
            Xl, Al = slice_rows(X, self.r0, self.r1), A.tocsr()
        elif world > 1:
            Xl, Al = slice_rows(X, self.r0, self.r1), widen(slice_rows(A, self.r0, self.r1))
        else:
            Xl, Al = X.tocsr(), A.tocsr()
        Xl, Al = canonical_csr(Xl), canonical_csr(Al)
        df = None
        if world > 1 and hot_max >= 32 and hot_density > 0:
            # same hot set on every rank: document frequencies summed over the row blocks
            df = np.bincount(Xl.indices, minlength=Xl.shape[1]).astype(np.int64)
            df = allreduce(df) if allreduce is not None else df
        self.hot_cols, X_hot, Xl = split_hot_columns(Xl, hot_density, hot_max, df, n)
        self.kh = 0 if self.hot_cols is None else len(self.hot_cols)
        self.hot_cols_p = _pinned(self.hot_cols) if self.kh else None
        self.hot_ptr = _pinned(np.ascontiguousarray(X_hot.indptr, dtype=np.int32)) if self.kh else None
        self.hot_col = _pinned(np.ascontiguousarray(X_hot.indices, dtype=np.uint16)) if self.kh else None  # kh <= 65536
        self.hot_val = _pinned(np.ascontiguousarray(X_hot.data, dtype=np.float32)) if self.kh else None
        self.X = HostCsr(Xl, chunk, row_groups=x_groups)  # the cold columns only when a hot block exists
        self.A = HostCsr(Al, chunk)
        self.XT = self.AT = None
        self.symmetric = True
        if need_backward:
            XTl = transpose_csr(Xl)
            self.XT = HostCsr(XTl, chunk, col_blocks=xt_blocks if xt_blocks > 0 else panel_col_blocks(XTl.shape[1]))
            self.symmetric = bool(is_symmetric(A) if assume_symmetric is None else assume_symmetric)
            if not self.symmetric:
                # A^T.G for a row block needs rows r0:r1 of A^T
                AT = transpose_csr(A)
                ATl = AT if (world == 1 or self.full_graph) else widen(slice_rows(AT, self.r0, self.r1))
                self.AT = HostCsr(ATl, chunk)
        self.nbytes = sum(c.nbytes for c in (self.X, self.XT, self.A, self.AT) if c is not None)
        if self.kh:
            self.nbytes += self.hot_ptr.nbytes + self.hot_col.nbytes + self.hot_val.nbytes + self.hot_cols_p.nbytes


class DeviceCsr:
    """A CSR matrix resident in HBM together with its SpMM work plan (gcnb_csr)."""

    def __init__(self, eng, host: HostCsr, tag):
        self.shape = host.shape
        self.nnz = host.nnz
        self.n_items, self.n_long = host.n_items, host.n_long
        i32 = lambda n: torch.empty(max(n, 1), dtype=torch.int32, device=eng.dev)
        self.t_rowptr = i32(host.rowptr.size)
        self.t_colidx = i32(host.colidx.size)
        self.t_val = torch.empty(max(host.val.size, 1), dtype=torch.float32, device=eng.dev)
        self.t_items = i32(host.items.size)
        self.t_long = i32(host.long_rows.size)
        self.t_col16 = torch.empty(host.colidx16.size + 8, dtype=torch.int16, device=eng.dev) \
            if host.colidx16 is not None else None
        s = GcnbCsr()
        s.n_rows, s.n_cols, s.nnz = host.shape[0], host.shape[1], host.nnz
        s.rowptr, s.colidx, s.val = self.t_rowptr.data_ptr(), self.t_colidx.data_ptr(), self.t_val.data_ptr()
        s.items, s.n_items = self.t_items.data_ptr(), host.n_items
        s.long_rows = self.t_long.data_ptr() if host.n_long else None
        s.n_long, s.n_slots, s.tag = host.n_long, host.n_slots, int(tag)
        # gather engine: -2 = chosen per call from operand size and K (spmm.cu pick_engine); an engine may force one
        # (Engine.spmm_engine, GCNB_SPMM_ENGINE: parity tests run every engine through the whole model on small graphs)
        s.engine, s.unroll = int(getattr(eng, "spmm_engine", -2)), 0
        self.struct = s
        self.refill(eng, host)

    def refill(self, eng, host, ctx=None):
        """(Re-)copy the host arrays into the existing device buffers: async H2D on the stream of ``ctx``
        (default: the engine stream)."""
        ctx = eng.ctx if ctx is None else ctx
        for dst, src in ((self.t_rowptr, host.rowptr), (self.t_colidx, host.colidx), (self.t_val, host.val),
                         (self.t_items, host.items), (self.t_long, host.long_rows)):
            if src is host.colidx and self.t_col16 is not None:
                ctx.call("gcnb_h2d", _ptr(self.t_col16), C.c_void_p(host.colidx16.ctypes.data), host.colidx16.nbytes)
                ctx.call("gcnb_expand_u16_i32", _ptr(self.t_col16), host.colidx16.size, _ptr(self.t_colidx))
            elif src.size:
                ctx.call("gcnb_h2d", _ptr(dst), C.c_void_p(src.ctypes.data), src.nbytes)

    def refill_rows(self, eng, host, a, b, ctx, first, phase="both"):
        """Copy the nonzeros of rows [a, b) (``first``: and the row pointers and the item plan) to the device.
        ``phase``: "copy" = only the host->device copies, "expand" = only the kernel that widens 16-bit column ids
        (so that a copy stream carries nothing but copies and never queues behind compute kernels), "both"."""
        at = lambda t, off: C.c_void_p(t.data_ptr() + off)
        lo, hi = int(host.rowptr[a]), int(host.rowptr[b])
        if phase in ("copy", "both"):
            if first:
                for dst, src in ((self.t_rowptr, host.rowptr), (self.t_items, host.items), (self.t_long, host.long_rows)):
                    if src.size:
                        ctx.call("gcnb_h2d", _ptr(dst), C.c_void_p(src.ctypes.data), src.nbytes)
            if hi > lo:
                if self.t_col16 is not None:
                    ctx.call("gcnb_h2d", at(self.t_col16, 2 * lo), C.c_void_p(host.colidx16.ctypes.data + 2 * lo), 2 * (hi - lo))
                else:
                    ctx.call("gcnb_h2d", at(self.t_colidx, 4 * lo), C.c_void_p(host.colidx.ctypes.data + 4 * lo), 4 * (hi - lo))
                ctx.call("gcnb_h2d", at(self.t_val, 4 * lo), C.c_void_p(host.val.ctypes.data + 4 * lo), 4 * (hi - lo))
        if phase in ("expand", "both") and hi > lo and self.t_col16 is not None:
            lo8 = lo & ~7  # the widening kernel wants 16-byte aligned ends; the few ids before lo are already there
            ctx.call("gcnb_expand_u16_i32", at(self.t_col16, 2 * lo8), hi - lo8, at(self.t_colidx, 4 * lo8))

    def group_struct(self, host, g):
        """The same matrix restricted to the items of row group ``g`` (HostCsr.row_bounds)."""
        s = GcnbCsr()
        C.memmove(C.byref(s), C.byref(self.struct), C.sizeof(GcnbCsr))
        i0, i1 = int(host.item_bounds[g]), int(host.item_bounds[g + 1])
        s.items, s.n_items = self.t_items.data_ptr() + 16 * i0, i1 - i0
        return s

    def retagged(self, tag):
        """Same device arrays booked under another profiling tag."""
        import copy
        other = copy.copy(self)
        s = GcnbCsr()
        C.memmove(C.byref(s), C.byref(self.struct), C.sizeof(GcnbCsr))
        s.tag = int(tag)
        other.struct = s
        return other

    def engine_for(self, eng, ldb, K):
        """Gather engine (0 LDG, 1 bulk copy, 2 L2-resident panels) gcnb_spmm_csr_f32 uses for this product."""
        return int(eng.lib.gcnb_spmm_engine_for(eng.ctx.h, C.byref(self.struct), int(ldb), int(K)))

    def touched_bytes(self, K):
        """SURVEY.md 8d B_touch: nnz*(4+4) + (rows+1)*4 + nnz*K*4 + rows*K*4."""
        return self.nnz * 8 + (self.shape[0] + 1) * 4 + self.nnz * K * 4 + self.shape[0] * K * 4


class PeerUnavailable(RuntimeError):
    """CUDA IPC peer memory could not be set up on every rank (the engine then uses the all-gather exchange)."""


class _RawDeviceArray:
    """``__cuda_array_interface__`` holder: lets torch view memory the library allocated (the IPC arena)."""

    def __init__(self, ptr, shape, typestr, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}
        self._owner = owner


class PeerArena:
    """One rank's arena of NVLink peer memory (csrc/peer.cu): ``gcnb_peer_alloc`` + IPC handle exchange + ``gcnb_peer_open``
    of every peer's arena + ``gcnb_peer_setup``.  All ranks must create arenas of the same size in the same order and
    carve them identically (the engine does: identical buffer plans), so that equal offsets mean equal buffers.
    Collective: every rank of ``eng.group`` must call it together."""

    FLAG_BYTES = 256

    def __init__(self, eng, nbytes):
        import torch.distributed as dist
        # no reference to the engine itself (it owns this arena: a cycle would keep both alive until a gc pass)
        self.ctx, self.dev, self.rank, self.world = eng.ctx, eng.dev, eng.rank, eng.world
        group = eng.group
        eng = None
        self.nbytes = int(-(-(int(nbytes) + self.FLAG_BYTES + 4096) // 4096) * 4096)
        self.ptr = C.c_void_p()
        self.peers = []
        self.off = self.FLAG_BYTES
        handle = (C.c_ubyte * 64)()
        ok = 1
        try:
            self.ctx.call("gcnb_peer_alloc", self.nbytes, C.byref(self.ptr), handle)
        except capi.GcnbError as e:
            self.err = str(e)
            ok = 0
        # sizes must agree on every rank, and every rank must have its memory
        info = torch.tensor([ok, -ok, self.nbytes, -self.nbytes], dtype=torch.int64, device=self.dev)
        dist.all_reduce(info, op=dist.ReduceOp.MIN, group=group)
        info = info.tolist()
        if info[0] == 0 or info[2] != -info[3]:
            self._free_local()
            raise PeerUnavailable("peer arena: allocation failed on a rank or sizes differ (%s)" % getattr(self, "err", info))
        mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8, device=self.dev)
        allh = torch.empty(64 * self.world, dtype=torch.uint8, device=self.dev)
        dist.all_gather_into_tensor(allh, mine, group=group)
        allh = bytes(allh.cpu().numpy().tobytes())
        bases = []
        ok = 1
        for q in range(self.world):
            if q == self.rank:
                bases.append(self.ptr.value)
                continue
            pp = C.c_void_p()
            buf = (C.c_ubyte * 64).from_buffer_copy(allh[64 * q:64 * q + 64])
            try:
                self.ctx.call("gcnb_peer_open", buf, C.byref(pp))
                self.peers.append(pp)
                bases.append(pp.value)
            except capi.GcnbError as e:
                self.err = str(e)
                ok = 0
                break
        flag = torch.tensor([ok], dtype=torch.int64, device=self.dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            self.close()
            raise PeerUnavailable("peer arena: cudaIpcOpenMemHandle failed on a rank (%s)" % getattr(self, "err", ""))
        table = (C.c_void_p * self.world)(*bases)
        self.ctx.call("gcnb_peer_setup", self.rank, self.world, table, self.nbytes, 0)
        dist.barrier(group=group)  # every rank's flag block is zeroed and registered before the first device barrier

    def take(self, rows, ld, dtype=torch.float32):
        """Carve a zero-filled (rows x ld) fp32 matrix (256-byte aligned) out of the arena as a torch view."""
        nbytes = int(rows) * int(ld) * 4
        off = self.off
        self.off = -(-(off + nbytes) // 256) * 256
        if self.off > self.nbytes:
            raise capi.GcnbError("peer arena exhausted: %d of %d bytes" % (self.off, self.nbytes))
        holder = _RawDeviceArray(self.ptr.value + off, (int(rows), int(ld)), "<f4", self)
        return torch.as_tensor(holder, device=self.dev)

    def _free_local(self):
        if self.ptr.value:
            self.ctx.call("gcnb_peer_free", self.ptr)
            self.ptr = C.c_void_p()

    def close(self):
        """Collective in spirit: every rank unmaps its peers before any rank frees (callers barrier around it)."""
        try:
            self.ctx.call("gcnb_peer_setup", 0, 0, None, 0, 0)
            for pp in self.peers:
                self.ctx.call("gcnb_peer_close", pp)
            self.peers = []
        finally:
            self._free_local()


class Engine:
    """One GPU's share of the GCN: weights (replicated), row block of the graph, activations."""

    SLICE_PANEL_BYTES_MAX = 80 << 20  # same L2 rule as the panel SpMM engine (spmm.cu pick_engine)

    def __init__(self, layout: ParamLayout, drop_out=0.0, regul_coef=0.0, nonlin="tanh", device=None,
                 group=None, spmm_chunk=SPMM_CHUNK_DEFAULT, keep_logits=False, hot_density=None, hot_max=None):
        if not torch.cuda.is_available():
            raise capi.GcnbError("geographconv_b200 needs a B200 GPU: torch.cuda.is_available() is False "
                                 "(there is no CPU fallback)")
        self.layout = layout
        self.drop_out = float(drop_out)
        self.regul_coef = float(regul_coef)
        self.act = ACT[nonlin]
        self.spmm_chunk = int(spmm_chunk)
        self.keep_logits = keep_logits
        # dense hot-column block of X (split_hot_columns): columns at least this dense, at most hot_max of them
        # (measured on C3: 1024 columns at >= 5% density; 1536 columns save 1.1 ms of SpMM and cost 2.1 ms of GEMM)
        self.hot_density = float(os.environ.get("GCNB_HOT_DENSITY", "0.05") if hot_density is None else hot_density)
        self.hot_max = int(os.environ.get("GCNB_HOT_MAX", "1024") if hot_max is None else hot_max)
        # test / tuning knobs: force one SpMM gather engine for every product (-2 = per-call choice) and the number of
        # column ranges the rows of X^T are cut into (0 = from the panel working-set rule)
        self.spmm_engine = int(os.environ.get("GCNB_SPMM_ENGINE", "-2"))
        self.xt_blocks = int(os.environ.get("GCNB_XT_BLOCKS", "0"))
        # repeated uploads (cache_device_inputs = False): X crosses PCIe in this many row groups and the first layer runs
        # group by group behind the copies (1 = whole matrix first)
        self.x_groups = max(1, int(os.environ.get("GCNB_X_GROUPS", "8")))
        self.group = group
        self.world = torch.distributed.get_world_size(group) if group is not None else 1
        self.rank = torch.distributed.get_rank(group) if group is not None else 0
        # pipelined exchange (world > 1): the dense operand of a graph convolution travels in column panels of this
        # many floats; panel c+1 is all-gathered on a side stream while the SpMM already multiplies panel c (each
        # nonzero still gathers one contiguous panel row, output panels are disjoint: no extra passes).  0 = one piece.
        # Measured on C3 with 2 GPUs (profiles/r1b_multigpu.md): the narrower SpMMs and the packing copies cost more
        # than the hidden exchange saves, so the default is one piece; the exchange still runs on its own stream and
        # overlaps the backward work that does not depend on it.
        self.exchange_panel = int(os.environ.get("GCNB_EXCHANGE_PANEL", "0")) if self.world > 1 else 0
        if device is None:
            device = torch.cuda.current_device()
        self.dev = torch.device("cuda", int(device))
        torch.cuda.set_device(self.dev)
        # one dedicated (non-default) stream carries every kernel, copy and collective of the engine;
        # torch allocations / fills happen on torch's current stream and are fenced by _fence()
        self.stream = torch.cuda.Stream(self.dev)
        self.comm = torch.cuda.Stream(self.dev, priority=-1) if self.world > 1 else None
        self.ctx = capi.Context(int(device), C.c_void_p(self.stream.cuda_stream))
        # repeated host->device copies of the graph (callers that do not cache device inputs: bench.py's end-to-end
        # leg) travel on their own stream in order of first use -- X, A_hat, X^T -- so that only the copy of X is
        # exposed; the forward pass runs under the copies of A_hat and X^T (see bind / _wait_upload)
        self.copy_stream = torch.cuda.Stream(self.dev)
        self.copy_ctx = capi.Context(int(device), C.c_void_p(self.copy_stream.cuda_stream))
        # Weight-gradient products (x^T.V: a reduction over the nodes on the tensor cores) feed nothing but the optimiser,
        # so they CAN leave the critical chain of the backward pass: GCNB_SIDE_STREAM=1 runs them on a second,
        # higher-priority stream with a context (= workspace) of their own, next to the L2- / HBM-bound SpMM and
        # element-wise kernels of the main stream.  Measured on C3, one B200 (profiles/r2e_bench_c3_side{0,1}.json): 40.1 ms
        # per step with the side stream against 37.5 ms in line -- the co-resident tcgen05 CTAs take registers and L2
        # bandwidth from the latency-bound panel SpMM (A_hat products 2.25 -> 2.87 ms, X^T.dz 4.9 -> 8.2 ms) and cost more
        # than the 3.9 ms of GEMM they hide.  Off by default; kept because the trade flips if the SpMM stops being
        # L2-bound.
        self.use_side = os.environ.get("GCNB_SIDE_STREAM", "0") == "1"
        self.side_stream = torch.cuda.Stream(self.dev, priority=-1)
        self.side_ctx = capi.Context(int(device), C.c_void_p(self.side_stream.cuda_stream))
        self.side_ws = None
        self._side_readers = {}   # data_ptr of a buffer -> event after which the side stream no longer reads it
        self._uploads = {}
        self._uploads_pending = False   # bind(force_upload) on a cached graph: forward() issues the copies
        self._x_group_events = []       # one event per row group of X while those copies are in flight
        # exchange design of the graph convolutions (module docstring); "slice" needs CUDA IPC between the ranks.
        # "auto" (default) decides per bound graph: sliced while this rank's 32-column panel of the operand (N x 128 B)
        # stays L2-resident, the all-gather + row-block product otherwise (measured at N = 2M, Hd = 512 on 8 GPUs:
        # 149 ms / step gathered vs 194 ms sliced, profiles/r2b_bench_c4_n8_*.json; at N = 500k 6.6 vs 11.1 ms)
        self.exchange_pref = os.environ.get("GCNB_EXCHANGE", "auto") if self.world > 1 else "none"
        if self.exchange_pref not in ("auto", "slice", "gather", "none"):
            raise ValueError("GCNB_EXCHANGE must be 'auto', 'slice' or 'gather'")
        self.exchange = self.exchange_pref
        self.peer_ok = False
        self.arena = None
        self._slice_plans = {}
        if self.world > 1 and "GCNB_PEER_TIMEOUT_S" in os.environ:
            self.ctx.set_option("peer_timeout_s", int(os.environ["GCNB_PEER_TIMEOUT_S"]))
        if self.exchange_pref in ("slice", "auto"):
            if self.world > 16:
                self.exchange_pref = self.exchange = "gather"
            else:
                try:  # probe once: a tiny arena, mapped by every peer, one device barrier
                    probe = PeerArena(self, 1 << 16)
                    self.ctx.call("gcnb_peer_barrier")
                    self.ctx.sync()
                    torch.distributed.barrier(group=self.group)
                    probe.close()
                    torch.distributed.barrier(group=self.group)
                    self.peer_ok = True
                except PeerUnavailable as e:
                    import logging
                    logging.warning("NVLink peer memory unavailable (%s): using the all-gather exchange", e)
                    self.exchange_pref = self.exchange = "gather"
        if "GCNB_SLICED_ENGINE" in os.environ:  # gather engine of the feature-sliced product (default: by size)
            self.ctx.set_option("spmm_sliced_engine", int(os.environ["GCNB_SLICED_ENGINE"]))
        if "GCNB_SPMM_PANEL" in os.environ:  # column-panel width of the panel engine (16 / 32 / 64 floats)
            self.ctx.set_option("spmm_panel", int(os.environ["GCNB_SPMM_PANEL"]))
        if self.world > 1 and "GCNB_SM_MARGIN" in os.environ:
            # SMs the persistent SpMM kernel leaves to concurrently running NCCL kernels
            self.ctx.set_option("sm_margin", int(os.environ["GCNB_SM_MARGIN"]))
        self.lib = self.ctx.lib
        L = layout
        self.params = torch.zeros(L.total, dtype=torch.float32, device=self.dev)
        self.grads = torch.zeros(L.total, dtype=torch.float32, device=self.dev)
        self.adam_m = torch.zeros(L.total, dtype=torch.float32, device=self.dev)
        self.adam_v = torch.zeros(L.total, dtype=torch.float32, device=self.dev)
        self.adam_state = torch.zeros(2, dtype=torch.float32, device=self.dev)
        # metrics: train {loss_sum, n_correct}, dev {loss_sum, n_correct}, reg_sum, pad
        self.metrics = torch.zeros(8, dtype=torch.float32, device=self.dev)
        self.metrics_sum = torch.zeros(8, dtype=torch.float32, device=self.dev)
        self.metrics_host = torch.zeros(8, dtype=torch.float32).pin_memory()
        self._fence()
        self.ws = None
        self._measuring = False
        self._measured = 0
        self.time_nccl = False
        self._nccl_events = []
        self._keepalive = []
        self.n = None  # rows bound (global)
        self.A = self.X = self.XT = self.AT = None
        self._bound_key = None
        self._bound_refs = None
        self.host = None
        self._idx_cache = {}
        self._row_labels = {}
        self.h2d_bytes_last_bind = 0
        self.step_count = 0

    # ------------------------------------------------------------------ small helpers
    def upload(self, arr):
        """Host ndarray -> fresh device tensor through gcnb_h2d (async on the engine stream)."""
        arr = np.ascontiguousarray(arr)
        if arr.dtype == np.float32:
            t = torch.empty(arr.size, dtype=torch.float32, device=self.dev)
        elif arr.dtype == np.int32:
            t = torch.empty(arr.size, dtype=torch.int32, device=self.dev)
        elif arr.dtype == np.uint8:
            t = torch.empty(arr.size, dtype=torch.uint8, device=self.dev)
        else:
            raise TypeError(arr.dtype)
        if arr.size:
            self.ctx.call("gcnb_h2d", _ptr(t), C.c_void_p(arr.ctypes.data), arr.nbytes)
            # pageable sources are staged by the driver before cudaMemcpyAsync returns; pinned
            # sources must outlive the copy -- keep a reference until the next sync
            self._keepalive.append(arr)
        return t

    def _ensure_ws(self, nbytes):
        nbytes = int(nbytes)
        if self.ws is None or self.ws.numel() < nbytes:
            self.ctx.sync()
            self.ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=self.dev)
            self.ctx.call("gcnb_set_workspace", _ptr(self.ws), self.ws.numel())

    def _pptr(self, name):
        e = self.layout.by_name[name]
        return C.c_void_p(self.params.data_ptr() + 4 * e["offset"]), e["ld"]

    def _gptr(self, name):
        e = self.layout.by_name[name]
        return C.c_void_p(self.grads.data_ptr() + 4 * e["offset"]), e["ld"]

    def _zeros(self, rows, ld):
        rows = max(int(rows), 1)
        if self._measuring:  # sizing pass of _alloc_buffers: count what the arena must hold
            self._measured += -(-(rows * int(ld) * 4) // 256) * 256
            return None
        if self.arena is not None:
            return self.arena.take(rows, ld)
        return torch.zeros((rows, ld), dtype=torch.float32, device=self.dev)

    def _release_arena(self):
        if self.arena is not None:
            self.ctx.sync()
            torch.distributed.barrier(group=self.group)  # nobody is still storing into a peer
            self.arena.close()
            torch.distributed.barrier(group=self.group)
            self.arena = None

    def close(self):
        """Release the peer arena (collective when world > 1).  Called by __del__ as a last resort."""
        try:
            if self.arena is not None:
                self.ctx.sync()
                self.arena.close()
                self.arena = None
        except Exception:
            pass

    def __del__(self):
        self.close()

    def _fence(self):
        """Order torch-side fills (current stream) before engine-stream kernels that use the buffers."""
        self.stream.wait_stream(torch.cuda.current_stream(self.dev))

    # ------------------------------------------------------------------ parameters
    def set_params(self, params):
        flat = self.layout.pack(params)
        self.ctx.call("gcnb_h2d", _ptr(self.params), C.c_void_p(flat.ctypes.data), flat.nbytes)
        self.ctx.sync()

    def get_params(self):
        flat = np.empty(self.layout.total, dtype=np.float32)
        self.ctx.call("gcnb_d2h", C.c_void_p(flat.ctypes.data), _ptr(self.params), flat.nbytes)
        self.ctx.sync()
        return self.layout.unpack(flat)

    def get_grads(self):
        flat = np.empty(self.layout.total, dtype=np.float32)
        self.ctx.call("gcnb_d2h", C.c_void_p(flat.ctypes.data), _ptr(self.grads), flat.nbytes)
        self.ctx.sync()
        return self.layout.unpack(flat)

    # ------------------------------------------------------------------ graph / features
    def bind(self, X, A, need_backward=True, force_upload=False, assume_symmetric=None):
        """Make CSR ``X`` (N x F) and ``A`` (N x N) resident (row block of this rank) and size buffers.

        The reference passes host SciPy matrices on every f_train / f_val call
        (gcnmain.py:221,226,231).  Host preparation (row block, transpose, plan, pinned copy) and
        the device copies are cached on object identity + (shape, nnz): only the first call pays
        for them.  ``force_upload`` repeats the host->device copies (bench.py's end-to-end leg).
        """
        if not sp.issparse(X) or not sp.issparse(A):
            raise ValueError("Input for this layer must be sparse")  # gcnmodel.py:34-36
        # identity + shape + nnz + a strided content fingerprint; the engine also keeps references to the bound
        # objects, so an id cannot be handed to a new matrix while its device copy is cached
        key = (id(X), X.shape, X.nnz, id(A), A.shape, A.nnz, _fingerprint(X), _fingerprint(A))
        same = self._bound_key == key and (self.host.need_backward or not need_backward)
        if same and not force_upload:
            return
        if not same:
            if X.shape[1] != self.layout.input_size:
                raise ValueError("X has %d columns, model input_size is %d" % (X.shape[1], self.layout.input_size))
            if A.shape[0] != A.shape[1] or A.shape[0] != X.shape[0]:
                raise ValueError("A must be N x N with N = X.shape[0]")
            self.ctx.sync()
            self.copy_ctx.sync()
            self._uploads = {}
            self._uploads_pending = False
            self._x_group_events = []
            if self.exchange_pref == "auto":
                self.exchange = "slice" if self.peer_ok and X.shape[0] * 128 <= self.SLICE_PANEL_BYTES_MAX else "gather"
            hg = HostGraph(X, A, self.world, self.rank, self.spmm_chunk, need_backward, assume_symmetric,
                           self.hot_density, self.hot_max, self.xt_blocks, allreduce=self._allreduce_host,
                           full_graph=self.exchange == "slice", x_groups=self.x_groups)
            self.host = hg
            self._bound_refs = (X, A)
            self.n, self.n_pad, self.r0, self.r1 = hg.n, hg.n_pad, hg.r0, hg.r1
            self.n_loc, self.n_tot, self.symmetric = hg.n_loc, hg.n_tot, hg.symmetric
            self.X = DeviceCsr(self, hg.X, capi.TAG_SPMM_X)
            self.A = DeviceCsr(self, hg.A, capi.TAG_SPMM_A)
            self.XT = DeviceCsr(self, hg.XT, capi.TAG_SPMM_XT) if hg.XT is not None else None
            self.AT = DeviceCsr(self, hg.AT, capi.TAG_SPMM_A) if hg.AT is not None else None
            self.kh = hg.kh
            self.X_hot = self.hot_idx = None
            if hg.kh:
                self.X_hot = torch.empty((max(hg.n_loc, 1), hg.kh), dtype=torch.float32, device=self.dev)
                self.hot_idx = torch.empty(hg.kh, dtype=torch.int32, device=self.dev)
                self.hot_csr = (torch.empty(hg.hot_ptr.size, dtype=torch.int32, device=self.dev),
                                torch.empty(max(hg.hot_col.size, 1), dtype=torch.int32, device=self.dev),
                                torch.empty(max(hg.hot_val.size, 1), dtype=torch.float32, device=self.dev))
                self.hot_col16 = torch.empty(hg.hot_col.size + 8, dtype=torch.int16, device=self.dev)
                self._upload_hot(hg)
            self._alloc_buffers(need_backward)
            self._bound_key = key
            self._idx_cache = {}
            self._row_labels = {}
        else:
            hg = self.host
            # same host objects, fresh copies: issued by forward() (its first consumer), after the caller's small index /
            # label copies.  One DMA engine serves every host->device copy of the process in issue order: a 400 KB index
            # copy queued on the engine stream BEHIND these 1.4 GB would hold the whole step back until they are through.
            self._uploads_pending = True
        self.A_out = self.A.retagged(capi.TAG_SPMM_A_NARROW)
        self.AT_out = self.AT.retagged(capi.TAG_SPMM_A_NARROW) if self.AT is not None else None
        self.h2d_bytes_last_bind = hg.nbytes

    def _issue_uploads(self):
        """Repeat the host->device copies of the bound graph on the copy stream, in order of first use, one event per
        group (bind with force_upload; bench.py's end-to-end leg)."""
        self._uploads_pending = False
        hg = self.host
        self.copy_stream.wait_stream(self.stream)  # kernels of the previous call are done with the buffers
        self._x_group_events = []
        for name, pairs in (("X", ((self.X, hg.X),)), ("A", ((self.A, hg.A), (self.AT, hg.AT))),
                            ("XT", ((self.XT, hg.XT),))):
            if name == "X" and len(hg.X.row_bounds) > 2:
                # X is what the first kernel of the step needs and the largest upload: it travels in row groups,
                # one event per group, and forward() runs the first layer group by group behind the copies
                rb = hg.X.row_bounds
                for g in range(len(rb) - 1):
                    # copies only: the kernels that widen the 16-bit ids and expand the hot block run on the engine
                    # stream in forward(), else they would queue behind the compute kernels and stall the copies
                    self.X.refill_rows(self, hg.X, rb[g], rb[g + 1], self.copy_ctx, first=(g == 0), phase="copy")
                    if hg.kh:
                        self._upload_hot(hg, self.copy_ctx, rows=(rb[g], rb[g + 1]), first=(g == 0), phase="copy")
                    ev = torch.cuda.Event()
                    ev.record(self.copy_stream)
                    self._x_group_events.append(ev)
                continue
            for d, h in pairs:
                if d is not None:
                    d.refill(self, h, self.copy_ctx)
            if name == "X" and hg.kh:
                self._upload_hot(hg, self.copy_ctx)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
            self._uploads[name] = ev

    def unbind(self):
        """Drop the cache key of the bound (X, A): the next ``bind`` prepares and uploads them again."""
        self.ctx.sync()
        self.copy_ctx.sync()
        self._bound_key = None
        self._bound_refs = None

    def _allreduce_host(self, arr):
        """Sum a small host int64 array over the ranks (NCCL moves device memory only)."""
        if self.world == 1:
            return arr
        t = torch.from_numpy(np.ascontiguousarray(arr)).to(self.dev)
        torch.distributed.all_reduce(t, group=self.group)
        return t.cpu().numpy()

    def _upload_hot(self, hg, ctx=None, rows=None, first=True, phase="both"):
        ctx = self.ctx if ctx is None else ctx
        if rows is not None:  # rows [a, b) of the hot block only (``first``: and the row pointers / column list)
            a, b = rows
            at = lambda t, off: C.c_void_p(t.data_ptr() + off)
            lo, hi = int(hg.hot_ptr[a]), int(hg.hot_ptr[b])
            if phase in ("copy", "both"):
                if first:
                    ctx.call("gcnb_h2d", _ptr(self.hot_csr[0]), C.c_void_p(hg.hot_ptr.ctypes.data), hg.hot_ptr.nbytes)
                    ctx.call("gcnb_h2d", _ptr(self.hot_idx), C.c_void_p(hg.hot_cols_p.ctypes.data), hg.hot_cols_p.nbytes)
                if hi > lo:
                    ctx.call("gcnb_h2d", at(self.hot_col16, 2 * lo), C.c_void_p(hg.hot_col.ctypes.data + 2 * lo), 2 * (hi - lo))
                    ctx.call("gcnb_h2d", at(self.hot_csr[2], 4 * lo), C.c_void_p(hg.hot_val.ctypes.data + 4 * lo), 4 * (hi - lo))
            if phase in ("expand", "both"):
                if hi > lo:
                    lo8 = lo & ~7
                    ctx.call("gcnb_expand_u16_i32", at(self.hot_col16, 2 * lo8), hi - lo8, at(self.hot_csr[1], 4 * lo8))
                if b > a:
                    ctx.call("gcnb_csr_to_dense_f32", at(self.hot_csr[0], 4 * a), _ptr(self.hot_csr[1]), _ptr(self.hot_csr[2]),
                             b - a, hg.kh, at(self.X_hot, 4 * a * hg.kh), hg.kh)
            return
        # the hot block travels as CSR (hot-local column ids) and is expanded to the dense N x Kh operand on the device
        for dst, src in zip(self.hot_csr, (hg.hot_ptr, hg.hot_col, hg.hot_val)):
            if src is hg.hot_col and src.size:  # uint16 over PCIe, widened on the device
                ctx.call("gcnb_h2d", _ptr(self.hot_col16), C.c_void_p(src.ctypes.data), src.nbytes)
                ctx.call("gcnb_expand_u16_i32", _ptr(self.hot_col16), src.size, _ptr(dst))
            elif src.size:
                ctx.call("gcnb_h2d", _ptr(dst), C.c_void_p(src.ctypes.data), src.nbytes)
        ctx.call("gcnb_csr_to_dense_f32", _ptr(self.hot_csr[0]), _ptr(self.hot_csr[1]), _ptr(self.hot_csr[2]),
                 hg.n_loc, hg.kh, _ptr(self.X_hot), hg.kh)
        ctx.call("gcnb_h2d", _ptr(self.hot_idx), C.c_void_p(hg.hot_cols_p.ctypes.data), hg.hot_cols_p.nbytes)

    def _wait_upload(self, *names):
        """Engine stream waits for the in-flight host->device copies of these groups ("X", "A", "XT")."""
        for name in names:
            ev = self._uploads.pop(name, None)
            if ev is not None:
                self.stream.wait_event(ev)

    def _alloc_buffers(self, need_backward):
        if self.exchange == "slice":
            # every dense buffer of the step lives in the NVLink peer arena (peers store result rows and operand slices
            # straight into them); sizing pass first, then the arena, then the real carve -- identical on every rank
            self._release_arena()
            self.lay = []  # drop views of the old arena
            self._measuring, self._measured = True, 0
            self._alloc_buffers_impl(need_backward)
            self._measuring = False
            self.arena = PeerArena(self, self._measured + (1 << 20))
        else:
            self._release_arena()
        self._alloc_buffers_impl(need_backward)

    def _slice_plan(self, K):
        """(col0, width, ldp) int32 arrays of the feature-sliced product at operand width K (partition.slice_columns)."""
        pl = self._slice_plans.get(int(K))
        if pl is None:
            pl = self._slice_plans[int(K)] = slice_columns(int(K), self.world)
        return pl

    def _alloc_buffers_impl(self, need_backward):
        L = self.layout
        n = self.n_pad if self.world > 1 else self.n_loc
        hd = L.hid[0]
        widths = [hd] + [l["n_out"] for l in L.layers]
        self.ldh = [ld_of(w) for w in widths]
        self.ldc = ld_of(L.output_size)
        self.nbuf = n
        self.H0 = self._zeros(n, self.ldh[0])
        self.lay = []
        for i, l in enumerate(L.layers):
            b = {"Y": self._zeros(n, self.ldh[i + 1])}
            if l["kind"] == "hw":
                b["H"] = self._zeros(n, self.ldh[i + 1])
                b["T"] = self._zeros(n, self.ldh[i + 1])
            self.lay.append(b)
        maxld = max(self.ldh + [self.ldc])
        self.S = self._zeros(n, maxld)  # A.x scratch (highway) / x.W scratch (plain, output)
        self.P = self._zeros(n, self.ldc)
        self.logits = self._zeros(n, self.ldc) if self.keep_logits else None
        self.gath = self.pack = self.XP = None
        if self.exchange == "slice":
            # panel buffer: this rank's column slice of the operand, all rows (peers push their row blocks into it)
            ldp_max = max(int(self._slice_plan(w)[2].max()) for w in set(widths + [L.output_size]))
            self.XP = self._zeros(self.n_tot, ldp_max)
        elif self.world > 1:
            self.gath = self._zeros(self.n_tot, maxld)   # gathered panels, laid out one after the other
            self.pack = self._zeros(n, maxld)            # this rank's panels made contiguous for the collective
        if need_backward:
            self.G = self._zeros(n, self.ldc)
            self.U = self._zeros(n, maxld)
            self.U2 = self._zeros(n, maxld) if self.use_side else self.U  # V alternates between U2 and U: the wgrad of
            self.dX = self._zeros(n, maxld)  # one layer (side stream) may still read its V while the next V is written
            self.dH = self._zeros(n, maxld)
            self.dT = self._zeros(n, maxld)
        # workspace: the largest scratch any op of the step needs
        need = 1 << 20
        ka = max(widths + [L.output_size])
        if self.exchange == "slice":  # a rank multiplies its column slice only
            ka = max(int(self._slice_plan(w)[1][self.rank]) for w in set(widths + [L.output_size]))
        for a in (self.A, self.AT):
            if a is not None:
                need = max(need, self.lib.gcnb_spmm_workspace_bytes(C.byref(a.struct), max(ka, 4)))
        need = max(need, self.lib.gcnb_spmm_workspace_bytes(C.byref(self.X.struct), hd))
        need = max(need, self.lib.gcnb_highway_workspace_bytes(0, max(widths)))
        wall = max(widths + [L.output_size, hd])
        need = max(need, 2 * self.lib.gcnb_gemm_workspace_bytes(0, max(n, 1), wall, wall))
        self.W0_hot = None
        if self.kh:
            self.W0_hot = self._zeros(self.kh, self.ldh[0])   # W0[hot, :] forward, dW0[hot, :] backward
            need = max(need, self.lib.gcnb_gemm_workspace_bytes(0, max(n, 1), hd, self.kh))
            if need_backward:
                need = max(need, self.lib.gcnb_gemm_workspace_bytes(1, self.kh, hd, max(n, 1)))
        if self._measuring:
            return
        if need_backward:
            need = max(need, self.lib.gcnb_spmm_workspace_bytes(C.byref(self.XT.struct), hd))
            wmax = max(widths + [L.output_size])
            need = max(need, self.lib.gcnb_gemm_workspace_bytes(1, wmax, wmax, max(n, 1)))
            need = max(need, 2 * self.lib.gcnb_colsum_workspace_bytes(n, wmax))
        self._ensure_ws(need)
        if need_backward and self.use_side:
            wmax = max(widths + [L.output_size])
            side_need = self.lib.gcnb_gemm_workspace_bytes(1, wmax, wmax, max(n, 1))
            if self.kh:
                side_need = max(side_need, self.lib.gcnb_gemm_workspace_bytes(1, self.kh, hd, max(n, 1)))
            side_need = max(int(side_need), 1 << 20)
            if self.side_ws is None or self.side_ws.numel() < side_need:
                self.side_ctx.sync()
                self.side_ws = torch.empty(side_need, dtype=torch.uint8, device=self.dev)
                self.side_ctx.call("gcnb_set_workspace", _ptr(self.side_ws), self.side_ws.numel())
            self.side_stream.wait_stream(torch.cuda.current_stream(self.dev))
        self._fence()

    def index_arrays(self, idx, labels=None, force_upload=False):
        """Global node indices (+ their labels, aligned with ``idx``) -> device int32 arrays of this
        rank's local rows.  Cached on the host arrays' identity and content."""
        idx = np.asarray(idx)
        if labels is not None:
            labels = np.asarray(labels)
            if len(labels) != len(idx):
                raise AssertionError("inputs and targets differ in length")  # gcnmodel.py:304
        # keyed on the content (index sets are a few MB at most): fit() builds Y[train_indices] afresh on every call and
        # malloc hands the same address to different label vectors, so addresses identify nothing
        key = (idx.shape, idx.dtype.str, hash(idx.tobytes()), None if labels is None else hash(labels.tobytes()))
        hit = self._idx_cache.get(key)
        if hit is not None and np.array_equal(hit[3], idx) and (labels is None or np.array_equal(hit[6], labels)):
            if force_upload:
                for dst, src in ((hit[0], hit[4]), (hit[1], hit[5])):
                    if dst is not None and src.size:
                        self.ctx.call("gcnb_h2d", _ptr(dst), C.c_void_p(src.ctypes.data), src.nbytes)
            return hit[0], hit[1], hit[2]
        have_labels = labels is not None
        if labels is None:
            labels = np.zeros(len(idx), np.int32)
        if self.world > 1:
            li, ll = local_index_split(idx, labels, self.r0, self.r1)
        else:
            li, ll = idx.astype(np.int32), np.asarray(labels).astype(np.int32)
        li, ll = _pinned(li), _pinned(ll)
        d_idx = self.upload(li)
        d_lab = self.upload(ll) if have_labels else None
        if have_labels and len(np.unique(li)) == len(li):
            # dense label map for the one-pass loss gradient (gcnb_xent_grad_dense_f32); an index set that names a row
            # twice counts it twice in the reference's mean (gcnmodel.py:376,382) and keeps the scatter-add kernel
            rl = np.full(max(self.nbuf, 1), -1, dtype=np.int32)
            rl[li] = ll
            self._row_labels[d_idx.data_ptr()] = self.upload(rl)
        self.ctx.sync()
        self._keepalive = []
        if len(self._idx_cache) >= 16:
            old = self._idx_cache.pop(next(iter(self._idx_cache)))
            self._row_labels.pop(old[0].data_ptr(), None)
        self._idx_cache[key] = (d_idx, d_lab, len(li), idx.copy(), li, ll, labels.copy() if have_labels else None)
        return d_idx, d_lab, len(li)

    # ------------------------------------------------------------------ ops
    def _spmm(self, csr, B, ldb, Cbuf, ldc, K, bias=None, act=0, softmax=0, accumulate=0, dropout_p=0.0,
              seed=0, logits=None):
        epi = GcnbEpilogue()
        epi.bias = bias.value if isinstance(bias, C.c_void_p) else bias
        epi.act, epi.softmax, epi.accumulate = int(act), int(softmax), int(accumulate)
        epi.dropout_p, epi.seed, epi.row0 = float(dropout_p), int(seed) & (2**64 - 1), int(self.r0)
        epi.logits = logits.data_ptr() if logits is not None else None
        Bp = B if isinstance(B, C.c_void_p) else _ptr(B)
        Cp = Cbuf if isinstance(Cbuf, C.c_void_p) else _ptr(Cbuf)
        self.ctx.call("gcnb_spmm_csr_f32", C.byref(csr.struct), Bp, ldb, Cp, ldc, K, C.byref(epi))

    def _gemm(self, tA, tB, M, N, K, A, lda, B, ldb, Cm, ldc, accumulate=0, bias=None, act=0):
        def p(x):
            return x if isinstance(x, C.c_void_p) or x is None else _ptr(x)
        if M == 0:
            return
        self.ctx.call("gcnb_gemm_f32", tA, tB, M, N, K, p(A), lda, p(B), ldb, p(Cm), ldc, accumulate, p(bias), act)

    def _panels(self, K, ld, whole):
        """Column panels (first column, width in floats, valid columns) an operand of K columns travels in."""
        w = self.exchange_panel
        if whole or w <= 0 or w >= ld:
            return [(0, ld, K)]
        return [(c0, min(w, ld - c0), min(K - c0, min(w, ld - c0))) for c0 in range(0, ld, w) if c0 < K]

    # ---- side stream: weight-gradient GEMMs off the critical chain ----
    def _side_wgrad(self, M, N, K, A, lda, B, ldb, Cg, ldc, reads):
        """Cg[M x N] = A^T . B (A: K x M, B: K x N; K = rows of this rank) on the side stream, after everything the main
        stream has enqueued so far.  ``reads``: the device tensors it reads that the main stream will overwrite later
        (``_before_write`` makes the main stream wait for this product first)."""
        if K == 0:
            return
        if not self.use_side:
            self._gemm(1, 0, M, N, K, A, lda, B, ldb, Cg, ldc)
            return
        ev = torch.cuda.Event()
        ev.record(self.stream)
        self.side_stream.wait_event(ev)
        p = lambda x: x if isinstance(x, C.c_void_p) or x is None else _ptr(x)
        self.side_ctx.call("gcnb_gemm_f32", 1, 0, M, N, K, p(A), lda, p(B), ldb, p(Cg), ldc, 0, None, 0)
        done = torch.cuda.Event()
        done.record(self.side_stream)
        for t in reads:
            self._side_readers[t.data_ptr()] = done

    def _before_write(self, *bufs):
        """Main stream: wait until the side stream has finished reading these buffers (no-op when it never did)."""
        for t in bufs:
            ev = self._side_readers.pop(t.data_ptr(), None)
            if ev is not None:
                self.stream.wait_event(ev)

    def _join_side(self):
        if self.use_side:
            self.stream.wait_stream(self.side_stream)
            self._side_readers = {}

    # per-op profiling over both contexts (bench.py)
    def prof_enable(self, on=True):
        self.ctx.prof_enable(on)
        self.side_ctx.prof_enable(on)

    def prof_mask(self, mask):
        self.ctx.set_option("prof_mask", mask)
        self.side_ctx.set_option("prof_mask", mask)

    def prof_reset(self):
        self.ctx.prof_reset()
        self.side_ctx.prof_reset()

    def prof_collect(self):
        a, b = self.ctx.prof_collect(), self.side_ctx.prof_collect()
        return {k: (a[k][0] + b[k][0], a[k][1] + b[k][1]) for k in a}

    def launch_count(self):
        return self.ctx.launch_count() + self.side_ctx.launch_count()

    # CUDA-event timing of the NCCL collectives (bench.py's split; off by default)
    def _nccl_tick(self, stream):
        if not self.time_nccl:
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record(stream)
        return e

    def _nccl_tock(self, t0, stream):
        if t0 is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record(stream)
            self._nccl_events.append((t0, e))

    def nccl_ms(self, reset=True):
        """Milliseconds spent inside NCCL collectives since the last reset (device time; blocks)."""
        self.ctx.sync()
        if self.comm is not None:
            self.comm.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in self._nccl_events)
        if reset:
            self._nccl_events = []
        return ms

    def _arm_push(self, K):
        """Feature-sliced runs: let the NEXT producer kernel store its K-column output straight into the owners' panel
        buffers (gcnb_push_arm).  Returns True when armed; pair with ``_push_done``."""
        if self.exchange != "slice":
            return False
        col0, width, ldp = self._slice_plan(K)
        rc = self.lib.gcnb_push_arm(self.ctx.h, _ptr(self.XP), int(K), C.c_void_p(col0.ctypes.data),
                                    C.c_void_p(width.ctypes.data), C.c_void_p(ldp.ctypes.data), int(self.r0))
        if rc == capi.E_UNSUPPORTED:
            return False
        self.ctx.check(rc)
        return True

    def _push_done(self, armed):
        """True when the producer that ran since ``_arm_push`` pushed its output (no explicit push needed)."""
        return bool(armed) and bool(self.lib.gcnb_push_consumed(self.ctx.h))

    def _conv_begin(self, x, K, whole=False, pushed=False):
        """Start moving the dense operand ``x`` (n_pad x ld, this rank's rows) of a graph convolution to the ranks that
        need it.  No-op on a single GPU.  ``slice``: push this rank's rows of every column slice into the owners' panel
        buffers (NVLink stores).  ``gather``: one all-gather per column panel on the side stream."""
        if self.world == 1:
            return x
        if self.exchange == "slice":
            if not pushed:
                col0, width, ldp = self._slice_plan(K)
                self.ctx.call("gcnb_slice_push_f32", _ptr(x), int(x.shape[1]), self.n_loc, int(self.r0), _ptr(self.XP),
                              C.c_void_p(col0.ctypes.data), C.c_void_p(width.ctypes.data), C.c_void_p(ldp.ctypes.data))
            return ("slice", int(K))
        ld = x.shape[1]
        panels = self._panels(K, ld, whole)
        srcs = []
        off = 0
        for c0, w, _ in panels:
            if len(panels) == 1:
                src = x
            else:  # make the panel contiguous
                src = self.pack.view(-1)[off * self.nbuf:(off + w) * self.nbuf].view(self.nbuf, w)
                self.ctx.call("gcnb_copy2d_f32", C.c_void_p(x.data_ptr() + 4 * c0), ld, _ptr(src), w, self.nbuf, w)
            srcs.append(src)
            off += w
        ready = torch.cuda.Event()
        ready.record(self.stream)
        self.comm.wait_event(ready)  # operand complete; earlier consumers of the gather buffers are done
        handle = []
        off = 0
        with torch.cuda.stream(self.comm):
            for (c0, w, kc), src in zip(panels, srcs):
                dst = self.gath.view(-1)[off * self.n_tot:(off + w) * self.n_tot].view(self.n_tot, w)
                t0 = self._nccl_tick(self.comm)
                torch.distributed.all_gather_into_tensor(dst, src, group=self.group)
                self._nccl_tock(t0, self.comm)
                ev = torch.cuda.Event()
                ev.record(self.comm)
                handle.append((dst, ev, c0, w, kc))
                off += w
        return handle

    def _conv_finish(self, handle, csr, out, ldo, K, bias=None, act=0, softmax=0, logits=None):
        """out = epilogue(A_hat . operand): one SpMM per arrived column panel, each writing its own output columns."""
        if self.world == 1:
            x = handle
            self._spmm(csr, x, x.shape[1], out, ldo, K, bias=bias, act=act, softmax=softmax, logits=logits)
            return
        if self.exchange == "slice":
            col0, width, ldp = self._slice_plan(K)
            q = self.rank
            epi = GcnbEpilogue()
            epi.bias = bias.value if isinstance(bias, C.c_void_p) else bias
            epi.act = int(act)
            self.ctx.call("gcnb_peer_barrier")  # every rank's rows of my slice have landed in XP
            self.ctx.call("gcnb_spmm_csr_sliced_f32", C.byref(csr.struct), _ptr(self.XP), int(ldp[q]), _ptr(out), int(ldo),
                          int(K), int(col0[q]), int(width[q]), int(self.n_pad), C.byref(epi))
            self.ctx.call("gcnb_peer_barrier")  # every rank's columns of my rows have landed in `out`
            if softmax:
                self.ctx.call("gcnb_row_softmax_f32", _ptr(out), int(ldo), self.n_loc, int(K),
                              _ptr(logits) if logits is not None else None)
            return
        for dst, ev, c0, w, kc in handle:
            self.stream.wait_event(ev)
            b = None if bias is None else C.c_void_p(bias.value + 4 * c0)
            self._spmm(csr, dst, w, C.c_void_p(out.data_ptr() + 4 * c0), ldo, kc, bias=b, act=act, softmax=softmax,
                       logits=logits)

    def _conv(self, x, csr, out, ldo, K, pushed=False, **epi):
        whole = bool(epi.get("softmax"))  # a row softmax needs every column of the row in one pass
        self._conv_finish(self._conv_begin(x, K, whole, pushed), csr, out, ldo, K, **epi)

    def conv_touched_bytes(self, K):
        """B_touch of one A_hat . H product of this rank: all columns of its rows, or -- feature-sliced --
        its column slice of all rows."""
        if self.exchange == "slice":
            w = int(self._slice_plan(K)[1][self.rank])
            return self.A.nnz * 8 + (self.A.shape[0] + 1) * 4 + self.A.nnz * w * 4 + self.A.shape[0] * w * 4
        return self.A.touched_bytes(K)

    def conv_exchange_bytes(self, K):
        """Bytes this rank sends over NVLink for one graph convolution at operand width K."""
        if self.world == 1:
            return 0
        k4 = (int(K) + 3) // 4 * 4
        if self.exchange == "slice":
            col0, width, _ = self._slice_plan(K)
            push = self.n_loc * (k4 - int(width[self.rank])) * 4
            back = (self.n - self.n_loc) * int(width[self.rank]) * 4
            return push + back
        return self.n_pad * ld_of(K) * 4 * (self.world - 1)

    # ------------------------------------------------------------------ forward
    def forward(self, train=False, seed=0, want_gates=False):
        """Fill ``self.P`` (and the per-layer buffers) for the bound graph.  Asynchronous."""
        L = self.layout
        n = self.n_loc
        hd = L.hid[0]
        W0, ldw0 = self._pptr("W0")
        b0, _ = self._pptr("b0")
        p = self.drop_out if train else 0.0
        if self._uploads_pending:
            self._issue_uploads()
        self._wait_upload("X")
        # SparseInputDenseLayer + dropout: one SpMM with bias/act/dropout fused in the epilogue; when X has a
        # dense hot-column block, X_hot . W0[hot] runs on the tensor cores first and the cold-column SpMM adds to it
        # a highway layer convolves H0 itself (S = A.H0): its column slices leave from this kernel's epilogue
        want_push = bool(L.layers and L.layers[0]["kind"] == "hw")
        group_events, self._x_group_events = self._x_group_events, []
        if group_events:
            # X is still crossing PCIe in row groups (bind with force_upload): the first layer follows the copies group
            # by group -- same kernels, same per-row arithmetic, items of one row group per launch
            rb = self.host.X.row_bounds
            x_pushed = want_push
            for g, ev in enumerate(group_events):
                self.stream.wait_event(ev)
                a, b = rb[g], rb[g + 1]
                at = lambda t, off: C.c_void_p(t.data_ptr() + off)
                self.X.refill_rows(self, self.host.X, a, b, self.ctx, first=False, phase="expand")
                if self.kh:
                    self._upload_hot(self.host, self.ctx, rows=(a, b), first=False, phase="expand")
                    if g == 0:
                        self.ctx.call("gcnb_gather_rows_f32", W0, ldw0, _ptr(self.hot_idx), self.kh, hd,
                                      _ptr(self.W0_hot), self.ldh[0])
                    self._gemm(0, 0, b - a, hd, self.kh, at(self.X_hot, 4 * a * self.kh), self.kh, self.W0_hot,
                               self.ldh[0], at(self.H0, 4 * a * self.ldh[0]), self.ldh[0])
                armed = self._arm_push(hd) if want_push else False
                part = copy.copy(self.X)
                part.struct = self.X.group_struct(self.host.X, g)
                self._spmm(part, W0, ldw0, self.H0, self.ldh[0], hd, bias=b0, act=self.act, dropout_p=p, seed=seed,
                           accumulate=2 if self.kh else 0)
                x_pushed = self._push_done(armed) and x_pushed
        else:
            if self.kh:
                self.ctx.call("gcnb_gather_rows_f32", W0, ldw0, _ptr(self.hot_idx), self.kh, hd, _ptr(self.W0_hot),
                              self.ldh[0])
                self._gemm(0, 0, n, hd, self.kh, self.X_hot, self.kh, self.W0_hot, self.ldh[0], self.H0, self.ldh[0])
            armed = self._arm_push(hd) if want_push else False
            self._spmm(self.X, W0, ldw0, self.H0, self.ldh[0], hd, bias=b0, act=self.act, dropout_p=p, seed=seed,
                       accumulate=2 if self.kh else 0)
            x_pushed = self._push_done(armed)
        x, ldx, width = self.H0, self.ldh[0], hd
        self._wait_upload("A")
        for i, l in enumerate(L.layers):
            b = self.lay[i]
            ldy = self.ldh[i + 1]
            if l["kind"] == "hw":
                k = l["i"]
                Wh, ldwh = self._pptr("Wh%d" % k)
                bh, _ = self._pptr("bh%d" % k)
                Wt, ldwt = self._pptr("Wt%d" % k)
                bt, _ = self._pptr("bt%d" % k)
                S = self.S.view(-1)[: self.nbuf * ldx].view(self.nbuf, ldx)
                self._conv(x, self.A, S, ldx, width, pushed=x_pushed)
                x_pushed = False
                self.ctx.call("gcnb_highway_fwd_f32", n, width, _ptr(S), ldx, _ptr(x), ldx, Wh, ldwh, bh, Wt, ldwt,
                              bt, self.act, _ptr(b["Y"]), ldy, _ptr(b["H"]), ldy, _ptr(b["T"]), ldy)
            else:
                k = l["i"]
                W, ldw = self._pptr("W%d" % k)
                bb, _ = self._pptr("b%d" % k)
                n_out = l["n_out"]
                # reference order (gcnmodel.py:126-133): dense product first, then A, then bias
                Q = self.S.view(-1)[: self.nbuf * ldy].view(self.nbuf, ldy)
                self._gemm(0, 0, n, n_out, width, x, ldx, W, ldw, Q, ldy)
                self._conv(Q, self.A, b["Y"], ldy, n_out, bias=bb, act=self.act)
                width = n_out
            x, ldx = b["Y"], ldy
        self.x_last, self.ld_last, self.w_last = x, ldx, width
        Wout, ldwo = self._pptr("Wout")
        bout, _ = self._pptr("bout")
        Cn = L.output_size
        Q = self.S.view(-1)[: self.nbuf * self.ldc].view(self.nbuf, self.ldc)
        self._gemm(0, 0, n, Cn, width, x, ldx, Wout, ldwo, Q, self.ldc)
        self._conv(Q, self.A_out, self.P, self.ldc, Cn, bias=bout, softmax=1, logits=self.logits)
        return self.P

    # ------------------------------------------------------------------ backward + Adam
    def backward(self, d_idx, d_lab, n_idx_local, n_train_global, seed):
        L = self.layout
        n = self.n_loc
        Cn = L.output_size
        csrT = self.A if self.AT is None else self.AT
        row_label = self._row_labels.get(d_idx.data_ptr())
        pushed = False
        if row_label is not None:  # one pass over G, column slices pushed from the same kernel in sliced runs
            armed = self._arm_push(Cn)
            self.ctx.call("gcnb_xent_grad_dense_f32", _ptr(self.P), self.ldc, Cn, n, _ptr(row_label),
                          1.0 / float(n_train_global), _ptr(self.G), self.ldc)
            pushed = self._push_done(armed)
        else:
            self.ctx.call("gcnb_xent_grad_f32", _ptr(self.P), self.ldc, Cn, self.nbuf, _ptr(d_idx), _ptr(d_lab),
                          n_idx_local, 1.0 / float(n_train_global), _ptr(self.G), self.ldc)
        U = self.U.view(-1)[: self.nbuf * self.ldc].view(self.nbuf, self.ldc)
        pending = self._conv_begin(self.G, Cn, pushed=pushed)
        x, ldx, width = self.x_last, self.ld_last, self.w_last
        gb, _ = self._gptr("bout")
        self.ctx.call("gcnb_colsum_f32", n, Cn, _ptr(self.G), self.ldc, gb, 0)  # dbout (overlaps the exchange)
        self._conv_finish(pending, self.A_out if self.AT is None else self.AT_out, U, self.ldc, Cn)
        gW, ldgw = self._gptr("Wout")
        Wout, ldwo = self._pptr("Wout")
        self._side_wgrad(width, Cn, n, x, ldx, U, self.ldc, gW, ldgw, reads=(self.U,))   # dWout = x^T.U
        dX = self.dX.view(-1)[: self.nbuf * ldx].view(self.nbuf, ldx)
        self._gemm(0, 1, n, width, Cn, U, self.ldc, Wout, ldwo, dX, ldx)        # dx = U.Wout^T
        for i in reversed(range(len(L.layers))):
            l = L.layers[i]
            b = self.lay[i]
            k = l["i"]
            xin = self.lay[i - 1]["Y"] if i > 0 else self.H0
            ldin = self.ldh[i]
            n_in, n_out = l["n_in"], l["n_out"]
            ldy = self.ldh[i + 1]
            Vbuf = self.U2 if (len(L.layers) - 1 - i) % 2 == 0 else self.U
            if l["kind"] == "hw":
                dH = self.dH.view(-1)[: self.nbuf * ldy].view(self.nbuf, ldy)
                dT = self.dT.view(-1)[: self.nbuf * ldy].view(self.nbuf, ldy)
                self._before_write(self.dT)
                gWh, ldgh = self._gptr("Wh%d" % k)
                gbh, _ = self._gptr("bh%d" % k)
                gWt, ldgt = self._gptr("Wt%d" % k)
                gbt, _ = self._gptr("bt%d" % k)
                # dHpre, dTpre, dx*(1-t) (in place over dX) and both bias gradients in one pass
                armed = self._arm_push(n_out)  # sliced runs: dHpre goes to its column owners, not to local memory
                self.ctx.call("gcnb_highway_bwd_bias_f32", n, n_out, ldy, _ptr(dX), _ptr(xin), _ptr(b["H"]),
                              _ptr(b["T"]), self.act, _ptr(dH), _ptr(dT), _ptr(dX), gbh, gbt)
                pending = self._conv_begin(dH, n_out, pushed=self._push_done(armed))
                V = Vbuf.view(-1)[: self.nbuf * ldy].view(self.nbuf, ldy)
                Wh, ldwh = self._pptr("Wh%d" % k)
                Wt, ldwt = self._pptr("Wt%d" % k)
                # dWt = x^T.dTpre needs nothing from the convolution: side stream, under the SpMM
                self._side_wgrad(n_in, n_out, n, xin, ldin, dT, ldy, gWt, ldgt, reads=(self.dT,))
                split_dgrad = self.exchange == "gather"  # a GEMM of its own keeps the all-gather of dHpre covered
                if split_dgrad:
                    self._gemm(0, 1, n, n_in, n_out, dT, ldy, Wt, ldwt, dX, ldin, accumulate=1)  # dx += dTpre.Wt^T
                self._before_write(Vbuf)
                self._conv_finish(pending, csrT, V, ldy, n_out)                     # V = A^T.dHpre
                self._side_wgrad(n_in, n_out, n, xin, ldin, V, ldy, gWh, ldgh, reads=(Vbuf,))   # dWh = x^T.V
                if split_dgrad:
                    self._gemm(0, 1, n, n_in, n_out, V, ldy, Wh, ldwh, dX, ldin, accumulate=1)   # dx += V.Wh^T
                elif n > 0:  # one pass over dx: dx += dTpre.Wt^T + V.Wh^T
                    self.ctx.call("gcnb_gemm_pair_f32", 1, n, n_in, n_out, _ptr(dT), ldy, Wt, ldwt, _ptr(V), ldy, Wh,
                                  ldwh, _ptr(dX), ldin, 1)
            else:
                dP = self.dH.view(-1)[: self.nbuf * ldy].view(self.nbuf, ldy)
                gb, _ = self._gptr("b%d" % k)
                armed = self._arm_push(n_out)
                self.ctx.call("gcnb_act_bwd_bias_f32", n, n_out, ldy, _ptr(dX), _ptr(b["Y"]), self.act, 0.0, 0, 0,
                              _ptr(dP), gb)
                pending = self._conv_begin(dP, n_out, pushed=self._push_done(armed))
                V = Vbuf.view(-1)[: self.nbuf * ldy].view(self.nbuf, ldy)
                self._before_write(Vbuf)
                self._conv_finish(pending, csrT, V, ldy, n_out)
                gW, ldgw = self._gptr("W%d" % k)
                W, ldw = self._pptr("W%d" % k)
                self._side_wgrad(n_in, n_out, n, xin, ldin, V, ldy, gW, ldgw, reads=(Vbuf,))
                dXn = self.dX.view(-1)[: self.nbuf * ldin].view(self.nbuf, ldin)
                self._gemm(0, 1, n, n_in, n_out, V, ldy, W, ldw, dXn, ldin)
                dX = dXn
        hd = L.hid[0]
        ld0 = self.ldh[0]
        dX = self.dX.view(-1)[: self.nbuf * ld0].view(self.nbuf, ld0)
        p = self.drop_out
        gW0, ldg0 = self._gptr("W0")
        gb0, _ = self._gptr("b0")
        self.ctx.call("gcnb_act_bwd_bias_f32", n, hd, ld0, _ptr(dX), _ptr(self.H0), self.act, p, int(seed) & (2**64 - 1),
                      int(self.r0), _ptr(dX), gb0)
        if self.kh:  # hot columns: dense wgrad on the tensor cores (side stream), under the cold-column SpMM
            self._side_wgrad(self.kh, hd, n, self.X_hot, self.kh, dX, ld0, self.W0_hot, self.ldh[0], reads=())
        self._wait_upload("XT")
        self._spmm(self.XT, dX, ld0, gW0, ldg0, hd)                                 # dW0 = X^T.dz (cold columns)
        self._join_side()                                                           # every weight gradient is in place
        if self.kh:
            self.ctx.call("gcnb_scatter_rows_f32", _ptr(self.W0_hot), self.ldh[0], _ptr(self.hot_idx), self.kh, hd,
                          gW0, ldg0)
        if self.world > 1:
            with torch.cuda.stream(self.stream):
                t0 = self._nccl_tick(self.stream)
                torch.distributed.all_reduce(self.grads, group=self.group)
                self._nccl_tock(t0, self.stream)
        if self.regul_coef > 0:
            reg = C.c_void_p(self.metrics.data_ptr() + 4 * 4)
            for off, size in L.weight_segments():
                self.ctx.call("gcnb_l1l2_f32", C.c_void_p(self.params.data_ptr() + 4 * off),
                              C.c_void_p(self.grads.data_ptr() + 4 * off), size, self.regul_coef, reg)

    def adam_step(self, lr=2e-3, beta1=0.9, beta2=0.999, eps=1e-8):
        self.ctx.call("gcnb_adam_f32", _ptr(self.params), _ptr(self.grads), _ptr(self.adam_m), _ptr(self.adam_v),
                      self.layout.total, _ptr(self.adam_state), lr, beta1, beta2, eps)

    def train_step(self, tr, dv, n_train, n_dev, seed, update=True):
        """One ``f_train`` call (gcnmodel.py:409): metrics of the dropout output, then Adam.

        ``tr`` / ``dv`` are ``index_arrays`` results.  Asynchronous; ``read_metrics`` syncs.
        """
        self.ctx.call("gcnb_memset", _ptr(self.metrics), 0, 32)
        self.forward(train=True, seed=seed)
        Cn = self.layout.output_size
        self.ctx.call("gcnb_xent_metrics_f32", _ptr(self.P), self.ldc, Cn, _ptr(tr[0]), _ptr(tr[1]), tr[2],
                      _ptr(self.metrics))
        if dv is not None:
            self.ctx.call("gcnb_xent_metrics_f32", _ptr(self.P), self.ldc, Cn, _ptr(dv[0]), _ptr(dv[1]), dv[2],
                          C.c_void_p(self.metrics.data_ptr() + 8))
        self.backward(tr[0], tr[1], tr[2], n_train, seed)
        if update:
            self.adam_step()
        self.step_count += 1
        self._n_train, self._n_dev = n_train, n_dev

    def read_metrics(self):
        """(train_loss, train_acc, dev_loss, dev_acc) of the last train_step; blocks."""
        src = self.metrics
        if self.world > 1:  # sum over ranks in a scratch copy: a second read of the same step must not double it
            with torch.cuda.stream(self.stream):
                self.metrics_sum.copy_(self.metrics)
                torch.distributed.all_reduce(self.metrics_sum[:4], group=self.group)
            src = self.metrics_sum
        self.ctx.call("gcnb_d2h", C.c_void_p(self.metrics_host.data_ptr()), _ptr(src), 32)
        self.ctx.sync()
        m = self.metrics_host.numpy()
        nt, nd = max(self._n_train, 1), max(self._n_dev, 1)
        loss = float(m[0]) / nt + self.regul_coef * float(m[4])
        return loss, float(m[1]) / nt, float(m[2]) / nd, float(m[3]) / nd

    # ------------------------------------------------------------------ outputs
    def gather_predictions(self, idx, want_probs=True):
        """argmax and probability rows of ``self.P`` at global indices ``idx`` (f_val outputs).  ``want_probs=False``
        returns (preds, device int64 tensor of the same predictions) and copies no probabilities to the host."""
        idx = np.ascontiguousarray(np.asarray(idx), dtype=np.int32)
        Cn = self.layout.output_size
        m = len(idx)
        P = self.P
        if self.world > 1:
            full = torch.empty((self.n_tot, self.ldc), dtype=torch.float32, device=self.dev)
            with torch.cuda.stream(self.stream):
                torch.distributed.all_gather_into_tensor(full, self.P, group=self.group)
            P = full
        self._keepalive = []
        d_idx = self.upload(idx)
        d_pred = torch.empty(max(m, 1), dtype=torch.int64, device=self.dev)
        d_prob = torch.empty((max(m, 1), Cn), dtype=torch.float32, device=self.dev) if want_probs else None
        self.ctx.call("gcnb_gather_argmax_f32", _ptr(P), self.ldc, Cn, _ptr(d_idx), m, _ptr(d_pred), _ptr(d_prob))
        preds = np.empty(m, dtype=np.int64)
        probs = np.empty((m, Cn), dtype=np.float32) if want_probs else None
        if m:
            self.ctx.call("gcnb_d2h", C.c_void_p(preds.ctypes.data), _ptr(d_pred), preds.nbytes)
            if want_probs:
                self.ctx.call("gcnb_d2h", C.c_void_p(probs.ctypes.data), _ptr(d_prob), probs.nbytes)
        self.ctx.sync()
        self._keepalive = []
        return (preds, probs) if want_probs else (preds, d_pred)

    def read_matrix(self, buf, rows, cols):
        """Device (rows_pad x ld) buffer -> host ndarray (rows x cols), gathered over ranks."""
        if self.world > 1:
            full = torch.empty((self.n_tot, buf.shape[1]), dtype=torch.float32, device=self.dev)
            with torch.cuda.stream(self.stream):
                torch.distributed.all_gather_into_tensor(full, buf.contiguous(), group=self.group)
            buf = full
        host = np.empty((buf.shape[0], buf.shape[1]), dtype=np.float32)
        self.ctx.call("gcnb_d2h", C.c_void_p(host.ctypes.data), _ptr(buf), host.nbytes)
        self.ctx.sync()
        return np.ascontiguousarray(host[:rows, :cols])

    def read_rows(self, buf, rows, cols):
        """Rows ``rows`` (global node ids) of a row-partitioned device matrix -> host (len(rows) x cols).  Collective
        when world > 1 (the row blocks are gathered on the device; only the selected rows cross PCIe).  Checker /
        diagnostics path (bench.py's in-run parity block), not part of the training step."""
        if self.world > 1:
            full = torch.empty((self.n_tot, buf.shape[1]), dtype=torch.float32, device=self.dev)
            with torch.cuda.stream(self.stream):
                torch.distributed.all_gather_into_tensor(full, buf.contiguous(), group=self.group)
            buf = full
        self.ctx.sync()
        idx = torch.as_tensor(np.asarray(rows, dtype=np.int64), device=self.dev)
        with torch.cuda.stream(self.stream):
            sel = buf.index_select(0, idx)[:, :cols].contiguous()
        self.ctx.sync()
        return sel.cpu().numpy()

    def checksum(self, buf, rows, cols):
        """Order-independent checksum of the bit patterns of buf[:rows, :cols], summed over the ranks: equal on 1 and on
        P GPUs exactly when the row-partitioned result is bit-identical to the single-GPU one."""
        self.ctx.sync()
        with torch.cuda.stream(self.stream):
            bits = buf[:rows, :cols].contiguous().view(torch.int32).to(torch.int64)
            mixed = (bits * 2654435761 + (bits >> 7)) & 0x7FFFFFFFFFFF
            tot = mixed.sum().reshape(1)
            if self.world > 1:
                torch.distributed.all_reduce(tot, group=self.group)
        self.ctx.sync()
        return int(tot.item()) & 0xFFFFFFFFFFFF

    def gates(self):
        """Gate activations T_i of the last forward, one N x Hd array per highway layer."""
        out = []
        for l, b in zip(self.layout.layers, self.lay):
            if l["kind"] == "hw":
                out.append(self.read_matrix(b["T"], self.n, l["n_out"]))
        return out

    def dropout_mask(self, seed):
        """Keep mask (N_local x Hd uint8) the fused epilogue draws for ``seed`` (tests feed the oracle)."""
        hd = self.layout.hid[0]
        m = torch.empty((max(self.n_loc, 1), hd), dtype=torch.uint8, device=self.dev)
        self.ctx.call("gcnb_dropout_mask_u8", self.n_loc, hd, self.drop_out, int(seed) & (2**64 - 1), int(self.r0),
                      _ptr(m))
        host = np.empty((self.n_loc, hd), dtype=np.uint8)
        if host.size:
            self.ctx.call("gcnb_d2h", C.c_void_p(host.ctypes.data), _ptr(m), host.nbytes)
        self.ctx.sync()
        return host
