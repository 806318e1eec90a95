"""Host-side layout helpers: leading dimensions, the flat parameter buffer, 1-D row partitioning.

Pure NumPy/SciPy -- importable without a GPU, exercised by the CPU test-suite (including the
world_size-2 gloo test of the row-partitioned exchange).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

LD_ALIGN = 32  # floats: every device matrix row starts on a 128-byte line


def round_up(x, m):
    return (int(x) + m - 1) // m * m


def ld_of(cols):
    return round_up(cols, LD_ALIGN)


class ParamLayout:
    """Flat fp32 buffer holding the weights in Lasagne ``get_all_param_values`` order.

    ``[W0, b0, (Wt_i, bt_i, Wh_i, bh_i)*, Wout, bout]`` for highway nets,
    ``[W0, b0, (W_i, b_i)*, Wout, bout]`` otherwise (gcnmodel.py:258,353-374,414).
    A (in, out) matrix is stored row-major with ld = round_up(out, 32); a bias occupies
    round_up(out, 32) floats; padding is zero, so Adam / L1+L2 over the whole buffer are exact.
    """

    def __init__(self, input_size, hid_size_list, output_size, highway):
        self.input_size = int(input_size)
        self.output_size = int(output_size)
        self.hid = [int(h) for h in hid_size_list]
        self.highway = bool(highway)
        self.entries = []
        off = 0

        def add(name, shape):
            nonlocal off
            if len(shape) == 2:
                ld = ld_of(shape[1])
                size = shape[0] * ld
            else:
                ld = ld_of(shape[0])
                size = ld
            self.entries.append(dict(name=name, shape=tuple(shape), ld=ld, offset=off, size=size))
            off += size

        hd = self.hid[0]
        add("W0", (self.input_size, hd))
        add("b0", (hd,))
        self.layers = []  # per hidden conv layer: dict(kind, in, out, names)
        prev = hd
        for i, h in enumerate(self.hid):
            if i == 0:
                continue
            if self.highway:
                add("Wt%d" % i, (prev, prev)); add("bt%d" % i, (prev,))
                add("Wh%d" % i, (prev, prev)); add("bh%d" % i, (prev,))
                self.layers.append(dict(kind="hw", i=i, n_in=prev, n_out=prev))
            else:
                add("W%d" % i, (prev, h)); add("b%d" % i, (h,))
                self.layers.append(dict(kind="gc", i=i, n_in=prev, n_out=h))
                prev = h
        self.last_width = prev
        add("Wout", (prev, self.output_size))
        add("bout", (self.output_size,))
        self.total = off
        self.by_name = {e["name"]: e for e in self.entries}

    def shapes(self):
        return [e["shape"] for e in self.entries]

    def check(self, params):
        if len(params) != len(self.entries):
            raise ValueError("expected %d parameter arrays, got %d" % (len(self.entries), len(params)))
        for p, e in zip(params, self.entries):
            if tuple(np.shape(p)) != e["shape"]:
                raise ValueError("parameter %s: expected shape %s, got %s" % (e["name"], e["shape"], np.shape(p)))

    def pack(self, params, out=None):
        self.check(params)
        flat = np.zeros(self.total, dtype=np.float32) if out is None else out
        if out is not None:
            flat[:] = 0
        for p, e in zip(params, self.entries):
            p = np.asarray(p, dtype=np.float32)
            if p.ndim == 2:
                view = flat[e["offset"]:e["offset"] + e["size"]].reshape(e["shape"][0], e["ld"])
                view[:, :e["shape"][1]] = p
            else:
                flat[e["offset"]:e["offset"] + e["shape"][0]] = p
        return flat

    def unpack(self, flat):
        out = []
        for e in self.entries:
            if len(e["shape"]) == 2:
                view = flat[e["offset"]:e["offset"] + e["size"]].reshape(e["shape"][0], e["ld"])
                out.append(np.array(view[:, :e["shape"][1]], dtype=np.float32))
            else:
                out.append(np.array(flat[e["offset"]:e["offset"] + e["shape"][0]], dtype=np.float32))
        return out

    def weight_segments(self):
        """(offset, size) of every W (not b): the tensors L1+L2 applies to (gcnmodel.py:383-387)."""
        return [(e["offset"], e["size"]) for e in self.entries if len(e["shape"]) == 2]


# ------------------------------------------------------------------------------------------
# 1-D row partition
# ------------------------------------------------------------------------------------------

def row_blocks(n, world):
    """Contiguous equal blocks: rank p owns rows [p*n_pad, min(n, (p+1)*n_pad)).

    Equal (padded) block size keeps the all-gather a single fixed-size collective and makes the
    gathered row index equal the global node id, so column indices of A need no remapping.
    """
    n_pad = (n + world - 1) // world
    return n_pad, [(min(n, p * n_pad), min(n, (p + 1) * n_pad)) for p in range(world)]


SLICE_UNIT = 16  # floats: column slices are cut at multiples of 16 columns (one 4-lane group of the panel SpMM)


def slice_columns(K, world):
    """Column slices of a K-wide dense operand for the feature-sliced graph convolution (csrc/peer.cu):
    rank q multiplies all rows of A_hat by columns [col0[q], col0[q] + width[q]).  Slices are runs of 16-column units,
    as even as possible (the first K/16 mod world ranks get one more), the last one is cut at K rounded up to 4.
    ldp[q] = leading dimension of q's panel buffer (its slice rounded up to 32 floats: rows start on 128-byte lines).
    K = 300 on 8 ranks: 19 units -> widths 48, 48, 48, 32, 32, 32, 32, 28; on 2 ranks 160 + 140.
    Returns three int32 arrays of ``world`` entries."""
    k4 = round_up(K, 4)
    units = (k4 + SLICE_UNIT - 1) // SLICE_UNIT
    base, extra = divmod(units, world)
    col0 = np.zeros(world, dtype=np.int32)
    width = np.zeros(world, dtype=np.int32)
    c = 0
    for q in range(world):
        u = base + (1 if q < extra else 0)
        col0[q] = min(c, k4)
        width[q] = max(0, min(u * SLICE_UNIT, k4 - c))
        c += u * SLICE_UNIT
    ldp = np.array([round_up(max(int(w), 4), LD_ALIGN) for w in width], dtype=np.int32)
    return col0, width, ldp


def slice_rows(M, r0, r1):
    M = M.tocsr()
    out = M[r0:r1]
    out.sort_indices()
    return out


def local_index_split(idx, labels, r0, r1):
    """Indices falling in [r0, r1) rebased to local rows, with their labels; order preserved."""
    idx = np.asarray(idx, dtype=np.int64)
    keep = (idx >= r0) & (idx < r1)
    return (idx[keep] - r0).astype(np.int32), np.asarray(labels)[keep].astype(np.int32)


def transpose_csr(M):
    """CSR of M^T with sorted indices (the operand of the X^T.dz product)."""
    T = M.tocsr().T.tocsr()
    T.sort_indices()
    return T


def is_symmetric(A, tol=1e-6):
    """A_hat of an undirected graph is symmetric (data.py:40, gcnmain.py:115-127) => A^T.G == A.G."""
    A = A.tocsr()
    if A.shape[0] != A.shape[1]:
        return False
    D = (A - A.T).tocsr()
    return D.nnz == 0 or float(np.abs(D.data).max()) <= tol
