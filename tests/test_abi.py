"""The C-ABI boundary without a GPU: the library loads, exports every symbol the header declares,
the ctypes prototypes cover exactly those symbols, and the host-side planner behaves."""
import os
import re

import numpy as np
import pytest

from geographconv_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "gcnb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gcnb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built_library):
    syms = _header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(built_library, s), "libgcnb200.so does not export %s" % s
    assert sorted(capi.SIGNATURES) == syms, "capi.SIGNATURES and include/gcnb200.h disagree"
    assert built_library.gcnb_version() == 100


def test_struct_layouts_match_header(tmp_path):
    """ctypes mirrors of gcnb_csr / gcnb_epilogue agree with what a C compiler lays out from the header."""
    import ctypes as C
    import subprocess
    fields = {"gcnb_csr": [f[0] for f in capi.GcnbCsr._fields_], "gcnb_epilogue": [f[0] for f in capi.GcnbEpilogue._fields_]}
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "gcnb200.h"', 'int main(void){']
    for st, fs in fields.items():
        src.append('printf("%s %%zu\\n", sizeof(%s));' % (st, st))
        for f in fs:
            src.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (st, f, st, f))
    src.append('return 0;}')
    cfile = tmp_path / "layout.c"
    cfile.write_text("\n".join(src))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(cfile), "-o", str(exe)])
    out = dict(line.split() for line in subprocess.check_output([str(exe)]).decode().splitlines())
    for st, cls in (("gcnb_csr", capi.GcnbCsr), ("gcnb_epilogue", capi.GcnbEpilogue)):
        assert int(out[st]) == C.sizeof(cls)
        for f in fields[st]:
            assert int(out["%s.%s" % (st, f)]) == getattr(cls, f).offset, (st, f)


def test_no_gpu_fails_loudly(built_library):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.GcnbError):
        capi.Context(0)
    from geographconv_b200.gcnmodel import GraphConv
    from geographconv_b200.synth import synthetic_problem
    A, X, Y, tr, dev, te, cfg = synthetic_problem("tiny")
    clf = GraphConv(cfg["f"], cfg["classes"], cfg["hid"], 0.0, 0.5)
    clf.build_model(A)
    with pytest.raises(capi.GcnbError):
        clf.predict(X, A, te)


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(capi.GcnbError):
        capi.load_library(str(tmp_path / "nope.so"))


def test_csr_plan_items_cover_rows(built_library):
    rowptr = np.array([0, 0, 3, 3, 1003, 1004, 1004 + 257], dtype=np.int32)
    items, long_rows, n_slots = capi.csr_plan(rowptr, 256)
    # rows 0 and 2 are empty, row 3 (1000 nnz) splits into 4 equal pieces, row 5 (257) into 2
    assert long_rows.tolist() == [[3, 0, 4], [5, 4, 2]] and n_slots == 6
    cover = {}
    for r, b, e, s in items.tolist():
        assert b <= e and e - b <= 256
        cover.setdefault(r, []).append((b, e, s))
    for r in range(len(rowptr) - 1):
        segs = sorted(cover[r])
        assert segs[0][0] == rowptr[r] and segs[-1][1] == rowptr[r + 1]
        for (b0, e0, _), (b1, e1, _) in zip(segs, segs[1:]):
            assert e0 == b1
        assert all(s == -1 for _, _, s in segs) == (len(segs) == 1)
    assert [s for _, _, s in sorted(cover[3])] == [0, 1, 2, 3]
    # empty matrix
    items, long_rows, n_slots = capi.csr_plan(np.array([0], dtype=np.int32), 256)
    assert len(items) == 0 and len(long_rows) == 0 and n_slots == 0


@pytest.mark.parametrize("n_blocks,chunk", [(1, 16), (2, 16), (2, 256), (3, 64), (4, 256)])
def test_column_blocked_plan_covers_rows_without_straddling(built_library, n_blocks, chunk):
    """plan_col_blocks: items tile every row in column order, never cross a column-range boundary, slots of a long
    row are consecutive in column order, an empty row keeps one empty item; one block == the C planner."""
    import scipy.sparse as sp
    from geographconv_b200.engine import plan_col_blocks
    rng = np.random.RandomState(0)
    n_cols = 1000
    M = sp.random(200, n_cols, density=0.05, random_state=rng, format="lil")
    M[5, :] = 1
    M[7, :] = 0
    M = M.tocsr()
    M.eliminate_zeros()
    M.sort_indices()
    items, long_rows, n_slots = plan_col_blocks(M.indptr, M.indices, n_cols, n_blocks, chunk)
    bounds = np.array([(n_cols * k + n_blocks - 1) // n_blocks for k in range(n_blocks + 1)])
    cover, slots = {}, []
    for r, b, e, s in items.tolist():
        assert 0 <= e - b <= chunk
        cover.setdefault(r, []).append((b, e, s))
        if e > b:
            blk = np.searchsorted(bounds, M.indices[b:e], side="right") - 1
            assert blk.min() == blk.max()
    for r in range(M.shape[0]):
        segs = sorted(cover[r])
        assert segs[0][0] == M.indptr[r] and segs[-1][1] == M.indptr[r + 1]
        assert all(a[1] == b[0] for a, b in zip(segs, segs[1:]))
        if len(segs) == 1:
            assert segs[0][2] == -1
        else:
            sl = [x[2] for x in segs]
            assert sl == list(range(sl[0], sl[0] + len(sl)))
            slots += sl
    assert sorted(slots) == list(range(n_slots))
    for r, s0, k in long_rows.tolist():
        assert len(cover[r]) == k and sorted(cover[r])[0][2] == s0
    if n_blocks == 1:
        it0, lr0, ns0 = capi.csr_plan(np.ascontiguousarray(M.indptr, dtype=np.int32), chunk)
        assert ns0 == n_slots and (lr0 == long_rows).all()
        assert sorted(map(tuple, it0.tolist())) == sorted(map(tuple, items.tolist()))
