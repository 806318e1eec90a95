"""Host logic: flat parameter layout, row blocks, and the row-partitioned exchange (world_size 2, gloo)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from geographconv_b200 import partition, synth
from oracle import gcn_ref


def test_param_layout_roundtrip_highway_and_plain():
    for highway, hid in [(True, [300, 300, 300]), (False, [40, 24, 56]), (True, [16])]:
        L = partition.ParamLayout(100, hid, 129, highway)
        params = gcn_ref.init_params(100, hid, 129, highway, 1)
        flat = L.pack(params)
        assert flat.size == L.total and L.total % 32 == 0
        for e in L.entries:
            assert e["offset"] % 32 == 0 and e["ld"] % 32 == 0
        back = L.unpack(flat)
        for a, b in zip(params, back):
            np.testing.assert_array_equal(a, b)
        # padding stays zero
        assert flat.sum() == pytest.approx(sum(float(p.sum(dtype=np.float64)) for p in params), rel=1e-4, abs=1e-2)
    L = partition.ParamLayout(100, [300, 300, 300], 129, True)
    assert [e["name"] for e in L.entries] == ["W0", "b0", "Wt1", "bt1", "Wh1", "bh1", "Wt2", "bt2", "Wh2", "bh2",
                                              "Wout", "bout"]
    with pytest.raises(ValueError):
        L.pack(gcn_ref.init_params(100, [300, 300], 129, True, 1))


def test_row_blocks_and_index_split():
    n_pad, blocks = partition.row_blocks(10, 4)
    assert n_pad == 3 and blocks == [(0, 3), (3, 6), (6, 9), (9, 10)]
    n_pad, blocks = partition.row_blocks(5, 8)
    assert n_pad == 1 and blocks[5] == (5, 5) and blocks[7] == (5, 5)
    idx = np.array([9, 0, 4, 3, 5], dtype=np.int32)
    lab = np.array([1, 2, 3, 4, 5], dtype=np.int32)
    li, ll = partition.local_index_split(idx, lab, 3, 6)
    assert li.tolist() == [1, 0, 2] and ll.tolist() == [3, 4, 5] and li.dtype == np.int32


def test_symmetry_and_transpose():
    A = synth.synthetic_graph(300, 6, 1)
    assert partition.is_symmetric(A)
    X = synth.synthetic_features(300, 50, 8, 1)
    assert not partition.is_symmetric(X)
    XT = partition.transpose_csr(X)
    assert XT.shape == (50, 300) and XT.has_sorted_indices
    np.testing.assert_allclose(XT.toarray(), X.toarray().T)


def _rank_main(rank, world, port, q):
    """Row-partitioned forward on CPU with the engine's exchange pattern (all-gather the dense
    operand, local rows of A), gloo backend; compared with the single-process oracle."""
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        A, X, Y, tr, dev, te, cfg = synth.synthetic_problem(dict(n=101, deg=5, f=40, xnnz=6, hid=[12, 12, 12], classes=5))
        params = gcn_ref.init_params(cfg["f"], cfg["hid"], cfg["classes"], True, 3)
        n = cfg["n"]
        n_pad, blocks = partition.row_blocks(n, world)
        r0, r1 = blocks[rank]
        n_tot = n_pad * world
        Al = partition.slice_rows(A, r0, r1)
        Al = sp.csr_matrix((Al.data, Al.indices, Al.indptr), shape=(r1 - r0, n_tot))
        Xl = partition.slice_rows(X, r0, r1)

        def gathered(x):
            loc = np.zeros((n_pad, x.shape[1]), dtype=np.float32)
            loc[: r1 - r0] = x
            out = torch.empty((n_tot, x.shape[1]), dtype=torch.float32)
            dist.all_gather_into_tensor(out, torch.from_numpy(loc))
            return out.numpy()

        W0, b0 = params[0], params[1]
        x = np.tanh(Xl @ W0 + b0)
        k = 2
        for _ in range(2):
            Wt, bt, Wh, bh = params[k:k + 4]
            k += 4
            h = np.tanh((Al @ gathered(x)) @ Wh + bh)
            t = gcn_ref.sigmoid(x @ Wt + bt)
            x = t * h + (1 - t) * x
        logits = Al @ gathered(x @ params[k]) + params[k + 1]
        probs = gcn_ref.softmax_rows(logits)
        # loss pieces and weight-gradient all-reduce pattern: local partial sums, global mean
        li, ll = partition.local_index_split(tr, Y[tr], r0, r1)
        part = torch.tensor([float(-np.log(probs[li, ll]).sum()), float(len(li))], dtype=torch.float64)
        dist.all_reduce(part)
        full = gathered(probs)[:n]
        q.put((rank, full, float(part[0] / part[1]), int(part[1])))
    finally:
        dist.destroy_process_group()


def test_row_partitioned_exchange_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    A, X, Y, tr, dev, te, cfg = synth.synthetic_problem(dict(n=101, deg=5, f=40, xnnz=6, hid=[12, 12, 12], classes=5))
    params = gcn_ref.init_params(cfg["f"], cfg["hid"], cfg["classes"], True, 3)
    ref = gcn_ref.forward(params, X, A, cfg["hid"], True)
    want_loss = gcn_ref.cross_entropy(ref["probs"][tr], Y[tr])
    for rank, full, loss, cnt in res:
        np.testing.assert_allclose(full, ref["probs"], rtol=1e-4, atol=1e-6)
        assert cnt == len(tr)
        assert abs(loss - want_loss) < 1e-4


def _p2p_exchange(dist, torch, send_blocks, recv_shapes, rank, world):
    """What the NVLink stores of csrc/peer.cu do, with gloo send/recv: block q of ``send_blocks`` lands on rank q."""
    recv = [torch.empty(sh, dtype=torch.float32) for sh in recv_shapes]
    recv[rank].copy_(torch.from_numpy(np.ascontiguousarray(send_blocks[rank])))
    reqs = []
    for q in range(world):
        if q != rank:
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(send_blocks[q])), dst=q))
            reqs.append(dist.irecv(recv[q], src=q))
    for r in reqs:
        r.wait()
    return [t.numpy() for t in recv]


def _rank_main_sliced(rank, world, port, q):
    """Feature-sliced graph convolution on CPU with the engine's exchange pattern (partition.slice_columns): push row
    blocks of every column slice to its owner, multiply ALL rows of A by the slice, send result rows home."""
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        A, X, Y, tr, dev, te, cfg = synth.synthetic_problem(dict(n=103, deg=5, f=40, xnnz=6, hid=[44, 44, 44], classes=5))
        params = gcn_ref.init_params(cfg["f"], cfg["hid"], cfg["classes"], True, 3)
        n = cfg["n"]
        n_pad, blocks = partition.row_blocks(n, world)
        r0, r1 = blocks[rank]
        Xl = partition.slice_rows(X, r0, r1)

        def conv(x):  # x: this rank's rows (r1 - r0) x K  ->  (A . x_all)[r0:r1]
            K = x.shape[1]
            col0, width, _ = partition.slice_columns(K, world)
            k4 = partition.round_up(K, 4)
            xpad = np.zeros((r1 - r0, k4), dtype=np.float32)
            xpad[:, :K] = x
            # push: my rows of slice q -> rank q
            send = [xpad[:, col0[p]:col0[p] + width[p]] for p in range(world)]
            shapes = [(blocks[p][1] - blocks[p][0], int(width[rank])) for p in range(world)]
            XP = np.concatenate(_p2p_exchange(dist, torch, send, shapes, rank, world), axis=0)   # all rows, my slice
            part = np.asarray(A @ XP, dtype=np.float32) if width[rank] else np.zeros((n, 0), np.float32)
            # result rows go home: rows of rank p, my columns -> rank p
            send = [part[blocks[p][0]:blocks[p][1]] for p in range(world)]
            shapes = [(r1 - r0, int(width[p])) for p in range(world)]
            cols = _p2p_exchange(dist, torch, send, shapes, rank, world)
            return np.concatenate(cols, axis=1)[:, :K]

        W0, b0 = params[0], params[1]
        x = np.tanh(Xl @ W0 + b0)
        first_conv = conv(x)
        k = 2
        for _ in range(2):
            Wt, bt, Wh, bh = params[k:k + 4]
            k += 4
            h = np.tanh(conv(x) @ Wh + bh)
            t = gcn_ref.sigmoid(x @ Wt + bt)
            x = t * h + (1 - t) * x
        logits = conv(x @ params[k]) + params[k + 1]
        q.put((rank, r0, r1, first_conv, gcn_ref.softmax_rows(logits)))
    finally:
        dist.destroy_process_group()


def test_feature_sliced_exchange_world3_gloo():
    """The sliced product must be BIT-identical to the single-process product (every row is summed in CSR order whatever
    the column slicing); the forward built on it matches the oracle."""
    import torch.multiprocessing as mp
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_rank_main_sliced, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    A, X, Y, tr, dev, te, cfg = synth.synthetic_problem(dict(n=103, deg=5, f=40, xnnz=6, hid=[44, 44, 44], classes=5))
    params = gcn_ref.init_params(cfg["f"], cfg["hid"], cfg["classes"], True, 3)
    x0 = np.tanh(X @ params[0] + params[1])
    want_first = np.asarray(A @ x0, dtype=np.float32)
    ref = gcn_ref.forward(params, X, A, cfg["hid"], True)
    for rank, r0, r1, first_conv, probs in res:
        np.testing.assert_array_equal(first_conv, want_first[r0:r1])
        np.testing.assert_allclose(probs, ref["probs"][r0:r1], rtol=1e-4, atol=1e-6)


def test_slice_columns_cover_the_operand():
    for K in (300, 256, 129, 512, 930, 7, 40, 16):
        for world in (2, 3, 4, 8, 16):
            col0, width, ldp = partition.slice_columns(K, world)
            k4 = partition.round_up(K, 4)
            assert int(width.sum()) == k4 and (width % 4 == 0).all() and (ldp >= width).all() and (ldp % 32 == 0).all()
            ends = col0 + width
            assert col0[0] == 0 and (col0[1:][width[1:] > 0] == ends[:-1][width[1:] > 0]).all()
            assert (col0 % 4 == 0).all() and int(width.max()) - int(width[width > 0].min()) <= 16 + 12
    assert partition.slice_columns(300, 8)[1].tolist() == [48, 48, 48, 32, 32, 32, 32, 28]


def test_hot_column_split_reassembles_x():
    """split_hot_columns: hot CSR (hot-local ids) + cold CSR partition X's nonzeros; hot ids ascending; nothing when the
    matrix has no dense columns."""
    import scipy.sparse as sp
    from geographconv_b200.engine import split_hot_columns, panel_col_blocks
    rng = np.random.RandomState(3)
    n, f = 400, 600
    dense_cols = rng.choice(f, size=100, replace=False)
    M = sp.random(n, f, density=0.01, random_state=rng, format="lil", dtype=np.float32)
    for c in dense_cols:
        rows = rng.choice(n, size=n // 3, replace=False)
        M[rows, c] = rng.rand(len(rows)).astype(np.float32) + 0.1
    X = M.tocsr().astype(np.float32)
    X.sort_indices()
    hot_cols, X_hot, X_cold = split_hot_columns(X, 0.2, 96)
    assert len(hot_cols) == 96 and (np.diff(hot_cols) > 0).all() and set(hot_cols) <= set(dense_cols.tolist())
    assert X_hot.shape == (n, 96) and X_cold.shape == X.shape and X_hot.nnz + X_cold.nnz == X.nnz
    back = X_cold.toarray()
    back[:, hot_cols] += X_hot.toarray()
    np.testing.assert_array_equal(back, X.toarray())
    assert not X_cold[:, hot_cols].nnz
    assert split_hot_columns(X, 0.9, 96)[0] is None          # no column that dense
    assert split_hot_columns(X, 0.2, 16)[0] is None          # fewer than 64 hot columns allowed: not worth a GEMM
    # panel working-set rule: 16 MB of 128-byte panel lines per block, at most 8 blocks
    assert panel_col_blocks(62_500) == 1 and panel_col_blocks(500_000) == 4 and panel_col_blocks(10_000_000) == 8


def test_hot_columns_are_chosen_globally():
    """Row blocks given the document frequencies of the WHOLE matrix pick the same hot set as the whole matrix does, so a
    row's sum is associated the same way on 1 and on P GPUs (bit-equal forward)."""
    from geographconv_b200.engine import split_hot_columns
    X = synth.synthetic_features(3000, 400, 40, 3)
    df = np.bincount(X.indices, minlength=X.shape[1])
    hot_all, Xh, Xc = split_hot_columns(X.tocsr(), 0.05, 128)
    assert hot_all is not None
    for r0, r1 in ((0, 1000), (1000, 3000)):
        blk = partition.slice_rows(X, r0, r1)
        hot_b, Xh_b, Xc_b = split_hot_columns(blk, 0.05, 128, df=df, n_total=X.shape[0])
        np.testing.assert_array_equal(hot_b, hot_all)
        np.testing.assert_array_equal(Xh_b.toarray(), Xh[r0:r1].toarray())
        np.testing.assert_array_equal(Xc_b.toarray(), Xc[r0:r1].toarray())


def test_row_grouped_item_plan_covers_every_row_once():
    """X's row-item plan in row groups (engine.HostCsr(row_groups=...)): every group holds exactly the items of its
    128-aligned row range, longest first, and the groups together are the ungrouped plan."""
    from geographconv_b200.engine import HostCsr
    rng = np.random.RandomState(3)
    n, f = 5000, 700
    deg = rng.poisson(12, size=n)
    deg[::17] = 0
    indptr = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(deg, out=indptr[1:])
    M = sp.csr_matrix((rng.randn(indptr[-1]).astype(np.float32), rng.randint(0, f, size=indptr[-1]).astype(np.int32),
                       indptr), shape=(n, f))
    M.sum_duplicates()
    M.sort_indices()
    flat = HostCsr(M, 1024)
    grp = HostCsr(M, 1024, row_groups=8)
    assert flat.row_bounds == [0, n] and flat.item_bounds == [0, n]
    rb, ib = grp.row_bounds, grp.item_bounds
    assert rb[0] == 0 and rb[-1] == n and all(b % 128 == 0 for b in rb[:-1]) and len(rb) == len(ib) and len(rb) > 3
    items = np.asarray(grp.items).reshape(-1, 4)
    seen = np.zeros(n, dtype=np.int64)
    for g in range(len(rb) - 1):
        part = items[ib[g]:ib[g + 1]]
        assert ((part[:, 0] >= rb[g]) & (part[:, 0] < rb[g + 1])).all()
        lens = part[:, 2] - part[:, 1]
        assert (np.diff(lens) <= 0).all()
        np.testing.assert_array_equal(part[:, 1], M.indptr[part[:, 0]])
        np.testing.assert_array_equal(part[:, 2], M.indptr[part[:, 0] + 1])
        seen[part[:, 0]] += 1
    assert (seen == 1).all()
    # a matrix with rows longer than one item keeps the single-group plan (partial sums are per matrix)
    hub = sp.vstack([M, sp.csr_matrix(np.ones((1, f), dtype=np.float32))]).tocsr()
    assert HostCsr(hub, 256, row_groups=8).row_bounds == [0, n + 1]
