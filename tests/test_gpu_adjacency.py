"""A_hat construction on the GPU (csrc/adjacency.cu) against the oracle's restatement of gcnmain.py:115-128.

Index arrays must be bit-exact; values too (the kernel does the reference's float64 arithmetic and rounds once)."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import gcn_ref

pytestmark = pytest.mark.gpu


def _oracle_from_edges(u, v, n):
    u = np.asarray(u, dtype=np.int64)
    v = np.asarray(v, dtype=np.int64)
    adj = sp.coo_matrix((np.ones(2 * len(u)), (np.concatenate([u, v]), np.concatenate([v, u]))), shape=(n, n)).tocsr()
    adj.data[:] = 1  # nx.Graph: a repeated edge is one edge
    return gcn_ref.normalize_adjacency(adj)


def _same(A, R):
    assert A.shape == R.shape and A.nnz == R.nnz
    assert A.indices.dtype == np.int32 and A.indptr.dtype == np.int32 and A.data.dtype == np.float32
    np.testing.assert_array_equal(A.indptr, R.indptr)
    np.testing.assert_array_equal(A.indices, R.indices)
    np.testing.assert_array_equal(A.data, R.data.astype(np.float32))


def test_known_answer_path_graph():
    """path 0-1-2 plus isolated node 3."""
    from geographconv_b200 import adjacency
    A = adjacency.normalized_adjacency_from_edges([0, 1], [1, 2], 4)
    want = np.array([[.5, .4082483, 0, 0], [.4082483, .3333333, .4082483, 0], [0, .4082483, .5, 0], [0, 0, 0, 1]],
                    dtype=np.float32)
    np.testing.assert_allclose(A.toarray(), want, rtol=0, atol=1e-7)
    assert A.nnz == 8
    _same(A, _oracle_from_edges([0, 1], [1, 2], 4))


@pytest.mark.parametrize("n,m,seed", [(1, 0, 0), (7, 0, 1), (50, 400, 2), (1000, 3000, 3), (5000, 80_000, 4)])
def test_random_graphs_match_oracle_bit_exact(n, m, seed):
    """duplicates, reversed duplicates, self loops, isolated nodes; rows of 1 .. a few hundred entries"""
    from geographconv_b200 import adjacency
    rng = np.random.RandomState(seed)
    u = rng.randint(0, n, size=m)
    v = rng.randint(0, n, size=m)
    if m:
        u = np.concatenate([u, v[: m // 10], u[: m // 20]])  # reversed and repeated edges
        v = np.concatenate([v, u[: m // 10], v[: m // 20]])
    A = adjacency.normalized_adjacency_from_edges(u, v, n)
    _same(A, _oracle_from_edges(u, v, n))


def test_hub_rows_take_the_cta_and_global_sort_paths():
    """a row of 3000 edge ends (CTA sort in shared memory) and one of 40000 (in-place sort in global memory), each
    with repeated edges"""
    from geographconv_b200 import adjacency
    rng = np.random.RandomState(11)
    n = 30_000
    hub1 = np.full(3000, 5)
    hub2 = np.full(40_000, 17)
    u = np.concatenate([hub1, hub2, rng.randint(0, n, size=50_000)])
    v = np.concatenate([rng.randint(0, 2000, size=3000), rng.randint(0, 20_000, size=40_000), rng.randint(0, n, size=50_000)])
    A = adjacency.normalized_adjacency_from_edges(u, v, n)
    R = _oracle_from_edges(u, v, n)
    assert np.diff(R.indptr)[5] > 128 and np.diff(R.indptr)[17] > 4096
    _same(A, R)
    again = adjacency.normalized_adjacency_from_edges(u, v, n)  # the atomic scatter order must not show
    _same(again, A)


def test_matches_synthetic_generator_and_properties_at_scale():
    """N = 200k, avg degree 32: equals the NumPy generator; symmetric; unit-degree-scaled diagonal; sorted columns"""
    from geographconv_b200 import adjacency, synth
    n, deg = 200_000, 32
    rng = np.random.RandomState(77)
    m = n * (deg - 1) // 2
    u = rng.randint(0, n, size=m, dtype=np.int64)
    v = rng.randint(0, n, size=m, dtype=np.int64)
    A = adjacency.normalized_adjacency_from_edges(u, v, n)
    _same(A, synth.normalized_adjacency_from_edges(u, v, n))
    d = np.diff(A.indptr)
    np.testing.assert_allclose(A.diagonal(), 1.0 / d, rtol=1e-6)
    assert (abs(A - A.T) > 0).nnz == 0
    assert A.has_sorted_indices and all(np.all(np.diff(A.indices[A.indptr[i]:A.indptr[i + 1]]) > 0) for i in range(0, n, 997))


def test_normalize_adjacency_drop_in_and_errors():
    from geographconv_b200 import adjacency, capi
    rng = np.random.RandomState(5)
    n = 300
    M = sp.random(n, n, density=0.02, random_state=rng, format="csr")
    M.data[:] = 1
    adj = ((M + M.T) > 0).astype(np.int64)
    adj.setdiag(1)  # existing self loops are replaced by the unit one either way
    A = adjacency.normalize_adjacency(adj)
    _same(A, gcn_ref.normalize_adjacency(adj))
    with pytest.raises(ValueError):
        adjacency.normalize_adjacency(sp.triu(adj, k=1))  # asymmetric
    # weighted graph (nx.adjacency_matrix(..., weight='w'), gcnmain.py:115): integer weights sum exactly in float64,
    # so the result is bit-identical to the oracle; real weights agree to the last float32 bit or one ulp (row-sum order)
    T = sp.triu(adj, k=1).tocoo()
    for weights, exact in ((rng.randint(1, 6, size=T.nnz).astype(np.float64), True), (rng.rand(T.nnz) + 0.25, False)):
        U = sp.csr_matrix((weights, (T.row, T.col)), shape=(n, n))
        W = (U + U.T).tocsr()
        W.setdiag(7.0)  # replaced by the unit self loop
        Aw = adjacency.normalize_adjacency(W)
        ref = gcn_ref.normalize_adjacency(W)
        if exact:
            _same(Aw, ref)
        else:
            ref.sort_indices()
            np.testing.assert_array_equal(Aw.indptr, ref.indptr)
            np.testing.assert_array_equal(Aw.indices, ref.indices)
            np.testing.assert_allclose(Aw.data, ref.data, rtol=1.2e-7, atol=0)
        assert Aw.dtype == np.float32 and Aw.indices.dtype == np.int32
    Wa = W.copy().tolil()
    Wa[0, 1] = Wa[0, 1] + 1.0 if Wa[0, 1] else 3.0
    with pytest.raises(ValueError):
        adjacency.normalize_adjacency(Wa.tocsr())  # asymmetric weights
    with pytest.raises(ValueError):
        adjacency.normalized_adjacency_from_edges([0, 5], [1, 2], 4)
    # the C ABI itself refuses an out-of-range id (device-side check)
    import ctypes as C
    import torch
    from geographconv_b200.layers import get_dev
    d = get_dev()
    u = d.upload(np.array([0, 9], dtype=np.int32))
    v = d.upload(np.array([1, 2], dtype=np.int32))
    work = torch.empty(int(d.ctx.lib.gcnb_adj_workspace_bytes(2, 4)), dtype=torch.uint8, device=d.dev)
    rowptr = torch.empty(5, dtype=torch.int32, device=d.dev)
    nnz = C.c_int64()
    d.fence()
    with pytest.raises(capi.GcnbError):
        d.ctx.call("gcnb_adj_build_rows", C.c_void_p(u.data_ptr()), C.c_void_p(v.data_ptr()), 2, 4,
                   C.c_void_p(work.data_ptr()), work.numel(), C.c_void_p(rowptr.data_ptr()), C.byref(nnz))
