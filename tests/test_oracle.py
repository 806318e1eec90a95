"""The oracle against everything that can pin it without the reference's runtime:
the derived 4-node known answer, SciPy's csr_matvecs, float64 finite differences, closed forms."""
import ctypes

import pytest

import numpy as np
import scipy.sparse as sp

from oracle import gcn_ref
from geographconv_b200 import synth


def test_normalized_adjacency_known_answer():
    # path graph 0-1-2 plus isolated node 3, gcnmain.py:115-128
    adj = sp.lil_matrix((4, 4))
    adj[0, 1] = adj[1, 0] = adj[1, 2] = adj[2, 1] = 1
    A = gcn_ref.normalize_adjacency(adj)
    want = np.array([[.5, .4082483, 0, 0], [.4082483, .3333333, .4082483, 0], [0, .4082483, .5, 0], [0, 0, 0, 1]],
                    dtype=np.float32)
    assert A.nnz == 8 and A.dtype == np.float32
    np.testing.assert_allclose(A.toarray(), want, atol=1e-7)
    assert abs(A - A.T).max() == 0


def test_synthetic_graph_matches_reference_normalisation():
    rng = np.random.RandomState(0)
    n = 200
    u, v = rng.randint(0, n, 600), rng.randint(0, n, 600)
    A = synth.normalized_adjacency_from_edges(u, v, n)
    adj = sp.coo_matrix((np.ones(len(u)), (u, v)), shape=(n, n)).tolil()
    adj = ((adj + adj.T) > 0).astype(np.float64)
    want = gcn_ref.normalize_adjacency(adj)
    assert A.has_sorted_indices and A.indices.dtype == np.int32 and A.dtype == np.float32
    np.testing.assert_array_equal(A.indptr, want.indptr)
    np.testing.assert_array_equal(A.indices, want.indices)
    np.testing.assert_allclose(A.data, want.data, rtol=1e-6)


def test_sd_csr_c_restatement_is_bit_exact_with_scipy(sdcsr):
    rng = np.random.RandomState(1)
    for n, m, k, dens in [(37, 53, 300, 0.2), (64, 64, 129, 0.05), (5, 9, 1, 0.9), (10, 10, 7, 0.0)]:
        A = sp.random(n, m, density=dens, format="csr", dtype=np.float32, random_state=rng)
        A.sort_indices()
        B = rng.randn(m, k).astype(np.float32)
        Z = np.full((n, k), np.nan, dtype=np.float32)
        f32p, i32p = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int32)
        sdcsr.sd_csr_f32(ctypes.c_int32(n), A.indptr.ctypes.data_as(i32p), A.indices.ctypes.data_as(i32p),
                         A.data.ctypes.data_as(f32p), B.ctypes.data_as(f32p), ctypes.c_int64(k),
                         Z.ctypes.data_as(f32p), ctypes.c_int64(k), ctypes.c_int32(k))
        np.testing.assert_array_equal(Z, gcn_ref.structured_dot(A, B))
        # gradient form: structured_dot(a.T, g)
        G = rng.randn(n, k).astype(np.float32)
        Zt = np.full((m, k), np.nan, dtype=np.float32)
        sdcsr.sd_csc_f32(ctypes.c_int32(n), ctypes.c_int32(m), A.indptr.ctypes.data_as(i32p),
                         A.indices.ctypes.data_as(i32p), A.data.ctypes.data_as(f32p), G.ctypes.data_as(f32p),
                         ctypes.c_int64(k), Zt.ctypes.data_as(f32p), ctypes.c_int64(k), ctypes.c_int32(k))
        np.testing.assert_allclose(Zt, (A.T.tocsr() @ G), rtol=1e-5, atol=1e-6)


def _tiny(highway=True, hid=(6, 6, 6), classes=4, n=30, f=20, seed=3):
    rng = np.random.RandomState(seed)
    A = synth.synthetic_graph(n, 4, seed)
    X = sp.random(n, f, density=0.3, format="csr", dtype=np.float32, random_state=rng)
    Y = rng.randint(0, classes, n).astype(np.int32)
    params = gcn_ref.init_params(f, list(hid), classes, highway, seed)
    params = [p + 0.1 * rng.randn(*p.shape).astype(np.float32) for p in params]
    return A, X, Y, params, list(hid)


def _loss64(params, A, X, Y, tr, hid, highway, scale, reg):
    f = gcn_ref.forward(params, X, A, hid, highway, scale, dtype="float64")
    loss = gcn_ref.cross_entropy(f["probs"][tr], Y[tr])
    if reg > 0:
        loss += reg * sum(np.abs(p).sum() + (p * p).sum() for p in params if np.ndim(p) == 2)
    return loss


def _fd_check(highway, hid, reg):
    A, X, Y, params, hid = _tiny(highway, hid)
    rng = np.random.RandomState(5)
    tr = np.arange(0, 18)
    scale = (rng.rand(X.shape[0], hid[0]) > 0.5).astype(np.float64) / 0.5
    P64 = [p.astype(np.float64) for p in params]
    r = gcn_ref.loss_and_grads(P64, X, A, Y, tr, hid, highway, scale, reg, dtype="float64")
    assert abs(r["train_loss"] - _loss64(P64, A, X, Y, tr, hid, highway, scale, reg)) < 1e-12
    eps = 1e-6
    for pi, p in enumerate(P64):
        flat = p.reshape(-1)
        for j in rng.choice(flat.size, size=min(6, flat.size), replace=False):
            old = flat[j]
            flat[j] = old + eps
            lp = _loss64(P64, A, X, Y, tr, hid, highway, scale, reg)
            flat[j] = old - eps
            lm = _loss64(P64, A, X, Y, tr, hid, highway, scale, reg)
            flat[j] = old
            fd = (lp - lm) / (2 * eps)
            g = r["grads"][pi].reshape(-1)[j]
            assert abs(fd - g) <= 1e-6 + 1e-5 * abs(fd), (pi, j, fd, g)


def test_backward_matches_finite_differences_highway():
    _fd_check(True, (6, 6, 6), 0.0)


def test_backward_matches_finite_differences_plain_and_regularised():
    _fd_check(False, (6, 5, 7), 1e-3)


def test_parameter_order_and_init():
    # get_all_param_values order: W0,b0,(Wt,bt,Wh,bh)*,Wout,bout (gcnmodel.py:258,274)
    P = gcn_ref.init_params(50, [8, 8, 8], 5, True, 77)
    assert [p.shape for p in P] == [(50, 8), (8,), (8, 8), (8,), (8, 8), (8,), (8, 8), (8,), (8, 8), (8,), (8, 5), (5,)]
    assert np.all(P[3] == -4.0) and np.all(P[5] == 0) and np.all(P[1] == 0)
    np.testing.assert_allclose(P[2] @ P[2].T, np.eye(8), atol=1e-5)  # Wt orthogonal
    assert np.abs(P[0]).max() <= np.sqrt(6.0 / 58) + 1e-7
    # the product host must draw the same stream
    from geographconv_b200.gcnmodel import initial_parameters
    Q = initial_parameters(50, [8, 8, 8], 5, True, 77)
    for a, b in zip(P, Q):
        np.testing.assert_array_equal(a, b)
    Pn = gcn_ref.init_params(50, [8, 6], 5, False, 77)
    Qn = initial_parameters(50, [8, 6], 5, False, 77)
    assert [p.shape for p in Pn] == [(50, 8), (8,), (8, 6), (6,), (6, 5), (5,)]
    for a, b in zip(Pn, Qn):
        np.testing.assert_array_equal(a, b)


def test_adam_first_step_closed_form():
    # lasagne.updates.adam, t=1: m=(1-b1)g, v=(1-b2)g^2, a=lr*sqrt(1-b2)/(1-b1) => step = lr*g/(|g|+eps')
    p = [np.array([1.0, -2.0, 3.0], dtype=np.float64)]
    g = [np.array([0.5, -0.25, 0.0], dtype=np.float64)]
    st = gcn_ref.AdamState(p)
    out = gcn_ref.adam_update(p, g, st)
    a = 2e-3 * np.sqrt(1 - 0.999) / (1 - 0.9)
    want = p[0] - a * (0.1 * g[0]) / (np.sqrt(0.001 * g[0] ** 2) + 1e-8)
    np.testing.assert_allclose(out[0], want, rtol=1e-12)
    assert st.t == 1


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = gcn_ref.philox4x32_10(*[np.array([c], np.uint32) for c in ctr], key[0], key[1])
        assert tuple(int(g[0]) for g in got) == want


def test_dropout_mask_rate_and_determinism():
    m = gcn_ref.dropout_keep_mask(99, 400, 300, 0.5)
    assert m.shape == (400, 300) and abs(m.mean() - 0.5) < 0.01
    assert (m == gcn_ref.dropout_keep_mask(99, 400, 300, 0.5)).all()
    assert (m[100:200] == gcn_ref.dropout_keep_mask(99, 100, 300, 0.5, row0=100)).all()
    assert gcn_ref.dropout_keep_mask(99, 4, 6, 0.0).all()


def test_forward_shapes_and_softmax():
    A, X, Y, params, hid = _tiny()
    f = gcn_ref.forward(params, X, A, hid, True)
    np.testing.assert_allclose(f["probs"].sum(1), 1.0, rtol=1e-5)
    assert len(f["gates"]) == 2 and f["gates"][0].shape == (30, 6)
    pr, pb = gcn_ref.predict(params, X, A, np.array([1, 5, 7]), hid, True)
    assert pr.dtype == np.int64 and pb.dtype == np.float32 and pb.shape == (3, 4)


def test_geo_oracle_known_answer_and_statistics():
    """haversine package's documented value (Lyon-Paris) pins the distance; geo_eval statistics follow gcnmain.py:57-63."""
    from oracle import geo_ref
    assert geo_ref.haversine((45.7597, 4.8422), (48.8567, 2.3508)) == pytest.approx(392.2172595594006, rel=1e-15)
    assert geo_ref.haversine((10.0, 20.0), (10.0, 20.0)) == 0.0
    lat = {"0": 40.0, "1": 30.0}
    lon = {"0": -100.0, "1": -90.0}
    loc = {"a": "40.0,-100.0", "b": "41.0,-100.0", "c": "30.0,-80.0"}
    mean, median, acc, dist, t, p = geo_ref.geo_eval([0, 0, 1], [0, 0, 1], ["a", "b", "c"], lat, lon, loc)
    assert dist[0] == 0.0 and dist[1] == pytest.approx(111.195, rel=1e-4) and dist[2] > 161
    assert acc == pytest.approx(100 * 2 / 3.0) and median == dist[1] and mean == pytest.approx(sum(dist) / 3)
    assert t == [[40.0, -100.0], [41.0, -100.0], [30.0, -80.0]] and p == [[40.0, -100.0], [40.0, -100.0], [30.0, -90.0]]
