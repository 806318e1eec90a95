"""Kernel-level parity on the GPU: every CUDA entry point of the C ABI against the oracle.

Tolerance (BASELINE.json north_star): 1e-3 relative in fp32 for floating-point results, bit-exact
for integer / index results (argmax, dropout mask)."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import gcn_ref

pytestmark = pytest.mark.gpu

RTOL = 1e-3


def _close(got, want, rtol=RTOL, atol_scale=1e-5):
    want = np.asarray(want)
    atol = atol_scale * max(float(np.abs(want).max()) if want.size else 0.0, 1e-30)
    np.testing.assert_allclose(got, want, rtol=rtol, atol=atol)


def _rand_csr(rng, n, m, avg, hub_rows=(), hub_deg=0, empty_every=0):
    deg = rng.poisson(avg, size=n)
    if empty_every:
        deg[::empty_every] = 0
    deg[1::7] = 1
    for r in hub_rows:
        deg[r] = hub_deg
    deg = np.minimum(deg, m)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(deg, out=rowptr[1:])
    cols = np.concatenate([np.sort(rng.choice(m, size=d, replace=False)) for d in deg]) if rowptr[-1] else np.zeros(0)
    vals = rng.randn(int(rowptr[-1])).astype(np.float32)
    return sp.csr_matrix((vals, cols.astype(np.int32), rowptr.astype(np.int32)), shape=(n, m))


@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("K", [1, 4, 129, 256, 300, 512, 600])
def test_spmm_plain(K, variant):
    from geographconv_b200 import layers
    rng = np.random.RandomState(K)
    A = _rand_csr(rng, 700, 900, 9, empty_every=5)
    B = rng.randn(900, K).astype(np.float32)
    got = layers.spmm(A, B, variant=variant)
    _close(got, gcn_ref.structured_dot(A, B))
    # rows without nonzeros are exactly zero
    assert not got[np.diff(A.indptr) == 0].any()


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_spmm_hub_rows_split_deterministically(variant):
    """A 100k-degree hub row (power-law graphs) is cut into items whose partial
    sums are added in item order: same answer on every run."""
    from geographconv_b200 import layers
    rng = np.random.RandomState(7)
    A = _rand_csr(rng, 300, 120_000, 6, hub_rows=(3, 250), hub_deg=100_000)
    B = rng.randn(120_000, 300).astype(np.float32)
    got = layers.spmm(A, B, variant=variant)
    want = (A.astype(np.float64) @ B.astype(np.float64))
    _close(got, want, atol_scale=2e-6)
    again = layers.spmm(A, B, variant=variant)
    np.testing.assert_array_equal(got, again)
    other_chunk = layers.spmm(A, B, chunk=1024, variant=variant)
    _close(other_chunk, want, atol_scale=2e-6)


def test_spmm_empty_and_degenerate():
    from geographconv_b200 import layers
    rng = np.random.RandomState(0)
    B = rng.randn(10, 8).astype(np.float32)
    assert layers.spmm(sp.csr_matrix((0, 10), dtype=np.float32), B).shape == (0, 8)
    Z = layers.spmm(sp.csr_matrix((5, 10), dtype=np.float32), B, bias=np.arange(8), act="linear")
    np.testing.assert_array_equal(Z, np.tile(np.arange(8, dtype=np.float32), (5, 1)))
    I = sp.identity(10, dtype=np.float32, format="csr")
    np.testing.assert_array_equal(layers.spmm(I, B), B)


@pytest.mark.parametrize("act", ["linear", "tanh", "relu", "sigmoid"])
def test_spmm_bias_activation_epilogue(act):
    from geographconv_b200 import layers
    rng = np.random.RandomState(3)
    A = _rand_csr(rng, 400, 500, 12)
    B = (0.3 * rng.randn(500, 300)).astype(np.float32)
    b = rng.randn(300).astype(np.float32)
    got = layers.spmm(A, B, bias=b, act=act)
    _close(got, gcn_ref._act(act)(gcn_ref.structured_dot(A, B) + b[None, :]), atol_scale=2e-5)


def test_spmm_dropout_epilogue_matches_replayed_mask():
    from geographconv_b200 import layers
    rng = np.random.RandomState(4)
    A = _rand_csr(rng, 300, 200, 10)
    B = (0.3 * rng.randn(200, 300)).astype(np.float32)
    b = rng.randn(300).astype(np.float32)
    seed, row0, p = 0xDEADBEEF12345, 1000, 0.5
    got = layers.spmm(A, B, bias=b, act="tanh", dropout_p=p, seed=seed, row0=row0)
    keep = gcn_ref.dropout_keep_mask(seed, 300, 300, p, row0=row0)
    want = np.tanh(gcn_ref.structured_dot(A, B) + b[None, :]) * keep / (1 - p)
    _close(got, want, atol_scale=2e-5)
    np.testing.assert_array_equal(got == 0, (keep == 0) | (want == 0))
    assert abs(keep.mean() - 0.5) < 0.02


@pytest.mark.parametrize("K", [7, 129, 256])
def test_spmm_softmax_epilogue(K):
    from geographconv_b200 import layers
    rng = np.random.RandomState(K)
    A = _rand_csr(rng, 350, 350, 8)
    B = rng.randn(350, K).astype(np.float32)
    b = rng.randn(K).astype(np.float32)
    P, Z = layers.spmm(A, B, bias=b, softmax=True, want_logits=True)
    zw = gcn_ref.structured_dot(A, B) + b[None, :]
    _close(Z, zw)
    _close(P, gcn_ref.softmax_rows(zw), atol_scale=1e-6)
    np.testing.assert_allclose(P.sum(1), 1.0, rtol=1e-5)
    clear = np.sort(zw, axis=1)
    clear = (clear[:, -1] - clear[:, -2]) > 1e-4
    np.testing.assert_array_equal(P.argmax(1)[clear], zw.argmax(1)[clear])


def test_spmm_accumulate():
    from geographconv_b200 import layers
    rng = np.random.RandomState(5)
    A = _rand_csr(rng, 200, 300, 10)
    B = rng.randn(300, 300).astype(np.float32)
    C0 = rng.randn(200, 300).astype(np.float32)
    got = layers.spmm(A, B, accumulate_into=C0)
    _close(got, C0 + gcn_ref.structured_dot(A, B))


@pytest.mark.parametrize("panel,unroll", [(16, 0), (16, 16), (32, 0), (32, 16), (64, 0)])
def test_spmm_panel_engine_epilogues(panel, unroll):
    """Engine 2 (L2-resident column panels): every epilogue the step uses, at each panel width; a row's sum is
    accumulated in CSR order like the other engines, so the engines agree to the last bit on plain products."""
    from geographconv_b200 import layers
    rng = np.random.RandomState(panel + unroll)
    A = _rand_csr(rng, 777, 640, 11, hub_rows=(5,), hub_deg=600, empty_every=9)
    K = 300
    B = (0.3 * rng.randn(640, K)).astype(np.float32)
    b = rng.randn(K).astype(np.float32)
    kw = dict(variant=2, panel=panel, unroll=unroll, chunk=256)
    plain = layers.spmm(A, B, **kw)
    np.testing.assert_array_equal(plain, layers.spmm(A, B, variant=0, chunk=256))
    _close(plain, gcn_ref.structured_dot(A, B))
    seed, row0, p = 0xABCDEF, 12345, 0.5
    got = layers.spmm(A, B, bias=b, act="tanh", dropout_p=p, seed=seed, row0=row0, **kw)
    keep = gcn_ref.dropout_keep_mask(seed, 777, K, p, row0=row0)
    _close(got, np.tanh(gcn_ref.structured_dot(A, B) + b[None, :]) * keep / (1 - p), atol_scale=2e-5)
    C0 = rng.randn(777, K).astype(np.float32)
    _close(layers.spmm(A, B, accumulate_into=C0, **kw), C0 + gcn_ref.structured_dot(A, B))
    got = layers.spmm(A, B, bias=b, act="tanh", accumulate_into=C0, accumulate_mode=2, **kw)
    _close(got, np.tanh(C0 + gcn_ref.structured_dot(A, B) + b[None, :]), atol_scale=2e-5)
    for Kc in (7, 129, 256):
        Bc = rng.randn(640, Kc).astype(np.float32)
        bc = rng.randn(Kc).astype(np.float32)
        P, Z = layers.spmm(A, Bc, bias=bc, softmax=True, want_logits=True, **kw)
        zw = gcn_ref.structured_dot(A, Bc) + bc[None, :]
        _close(Z, zw)
        _close(P, gcn_ref.softmax_rows(zw), atol_scale=1e-6)
        P0 = layers.spmm(A, Bc, bias=bc, softmax=True, variant=0, chunk=256)
        np.testing.assert_array_equal(P, P0)


@pytest.mark.parametrize("tA,tB", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(257, 300, 300), (1000, 129, 300), (300, 300, 5000), (3, 5, 2), (129, 300, 20000)])
def test_gemm_simt(tA, tB, M, N, K):
    from geographconv_b200 import layers
    rng = np.random.RandomState(M + N + K)
    A = rng.randn(K, M).astype(np.float32) if tA else rng.randn(M, K).astype(np.float32)
    B = rng.randn(N, K).astype(np.float32) if tB else rng.randn(K, N).astype(np.float32)
    want = (A.T if tA else A).astype(np.float64) @ (B.T if tB else B).astype(np.float64)
    got = layers.gemm(A, B, transA=bool(tA), transB=bool(tB), tc=0)
    _close(got, want, atol_scale=3e-6)


@pytest.fixture(params=[1, 2], ids=["one_cta_per_sm", "two_ctas_per_sm"])
def gemm_v(request):
    """Both tcgen05 GEMM kernels (ctx option gemm_v): the default one-tile-per-SM kernel with k-blocks of 32 floats and
    the opt-in kernel with two co-resident CTAs per SM and k-blocks of 16."""
    from geographconv_b200 import layers
    d = layers.get_dev()
    old = d.ctx.get_option("gemm_v")
    d.ctx.set_option("gemm_v", request.param)
    yield request.param
    d.ctx.set_option("gemm_v", old)


@pytest.mark.parametrize("tB", [0, 1])
@pytest.mark.parametrize("M,N,K", [(257, 300, 300), (1000, 129, 300), (1000, 300, 256), (128, 32, 8), (3000, 512, 300),
                                   (77, 5, 3), (130, 160, 1000), (1000, 304, 17), (513, 170, 48), (4000, 256, 1024)])
def test_gemm_tcgen05_3xtf32(tB, M, N, K, gemm_v):
    """tcgen05 path (3xTF32 split): fp32-grade accuracy against a float64 product."""
    from geographconv_b200 import layers
    rng = np.random.RandomState(M + N + K + tB)
    A = rng.randn(M, K).astype(np.float32)
    B = rng.randn(N, K).astype(np.float32) if tB else rng.randn(K, N).astype(np.float32)
    want = A.astype(np.float64) @ (B.T if tB else B).astype(np.float64)
    d = layers.get_dev()
    before = d.ctx.get_option("tc_launches")
    got = layers.gemm(A, B, transB=bool(tB), tc=1)
    assert d.ctx.get_option("tc_launches") == before + 1, "the tcgen05 kernel did not run"
    _close(got, want, atol_scale=3e-6)


@pytest.mark.parametrize("M,N,K", [(300, 300, 5000), (129, 300, 20000), (300, 256, 777), (40, 24, 100), (512, 512, 4096),
                                   (300, 300, 31), (7, 3, 5), (1536, 300, 3000), (2048, 129, 700)])
def test_wgrad_tcgen05_split_k(M, N, K, gemm_v):
    """dW = x^T . V on tensor cores: MN-major operands, split over the node dimension, deterministic."""
    from geographconv_b200 import layers
    rng = np.random.RandomState(M + N + K)
    A = rng.randn(K, M).astype(np.float32)
    B = rng.randn(K, N).astype(np.float32)
    want = A.T.astype(np.float64) @ B.astype(np.float64)
    d = layers.get_dev()
    before = d.ctx.get_option("tc_launches")
    got = layers.gemm(A, B, transA=True, tc=1)
    assert d.ctx.get_option("tc_launches") == before + 1, "the tcgen05 wgrad kernel did not run"
    _close(got, want, atol_scale=3e-6)
    np.testing.assert_array_equal(got, layers.gemm(A, B, transA=True, tc=1))
    C0 = rng.randn(M, N).astype(np.float32)
    _close(layers.gemm(A, B, transA=True, accumulate_into=C0, tc=1), C0 + want, atol_scale=3e-6)


def test_gemm_tcgen05_bias_act_accumulate(gemm_v):
    from geographconv_b200 import layers
    rng = np.random.RandomState(12)
    A = (0.1 * rng.randn(700, 300)).astype(np.float32)
    B = rng.randn(300, 300).astype(np.float32)
    b = rng.randn(300).astype(np.float32)
    _close(layers.gemm(A, B, bias=b, act="tanh", tc=1), np.tanh(A.astype(np.float64) @ B + b), atol_scale=1e-5)
    C0 = rng.randn(700, 300).astype(np.float32)
    _close(layers.gemm(A, B, transB=True, accumulate_into=C0, tc=1), C0 + A.astype(np.float64) @ B.T.astype(np.float64),
           atol_scale=1e-5)


def test_gemm_bias_act_and_accumulate():
    from geographconv_b200 import layers
    rng = np.random.RandomState(11)
    A = (0.1 * rng.randn(500, 300)).astype(np.float32)
    B = rng.randn(300, 300).astype(np.float32)
    b = rng.randn(300).astype(np.float32)
    _close(layers.gemm(A, B, bias=b, act="sigmoid", tc=0), gcn_ref.sigmoid(A @ B + b), atol_scale=1e-5)
    C0 = rng.randn(500, 300).astype(np.float32)
    _close(layers.gemm(A, B, accumulate_into=C0, tc=0), C0 + A @ B, atol_scale=1e-5)


@pytest.mark.parametrize("tc", [0, 1])
@pytest.mark.parametrize("n,hd", [(1000, 300), (777, 40), (129, 512), (5000, 129)])
def test_highway_forward(n, hd, tc, gemm_v):
    from geographconv_b200 import layers
    rng = np.random.RandomState(n)
    S = rng.randn(n, hd).astype(np.float32)
    X = rng.randn(n, hd).astype(np.float32)
    Wh = (rng.randn(hd, hd) / np.sqrt(hd)).astype(np.float32)
    Wt = (rng.randn(hd, hd) / np.sqrt(hd)).astype(np.float32)
    bh = rng.randn(hd).astype(np.float32)
    bt = (rng.randn(hd) - 1).astype(np.float32)
    d = layers.get_dev()
    before = d.ctx.get_option("tc_launches")
    Y, H, T = layers.highway(S, X, Wh, bh, Wt, bt, tc=tc)
    assert d.ctx.get_option("tc_launches") == before + tc
    h = np.tanh(S.astype(np.float64) @ Wh + bh)
    t = gcn_ref.sigmoid(X.astype(np.float64) @ Wt + bt)
    _close(H, h, atol_scale=2e-5)
    _close(T, t, atol_scale=2e-5)
    _close(Y, t * h + (1 - t) * X, atol_scale=2e-5)


@pytest.mark.parametrize("n,k", [(1000, 300), (777, 129), (5000, 512), (64, 7), (300, 600)])
def test_fused_backward_with_bias_gradients(n, k):
    """gcnb_highway_bwd_bias_f32 / gcnb_act_bwd_bias_f32: element-wise outputs identical to the unfused kernels,
    bias gradients = column sums (float64 reference), run-to-run identical."""
    import ctypes as C
    import torch
    from geographconv_b200.layers import get_dev
    from geographconv_b200.partition import ld_of
    d = get_dev()
    rng = np.random.RandomState(n + k)
    ld = ld_of(k)
    host = {name: rng.randn(n, k).astype(np.float32) for name in ("dY", "X")}
    host["H"] = np.tanh(rng.randn(n, k)).astype(np.float32)
    host["T"] = (1 / (1 + np.exp(-rng.randn(n, k)))).astype(np.float32)
    dev = {name: d.dense(v)[0] for name, v in host.items()}
    outs = {name: torch.zeros(n * ld, dtype=torch.float32, device=d.dev) for name in
            ("dH", "dT", "dX", "dH2", "dT2", "dX2", "dZ", "dZ2")}
    db = {name: torch.zeros(ld, dtype=torch.float32, device=d.dev) for name in ("bh", "bt", "bz", "bh_again")}
    d.ensure_ws(2 * d.ctx.lib.gcnb_colsum_workspace_bytes(n, k))
    p = lambda t: C.c_void_p(t.data_ptr())
    d.fence()
    d.ctx.call("gcnb_highway_bwd_f32", n, k, ld, p(dev["dY"]), p(dev["X"]), p(dev["H"]), p(dev["T"]), 1, p(outs["dH"]),
               p(outs["dT"]), p(outs["dX"]))
    d.ctx.call("gcnb_highway_bwd_bias_f32", n, k, ld, p(dev["dY"]), p(dev["X"]), p(dev["H"]), p(dev["T"]), 1,
               p(outs["dH2"]), p(outs["dT2"]), p(outs["dX2"]), p(db["bh"]), p(db["bt"]))
    d.ctx.call("gcnb_highway_bwd_bias_f32", n, k, ld, p(dev["dY"]), p(dev["X"]), p(dev["H"]), p(dev["T"]), 1,
               p(outs["dH2"]), p(outs["dT2"]), p(outs["dX2"]), p(db["bh_again"]), p(db["bt"]))
    seed, row0, pdrop = 99, 7, 0.5
    d.ctx.call("gcnb_act_bwd_f32", n, k, ld, p(dev["dY"]), p(dev["H"]), 1, pdrop, seed, row0, p(outs["dZ"]))
    d.ctx.call("gcnb_act_bwd_bias_f32", n, k, ld, p(dev["dY"]), p(dev["H"]), 1, pdrop, seed, row0, p(outs["dZ2"]), p(db["bz"]))
    got = {name: d.download(t, n, ld, k) for name, t in outs.items()}
    gb = {name: d.download(t, 1, ld, k)[0] for name, t in db.items()}
    for a in ("dH", "dT", "dX", "dZ"):
        np.testing.assert_array_equal(got[a], got[a + "2"])
    g, x, h, t = (host[q].astype(np.float64) for q in ("dY", "X", "H", "T"))
    _close(got["dH"], g * t * (1 - h * h))
    _close(got["dT"], g * (h - x) * t * (1 - t))
    _close(got["dX"], g * (1 - t))
    for name, ref in (("bh", got["dH"]), ("bt", got["dT"]), ("bz", got["dZ"])):
        want = ref.astype(np.float64).sum(0)
        np.testing.assert_allclose(gb[name], want, rtol=1e-4, atol=1e-4 * np.abs(ref).sum(0).max())
    np.testing.assert_array_equal(gb["bh"], gb["bh_again"])


@pytest.mark.parametrize("M,N,K,acc", [(1000, 300, 300, 1), (257, 300, 300, 0), (4000, 129, 512, 1)])
def test_gemm_pair_one_pass(M, N, K, acc, gemm_v):
    """gcnb_gemm_pair_f32: C (+)= A1.B1^T + A2.B2^T with both k-loops in one tcgen05 kernel (fp32-grade, 3xTF32)."""
    import ctypes as C
    import torch
    from geographconv_b200.layers import get_dev
    d = get_dev()
    rng = np.random.RandomState(M + N)
    A1, A2 = rng.randn(M, K).astype(np.float32), rng.randn(M, K).astype(np.float32)
    B1, B2 = rng.randn(N, K).astype(np.float32), rng.randn(N, K).astype(np.float32)  # used transposed (dgrad: W is (in, out))
    C0 = rng.randn(M, N).astype(np.float32)
    dA1, lda = d.dense(A1)
    dA2, _ = d.dense(A2)
    dB1, ldb = d.dense(B1)
    dB2, _ = d.dense(B2)
    dC, ldc = d.dense(C0)
    d.ensure_ws(2 * d.ctx.lib.gcnb_gemm_workspace_bytes(0, M, N, K))
    p = lambda t: C.c_void_p(t.data_ptr())
    d.fence()
    before = d.ctx.get_option("tc_launches")
    d.ctx.call("gcnb_gemm_pair_f32", 1, M, N, K, p(dA1), lda, p(dB1), ldb, p(dA2), lda, p(dB2), ldb, p(dC), ldc, acc)
    got = d.download(dC, M, ldc, N)
    assert d.ctx.get_option("tc_launches") == before + 1
    want = A1.astype(np.float64) @ B1.astype(np.float64).T + A2.astype(np.float64) @ B2.astype(np.float64).T
    if acc:
        want = want + C0
    _close(got, want, atol_scale=3e-6)
