"""The layer-level surface of the reference module (gcnmodel.py:29-313; SURVEY.md 8a last row, 8f rank 4): every class and
helper ``GraphConv`` does not reach is a thin composition of the same device ops.  Checked against NumPy/SciPy restatements
of the reference's ``get_output_for`` bodies, 1e-3 relative like the rest of the floating-point parity."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import gcn_ref

pytestmark = pytest.mark.gpu


def _close(got, want, rtol=1e-3):
    want = np.asarray(want)
    np.testing.assert_allclose(got, want, rtol=rtol, atol=2e-5 * max(float(np.abs(want).max()), 1e-30))


@pytest.fixture(scope="module")
def data():
    rng = np.random.RandomState(1)
    n, f, hd = 500, 300, 64
    A = gcn_ref.normalize_adjacency(sp.random(n, n, density=0.02, random_state=rng, format="csr") > 0)
    X = sp.random(n, f, density=0.05, random_state=rng, format="csr", dtype=np.float32)
    x = (0.5 * rng.randn(n, hd)).astype(np.float32)
    ti = rng.choice(n, size=77, replace=False)
    return A, X, x, ti, rng


def test_convolution_variants(data):
    from geographconv_b200 import gcnmodel as m
    A, X, x, ti, rng = data
    for cls, kw in ((m.ConvolutionDenseLayer_zero, {}), (m.ConvolutionDenseLayer, {"target_indices": ti})):
        lay = cls(48, 'tanh', A=A)
        out = lay.get_output_for(x, **kw)
        want = np.tanh(A @ (x @ lay.W) + lay.b[None, :])
        _close(out, want[ti] if kw else want)
    lay = m.ConvolutionLayer(use_target_indices=True, A=A)
    _close(lay.get_output_for(x, target_indices=ti), (A @ x)[ti])
    _close(m.ConvolutionLayer(A=A, nonlinearity='relu').get_output_for(x, target_indices=ti), np.maximum(A @ x, 0))
    lay = m.DenseLayer2(32, 'sigmoid', use_target_indices=True)
    out = lay.get_output_for(x, target_indices=ti)
    _close(out, gcn_ref.sigmoid(x @ lay.W + lay.b[None, :])[ti])
    # ConvolutionDenseLayer2: the gather needs use_target_indices AND target_indices (gcnmodel.py:121-123,134-135)
    lay = m.ConvolutionDenseLayer2(40, 'tanh', use_target_indices=True)
    lay._init(x.shape[1])
    want = np.tanh(A @ (x @ lay.W) + lay.b[None, :])
    _close(lay.get_output_for(x, A=A, target_indices=ti), want[ti])
    _close(lay.get_output_for(x, A=A), want)
    lay2 = m.ConvolutionDenseLayer2(40, 'tanh')
    lay2._init(x.shape[1])
    _close(lay2.get_output_for(x, A=A, target_indices=ti), np.tanh(A @ (x @ lay2.W) + lay2.b[None, :]))
    # ConvolutionDenseLayer3 without a graph: softmax(x.W + b) (gcnmodel.py:152-157)
    lay3 = m.ConvolutionDenseLayer3(9)
    lay3._init(x.shape[1])
    _close(lay3.get_output_for(x), gcn_ref.softmax_rows(x @ lay3.W + lay3.b[None, :]))
    _close(lay3.get_output_for(x, A=A), gcn_ref.softmax_rows(A @ (x @ lay3.W) + lay3.b[None, :]))


def test_sparse_input_variants(data):
    from geographconv_b200 import gcnmodel as m
    A, X, x, ti, rng = data
    for lay, a_call in ((m.SparseConvolutionDenseLayer(40, 'tanh', A=A), None), (m.SparseConvolutionDenseLayer2(40, 'tanh'), A)):
        out = lay.get_output_for(X, A=a_call)
        _close(out, np.tanh(A @ (X @ lay.W) + lay.b[None, :]))
    lay = m.SparseConvolutionDenseLayer2(40, 'relu')
    _close(lay.get_output_for(X), np.maximum(X @ lay.W + lay.b[None, :], 0))      # falsy A: no convolution
    for cls in (m.SparseConvolutionDenseLayer2, m.SparseInputDenseLayer):
        with pytest.raises(ValueError, match="must be sparse"):
            cls(8).get_output_for(x)
    drop = m.SparseInputDropoutLayer(p=0.5)
    assert drop.get_output_for(X, deterministic=True) is X
    with pytest.raises(ValueError, match="must be sparse"):
        drop.get_output_for(x)
    Xd = drop.get_output_for(X, seed=123)
    keep = gcn_ref.dropout_keep_mask(123, 1, X.nnz, 0.5)[0]
    np.testing.assert_array_equal(Xd.indices, X.indices)
    np.testing.assert_array_equal(Xd.data, X.data * np.float32(2.0) * keep)
    assert 0.4 < keep.mean() < 0.6


def test_gating_highway_residual(data):
    from geographconv_b200 import gcnmodel as m
    A, X, x, ti, rng = data
    t = gcn_ref.sigmoid(rng.randn(*x.shape)).astype(np.float32)
    h = np.tanh(rng.randn(*x.shape)).astype(np.float32)
    _close(m.MultiplicativeGatingLayer().get_output_for([t, h, x]), t * h + (1.0 - t) * x, rtol=1e-6)
    with pytest.raises(AssertionError):
        m.MultiplicativeGatingLayer().get_output_for([t, h[:, :5], x])
    hd = x.shape[1]
    Wh, Wt = (0.1 * rng.randn(hd, hd)).astype(np.float32), (0.1 * rng.randn(hd, hd)).astype(np.float32)
    bh, bt = rng.randn(hd).astype(np.float32), np.full(hd, -4.0, np.float32)
    for gconv in (True, False):
        y, tt = m.highway_dense(x, A, gconv=gconv, Wh=Wh, bh=bh, Wt=Wt, bt=bt, nonlinearity='tanh')
        hw = np.tanh((A @ (x @ Wh) if gconv else x @ Wh) + bh[None, :])
        tw = gcn_ref.sigmoid(x @ Wt + bt[None, :])
        _close(tt, tw)
        _close(y, tw * hw + (1 - tw) * x)
    np.random.seed(5)
    y, tt = m.highway_dense(x, A, gconv=True)          # default initialisers: bt = -4 keeps the carry path open
    assert y.shape == x.shape and float(tt.max()) < 0.5
    W = (0.1 * rng.randn(hd, hd)).astype(np.float32)
    z = (A @ (x @ W)) + x
    selu = 1.0507009873554805 * np.where(z > 0, z, 1.6732632423543772 * np.expm1(z))
    _close(m.residual_dense(x, A, W=W), selu)


def test_numpy_helpers():
    from geographconv_b200 import gcnmodel as m
    x = np.array([[1.0, 2.0], [3.0, 4.0]])
    p = m.np_softmax(x)
    assert p.shape == x.shape and abs(p.sum() - 1.0) < 1e-12           # over all entries, as the reference writes it
    a, b = np.arange(10).reshape(10, 1), np.arange(10)
    batches = list(m.iterate_minibatches(a, b, 4))
    assert [len(t) for _, t in batches] == [4, 4] and (batches[1][1] == [4, 5, 6, 7]).all()
    np.random.seed(0)
    got = np.concatenate([t for _, t in m.iterate_minibatches(a, b, 5, shuffle=True)])
    assert sorted(got.tolist()) == list(range(10))
    with pytest.raises(AssertionError):
        list(m.iterate_minibatches(a, b[:5], 2))
