"""Device geo-evaluation (csrc/geo.cu, geographconv_b200/geo.py) against the oracle's restatement of
gcnmain.geo_eval (gcnmain.py:43-63).  float64: distances within 1e-12 relative (libm vs CUDA double math differ in
the last ulp), counts exact."""
import numpy as np
import pytest

from oracle import geo_ref

pytestmark = pytest.mark.gpu


def _problem(n, C, seed):
    rng = np.random.RandomState(seed)
    lat = {str(c): float(v) for c, v in enumerate(rng.uniform(-60, 70, size=C))}
    lon = {str(c): float(v) for c, v in enumerate(rng.uniform(-180, 180, size=C))}
    users = ["u%d" % i for i in range(n)]
    loc = {u: "%r,%r" % (float(a), float(b)) for u, a, b in zip(users, rng.uniform(-60, 70, n), rng.uniform(-180, 180, n))}
    y_pred = rng.randint(0, C, size=n)
    return y_pred, users, lat, lon, loc


@pytest.mark.parametrize("n,C", [(1, 1), (1000, 129), (100_000, 256)])
def test_geo_eval_matches_oracle(n, C):
    from geographconv_b200 import geo
    y_pred, users, lat, lon, loc = _problem(n, C, n + C)
    got = geo.geo_eval(y_pred, y_pred, users, lat, lon, loc)
    want = geo_ref.geo_eval(y_pred, y_pred, users, lat, lon, loc)
    np.testing.assert_allclose(got[3], want[3], rtol=1e-12, atol=1e-9)
    assert got[2] == want[2]                      # Acc@161: a count
    np.testing.assert_allclose(got[0], want[0], rtol=1e-12)
    np.testing.assert_allclose(got[1], want[1], rtol=1e-12)
    assert got[4] == want[4] and got[5] == want[5]


def test_geo_known_answers_and_errors():
    from geographconv_b200 import geo
    lat, lon = {"0": 48.8567, "1": 45.7597}, {"0": 2.3508, "1": 4.8422}
    loc = {"lyon": "45.7597,4.8422", "paris": "48.8567,2.3508"}
    mean, median, acc, dist, _, _ = geo.geo_eval([0, 0, 1], [0, 0, 1], ["lyon", "paris", "lyon"], lat, lon, loc)
    assert dist[0] == pytest.approx(392.2172595594006, rel=1e-13) and dist[1] == 0.0 and dist[2] == 0.0
    assert acc == pytest.approx(100 * 2 / 3.0)
    with pytest.raises(AssertionError):
        geo.geo_eval([0], [0, 1], ["lyon"], lat, lon, loc)
    with pytest.raises(KeyError):
        geo.geo_eval([0], [5], ["lyon"], lat, lon, loc)


def test_predict_classes_feeds_geo_eval_on_device():
    """predictions stay on the GPU between the forward pass and the distance kernel"""
    from geographconv_b200 import geo, synth
    from geographconv_b200.gcnmodel import GraphConv
    data = synth.synthetic_dump("tiny")
    A, X_tr, Y_tr, X_dev, Y_dev, X_te, Y_te, U_tr, U_dev, U_te, clat, clon, uloc = data
    import scipy.sparse as sp
    X = sp.vstack([X_tr, X_dev, X_te]).tocsr().astype("float32")
    n_tr, n_dev = X_tr.shape[0], X_dev.shape[0]
    te = np.arange(n_tr + n_dev, X.shape[0]).astype("int32")
    clf = GraphConv(X.shape[1], len(clat), [40, 40, 40], 0.0, 0.5, device=0, shard=False)
    clf.build_model(A, seed=3)
    preds, probs = clf.predict(X, A, te)
    preds2, dev_preds = clf.predict_classes(X, A, te)
    np.testing.assert_array_equal(preds, preds2)
    got = geo.geo_eval(Y_te, preds2, U_te, clat, clon, uloc, preds_device=dev_preds)
    want = geo_ref.geo_eval(Y_te, preds, U_te, clat, clon, uloc)
    np.testing.assert_allclose(got[3], want[3], rtol=1e-12, atol=1e-9)
    assert got[2] == want[2]
