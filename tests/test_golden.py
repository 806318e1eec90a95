"""Committed fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py): the oracle must
still reproduce them (CPU) and the CUDA path must match them (GPU)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import gcn_ref

HERE = os.path.dirname(os.path.abspath(__file__))
import sys
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402


def _load(name):
    z = np.load(os.path.join(HERE, "golden", name + ".npz"))
    c = make_golden.CASES[name]
    cfg = c["cfg"]
    A = sp.csr_matrix((z["A_data"], z["A_indices"], z["A_indptr"]), shape=(cfg["n"], cfg["n"]))
    X = sp.csr_matrix((z["X_data"], z["X_indices"], z["X_indptr"]), shape=(cfg["n"], cfg["f"]))
    n_par = len([k for k in z.files if k.startswith("param_")])
    params = [z["param_%d" % i] for i in range(n_par)]
    return z, c, cfg, A, X, params


@pytest.mark.parametrize("name", sorted(make_golden.CASES))
def test_oracle_reproduces_golden(name):
    z = np.load(os.path.join(HERE, "golden", name + ".npz"))
    fresh = make_golden.build(name)
    assert sorted(fresh) == sorted(z.files)
    for k in z.files:
        if np.asarray(fresh[k]).dtype.kind in "iub":
            np.testing.assert_array_equal(fresh[k], z[k], err_msg=k)
        else:  # BLAS summation order may differ between hosts
            np.testing.assert_allclose(fresh[k], z[k], rtol=2e-4, atol=1e-6, err_msg=k)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(make_golden.CASES))
def test_cuda_path_matches_golden(name):
    from geographconv_b200.gcnmodel import GraphConv
    z, c, cfg, A, X, params = _load(name)
    clf = GraphConv(cfg["f"], cfg["classes"], cfg["hid"], regul_coef=c["reg"], drop_out=c["p"],
                    highway=c["highway"], shard=False)
    clf.build_model(A, seed=11)
    for a, b in zip(clf.init_params, params):
        np.testing.assert_array_equal(a, b)  # same init stream as the fixture
    Y, tr, dev, te = z["Y"], z["tr"], z["dev"], z["te"]
    preds, probs = clf.predict(X, A, te)
    np.testing.assert_allclose(probs, z["det_probs"][te], rtol=1e-3, atol=1e-7)
    np.testing.assert_array_equal(preds, z["det_probs"][te].argmax(1))
    for i, g in enumerate(clf.get_gates(X, A)):
        np.testing.assert_allclose(g, z["gate_%d" % i], rtol=1e-3, atol=1e-7)
    seed = int(z["seed"])
    out = clf.f_train(X, Y[tr], Y[dev], A, tr, dev, seed=seed)
    eng = clf._get_engine()
    np.testing.assert_array_equal(eng.dropout_mask(seed), z["keep"])
    np.testing.assert_allclose(out, z["metrics"], rtol=1e-3, atol=1e-6)
    np.testing.assert_allclose(clf.last_output(), z["train_probs"], rtol=1e-3, atol=1e-7)
    for i, g in enumerate(eng.get_grads()):
        want = z["grad_%d" % i]
        np.testing.assert_allclose(g, want, rtol=1e-3, atol=1e-4 * float(np.abs(want).max()) + 1e-12)
