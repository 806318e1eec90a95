"""GPU parity at the EXACT size of BASELINE.json configs[0] (C1: N=9.5k users, avg-degree 8, 10k BoW features with
128 terms per user, 3x300 hidden + highway, 129 classes) -- the configuration the reference runs on the CPU for
"correctness + baseline".  Everything goes through the reference-facing surface (GraphConv -> ctypes -> C ABI) and is
compared with the oracle: probabilities, logits, argmax on every row, one training step (losses, accuracies, the
replayed dropout mask, every gradient in float64, the Adam update)."""
import numpy as np
import pytest

from geographconv_b200 import synth
from oracle import gcn_ref
from parity_util import assert_argmax_parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c1():
    return synth.synthetic_problem("C1")


def _clf(cfg, **kw):
    from geographconv_b200.gcnmodel import GraphConv
    clf = GraphConv(cfg["f"], cfg["classes"], cfg["hid"], regul_coef=kw.get("reg", 0.0), drop_out=0.5, highway=True,
                    shard=False)
    return clf


def test_c1_shapes_are_the_named_ones(c1):
    A, X, Y, tr, dev, te, cfg = c1
    assert cfg == dict(n=9500, deg=8, f=10000, xnnz=128, hid=[300, 300, 300], classes=129)
    assert A.shape == (9500, 9500) and X.shape == (9500, 10000)
    assert 7.0 < A.nnz / 9500 < 9.0 and 110 < X.nnz / 9500 < 140
    assert int(Y.max()) + 1 == 129 and len(tr) + len(dev) + len(te) == 9500


def test_c1_predict_probs_logits_argmax(c1):
    A, X, Y, tr, dev, te, cfg = c1
    clf = _clf(cfg)
    clf.build_model(A, seed=77)
    params = [p.copy() for p in clf.init_params]
    all_rows = np.arange(cfg["n"], dtype=np.int32)
    preds, probs = clf.predict(X, A, all_rows)
    assert preds.dtype == np.int64 and probs.dtype == np.float32 and probs.shape == (cfg["n"], 129)
    ref32 = gcn_ref.forward(params, X, A, cfg["hid"], True)
    ref64 = gcn_ref.forward(params, X, A, cfg["hid"], True, dtype="float64")
    np.testing.assert_allclose(probs, ref32["probs"], rtol=1e-3, atol=1e-7)
    np.testing.assert_allclose(probs, ref64["probs"], rtol=1e-3, atol=1e-7)
    n_mis, worst = assert_argmax_parity(preds, probs, ref64["probs"])
    assert n_mis <= 3, (n_mis, worst)
    eng = clf._get_engine()
    eng.keep_logits = True
    eng.unbind()
    clf.predict(X, A, te)
    logits = eng.read_matrix(eng.logits, eng.n, 129)
    np.testing.assert_allclose(logits, ref64["logits"], rtol=1e-3, atol=1e-3 * float(np.abs(ref64["logits"]).max()))
    # the same through the dev / test index sets the driver uses (gcnmain.py:226,231)
    for idx in (dev, te):
        p, pr = clf.predict(X, A, idx)
        np.testing.assert_array_equal(p, preds[idx])
        np.testing.assert_array_equal(pr, probs[idx])


@pytest.mark.parametrize("reg", [0.0, 1e-5])
def test_c1_train_step_every_gradient(c1, reg):
    A, X, Y, tr, dev, te, cfg = c1
    hid = cfg["hid"]
    clf = _clf(cfg, reg=reg)
    clf.build_model(A, seed=77)
    params = [p.copy() for p in clf.init_params]
    seed = 20181187
    out = clf.f_train(X, Y[tr], Y[dev], A, tr, dev, seed=seed)
    eng = clf._get_engine()
    keep = eng.dropout_mask(seed)
    np.testing.assert_array_equal(keep, gcn_ref.dropout_keep_mask(seed, cfg["n"], hid[0], 0.5))
    scale = keep.astype(np.float32) / np.float32(0.5)
    r64 = gcn_ref.loss_and_grads(params, X, A, Y, tr, hid, True, scale, reg, dtype="float64", dev_idx=dev)
    np.testing.assert_allclose(out[0], r64["train_loss"], rtol=1e-3)
    np.testing.assert_allclose(out[2], r64["dev_loss"], rtol=1e-3)
    np.testing.assert_allclose(out[1], r64["train_acc"], atol=3.0 / len(tr))
    np.testing.assert_allclose(out[3], r64["dev_acc"], atol=3.0 / len(dev))
    names = [e["name"] for e in eng.layout.entries]
    gpu_grads = eng.get_grads()
    for name, g, rg in zip(names, gpu_grads, r64["grads"]):
        np.testing.assert_allclose(g, rg, rtol=1e-3, atol=1e-4 * float(np.abs(rg).max()) + 1e-12, err_msg=name)
    upd = gcn_ref.adam_update(params, gpu_grads, gcn_ref.AdamState(params))
    for name, p, rp in zip(names, eng.get_params(), upd):
        np.testing.assert_allclose(p, rp, rtol=1e-5, atol=2e-6, err_msg=name)
