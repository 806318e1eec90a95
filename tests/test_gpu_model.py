"""Model-level parity on the GPU through the reference-facing surface (GraphConv) vs the oracle."""
import numpy as np
import pytest
import scipy.sparse as sp

from geographconv_b200 import synth
from oracle import gcn_ref
from parity_util import assert_argmax_parity

pytestmark = pytest.mark.gpu

SMALL = dict(n=2000, deg=8, f=1500, xnnz=40, hid=[300, 300, 300], classes=129)


def _model(cfg, highway=True, reg=0.0, p=0.5, hid=None):
    from geographconv_b200.gcnmodel import GraphConv
    hid = hid or cfg["hid"]
    clf = GraphConv(cfg["f"], cfg["classes"], hid, regul_coef=reg, drop_out=p, highway=highway, shard=False)
    return clf


@pytest.fixture(scope="module")
def problem():
    return synth.synthetic_problem(SMALL)


@pytest.mark.parametrize("highway,hid", [(True, [300, 300, 300]), (False, [300, 200, 256]), (True, [64]),
                                         (True, [48] * 6)])
def test_predict_matches_oracle(problem, highway, hid):
    A, X, Y, tr, dev, te, cfg = problem
    clf = _model(cfg, highway, hid=hid)
    clf.build_model(A, seed=77)
    params = [p.copy() for p in clf.init_params]
    preds, probs = clf.predict(X, A, te)
    rp, rprob = gcn_ref.predict(params, X, A, te, hid, highway)
    assert preds.dtype == np.int64 and probs.dtype == np.float32 and probs.shape == (len(te), cfg["classes"])
    np.testing.assert_allclose(probs, rprob, rtol=1e-3, atol=1e-7)
    # argmax on ALL rows against the float64 oracle; a mismatch is only tolerated on a numerical tie
    _, rprob64 = gcn_ref.predict(params, X, A, te, hid, highway, dtype="float64")
    n_mis, _ = assert_argmax_parity(preds, probs, rprob64)
    assert n_mis <= 2, n_mis
    # logits (pre-softmax) parity: north_star's "logits within 1e-3"
    eng = clf._get_engine()
    eng.keep_logits = True
    eng._bound_key = None
    clf.predict(X, A, te)
    logits = eng.read_matrix(eng.logits, eng.n, cfg["classes"])
    ref = gcn_ref.forward(params, X, A, hid, highway)["logits"]
    np.testing.assert_allclose(logits, ref, rtol=1e-3, atol=1e-3 * float(np.abs(ref).max()))


@pytest.mark.parametrize("highway,hid,reg", [(True, [300, 300, 300], 0.0), (True, [64, 64], 1e-4),
                                             (False, [96, 64, 80], 0.0)])
def test_train_step_matches_oracle(problem, highway, hid, reg):
    A, X, Y, tr, dev, te, cfg = problem
    clf = _model(cfg, highway, reg=reg, hid=hid)
    clf.build_model(A, seed=5)
    params = [p.copy() for p in clf.init_params]
    seed = 4242
    out = clf.f_train(X, Y[tr], Y[dev], A, tr, dev, seed=seed)
    eng = clf._get_engine()
    keep = eng.dropout_mask(seed)
    np.testing.assert_array_equal(keep, gcn_ref.dropout_keep_mask(seed, cfg["n"], hid[0], 0.5))
    scale = keep.astype(np.float32) / 0.5
    state = gcn_ref.AdamState(params)
    new_params, r = gcn_ref.train_step(params, state, X, A, Y, tr, dev, hid, highway, scale, reg)
    np.testing.assert_allclose(out[0], r["train_loss"], rtol=1e-3)
    np.testing.assert_allclose(out[1], r["train_acc"], atol=2.0 / len(tr))
    np.testing.assert_allclose(out[2], r["dev_loss"], rtol=1e-3)
    np.testing.assert_allclose(out[3], r["dev_acc"], atol=2.0 / len(dev))
    np.testing.assert_allclose(clf.last_output(), r["probs"], rtol=1e-3, atol=1e-7)
    r64 = gcn_ref.loss_and_grads(params, X, A, Y, tr, hid, highway, scale, reg, dtype="float64")
    for name, g, rg in zip([e["name"] for e in eng.layout.entries], eng.get_grads(), r64["grads"]):
        scale_g = float(np.abs(rg).max())
        np.testing.assert_allclose(g, rg, rtol=1e-3, atol=1e-4 * scale_g + 1e-12, err_msg=name)
    # Adam's first step moves every weight by ~lr*sign(g), so a gradient that is pure rounding noise
    # may flip sign between the two implementations: check the update rule itself on the GPU's
    # own gradients (exact comparison), and the end-to-end weights on all but such elements.
    gpu_grads = eng.get_grads()
    upd = gcn_ref.adam_update(params, gpu_grads, gcn_ref.AdamState(params))
    names = [e["name"] for e in eng.layout.entries]
    for name, p, rp, rp2 in zip(names, eng.get_params(), upd, new_params):
        np.testing.assert_allclose(p, rp, rtol=1e-5, atol=2e-6, err_msg=name)
        bad = np.abs(p - rp2) > 4e-4 + 1e-3 * np.abs(rp2)
        assert bad.mean() < 1e-4, (name, float(bad.mean()))


def test_two_steps_adam_state_advances(problem):
    A, X, Y, tr, dev, te, cfg = problem
    hid = [64, 64]
    clf = _model(cfg, True, hid=hid, p=0.0)
    clf.build_model(A, seed=9)
    params = [p.copy() for p in clf.init_params]
    state = gcn_ref.AdamState(params)
    for step in range(3):
        out = clf.f_train(X, Y[tr], Y[dev], A, tr, dev, seed=step)
        params, r = gcn_ref.train_step(params, state, X, A, Y, tr, dev, hid, True, None)
        np.testing.assert_allclose(out[0], r["train_loss"], rtol=2e-3)
    for p, rp in zip(clf.get_all_param_values(), params):
        bad = np.abs(p - rp) > 2e-3 + 1e-2 * np.abs(rp)
        assert bad.mean() < 1e-3


def test_fit_predict_save_load_reset_gates(problem, tmp_path):
    import gzip
    import pickle
    A, X, Y, tr, dev, te, cfg = problem
    # learnable labels: a linear function of the features
    rng = np.random.RandomState(0)
    Wtrue = rng.randn(cfg["f"], 5)
    Yl = np.asarray((A @ (X @ Wtrue))).argmax(1).astype(np.int32)
    from geographconv_b200.gcnmodel import GraphConv
    clf = GraphConv(cfg["f"], 5, [64, 64], regul_coef=0.0, drop_out=0.2, highway=True, shard=False)
    clf.build_model(A, seed=77)
    assert not clf.fitted
    clf.save(lambda obj, fn: None, str(tmp_path / "m.pkl"))  # warns only (gcnmodel.py:463-464)
    clf.fit(X, A, Yl, tr, dev, n_epochs=40, max_down=5, verbose=False)
    assert clf.fitted and len(clf.best_params) == 8
    preds, probs = clf.predict(X, A, te)
    acc = float((preds == Yl[te]).mean())
    assert acc > 0.45, acc  # 5 classes, chance 0.2
    rp, rprob = gcn_ref.predict(clf.best_params, X, A, te, [64, 64], True)
    np.testing.assert_allclose(probs, rprob, rtol=1e-3, atol=1e-6)

    def dump(obj, fn):  # data.py:28-34
        with gzip.open(fn, "wb") as f:
            pickle.dump(obj, f, -1)

    def load(fn):
        with gzip.open(fn, "rb") as f:
            return pickle.load(f)

    fn = str(tmp_path / "model.pkl")
    clf.save(dump, fn)
    clf2 = GraphConv(cfg["f"], 5, [64, 64], regul_coef=0.0, drop_out=0.2, highway=True, shard=False)
    clf2.build_model(A, seed=1)
    clf2.load(load, fn)
    assert clf2.fitted
    p2, pr2 = clf2.predict(X, A, te)
    np.testing.assert_array_equal(p2, preds)
    np.testing.assert_array_equal(pr2, probs)
    gates = clf2.get_gates(X, A)
    ref_g = gcn_ref.get_gates(clf.best_params, X, A, [64, 64], True)
    assert len(gates) == 1 and gates[0].shape == (cfg["n"], 64)
    np.testing.assert_allclose(gates[0], ref_g[0], rtol=1e-3, atol=1e-6)
    clf2.reset()
    for a, b in zip(clf2.get_all_param_values(), clf2.init_params):
        np.testing.assert_array_equal(a, b)
    # a different N through the same weights (feature_report's identity graph, gcnmain.py:244-251)
    V = 50
    Xv = sp.identity(cfg["f"], dtype=np.float32, format="csr")[:V]
    Av = sp.identity(V, dtype=np.float32, format="csr")
    pv, prv = clf.predict(Xv, Av, np.arange(V, dtype=np.int32))
    rpv, rprv = gcn_ref.predict(clf.best_params, Xv, Av, np.arange(V), [64, 64], True)
    np.testing.assert_allclose(prv, rprv, rtol=1e-3, atol=1e-6)


def test_errors_mirror_reference(problem):
    A, X, Y, tr, dev, te, cfg = problem
    clf = _model(cfg, True, hid=[32, 32])
    clf.build_model(A)
    with pytest.raises(ValueError, match="must be sparse"):
        clf.predict(X.toarray(), A, te)  # gcnmodel.py:34-36
    with pytest.raises(ValueError):
        clf.set_all_param_values(clf.init_params[:-1])


def test_nonsymmetric_graph_uses_transpose(problem):
    A, X, Y, tr, dev, te, cfg = problem
    rng = np.random.RandomState(1)
    An = A.copy()
    An.data = (An.data * rng.uniform(0.5, 1.5, size=An.nnz)).astype(np.float32)
    hid = [48, 48]
    clf = _model(cfg, True, hid=hid, p=0.0)
    clf.build_model(An, seed=3)
    params = [p.copy() for p in clf.init_params]
    clf.f_train(X, Y[tr], Y[dev], An, tr, dev, seed=1, update=False)
    eng = clf._get_engine()
    assert not eng.symmetric
    r64 = gcn_ref.loss_and_grads(params, X, An, Y, tr, hid, True, None, 0.0, dtype="float64")
    for g, rg in zip(eng.get_grads(), r64["grads"]):
        np.testing.assert_allclose(g, rg, rtol=1e-3, atol=1e-4 * float(np.abs(rg).max()) + 1e-12)


def test_uncached_inputs_reupload_gives_identical_results(problem):
    """cache_device_inputs=False (bench.py's end-to-end leg): every call copies X, A_hat, X^T again on the copy
    stream, overlapped with the forward pass; results must be bit-identical to the cached run.  The host-side
    pinned staging copies are prepared once per (X, A) identity; a matrix edited in place needs invalidate_inputs()."""
    A, X, Y, tr, dev, te, cfg = problem
    hid = [300, 300]
    outs = []
    for cached in (True, False):
        clf = _model(cfg, True, hid=hid)
        clf.build_model(A, seed=11)
        clf.cache_device_inputs = cached
        o = [clf.f_train(X, Y[tr], Y[dev], A, tr, dev, seed=s, update=True) for s in (1, 2, 3)]
        outs.append((o, clf.get_all_param_values(), clf.predict(X, A, te)[1]))
    for a, b in zip(outs[0][0], outs[1][0]):
        assert a == b
    for a, b in zip(outs[0][1], outs[1][1]):
        np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(outs[0][2], outs[1][2])
    # an in-place edit of the host matrices is picked up after invalidate_inputs()
    clf = _model(cfg, True, hid=hid, p=0.0)
    clf.build_model(A, seed=11)
    clf.cache_device_inputs = False
    X2, A2 = X.copy(), A.copy()
    l0 = clf.f_train(X2, Y[tr], Y[dev], A2, tr, dev, seed=1, update=False)[0]
    X2.data *= np.float32(0.5)
    clf.invalidate_inputs()
    l1 = clf.f_train(X2, Y[tr], Y[dev], A2, tr, dev, seed=1, update=False)[0]
    params = [p.copy() for p in clf.init_params]
    r = gcn_ref.loss_and_grads(params, X2, A2, Y, tr, hid, True, None, 0.0)
    assert l0 != l1
    np.testing.assert_allclose(l1, r["train_loss"], rtol=1e-3)


@pytest.mark.parametrize("engine,xt_blocks", [(0, 1), (1, 2), (2, 4)])
def test_whole_model_parity_with_every_spmm_engine(problem, monkeypatch, engine, xt_blocks):
    """The per-call engine choice picks the L2-panel engine and the column-blocked X^T plan only for operands far larger
    than this test graph, so force each engine (and a blocked plan) through the whole model: predict, one training step
    and every gradient against the oracle."""
    monkeypatch.setenv("GCNB_SPMM_ENGINE", str(engine))
    monkeypatch.setenv("GCNB_XT_BLOCKS", str(xt_blocks))
    A, X, Y, tr, dev, te, cfg = problem
    hid = [300, 300, 300]
    clf = _model(cfg, True, hid=hid)
    clf.build_model(A, seed=21)
    params = [p.copy() for p in clf.init_params]
    preds, probs = clf.predict(X, A, te)
    eng = clf._get_engine()
    assert eng.spmm_engine == engine and eng.A.engine_for(eng, eng.ldh[0], hid[0]) == engine
    rp, rprob = gcn_ref.predict(params, X, A, te, hid, True)
    np.testing.assert_allclose(probs, rprob, rtol=1e-3, atol=1e-7)
    _, rprob64 = gcn_ref.predict(params, X, A, te, hid, True, dtype="float64")
    assert_argmax_parity(preds, probs, rprob64)
    seed = 777
    out = clf.f_train(X, Y[tr], Y[dev], A, tr, dev, seed=seed, update=False)
    assert eng.host.XT.col_blocks == xt_blocks
    keep = gcn_ref.dropout_keep_mask(seed, cfg["n"], hid[0], 0.5)
    r = gcn_ref.loss_and_grads(params, X, A, Y, tr, hid, True, keep.astype(np.float32) / 0.5, 0.0, dtype="float64",
                               dev_idx=dev)
    np.testing.assert_allclose(out[0], r["train_loss"], rtol=1e-3)
    np.testing.assert_allclose(out[2], r["dev_loss"], rtol=1e-3)
    for name, g, rg in zip([e["name"] for e in eng.layout.entries], eng.get_grads(), r["grads"]):
        np.testing.assert_allclose(g, rg, rtol=1e-3, atol=1e-4 * float(np.abs(rg).max()) + 1e-12, err_msg=name)


def test_whole_model_parity_with_the_two_cta_gemm_kernel(problem, monkeypatch):
    """GCNB_GEMM_V=2 (two co-resident tcgen05 CTAs per SM, k-blocks of 16 floats, opt-in): predict, one training step and
    every gradient against the oracle, like the default kernel."""
    monkeypatch.setenv("GCNB_GEMM_V", "2")
    A, X, Y, tr, dev, te, cfg = problem
    hid = [300, 300, 300]
    clf = _model(cfg, True, hid=hid)
    clf.build_model(A, seed=21)
    params = [p.copy() for p in clf.init_params]
    preds, probs = clf.predict(X, A, te)
    eng = clf._get_engine()
    assert eng.ctx.get_option("gemm_v") == 2
    _, rprob64 = gcn_ref.predict(params, X, A, te, hid, True, dtype="float64")
    np.testing.assert_allclose(probs, rprob64, rtol=1e-3, atol=1e-7)
    assert_argmax_parity(preds, probs, rprob64)
    seed = 778
    out = clf.f_train(X, Y[tr], Y[dev], A, tr, dev, seed=seed, update=False)
    keep = gcn_ref.dropout_keep_mask(seed, cfg["n"], hid[0], 0.5)
    r = gcn_ref.loss_and_grads(params, X, A, Y, tr, hid, True, keep.astype(np.float32) / 0.5, 0.0, dtype="float64",
                               dev_idx=dev)
    np.testing.assert_allclose(out[0], r["train_loss"], rtol=1e-3)
    np.testing.assert_allclose(out[2], r["dev_loss"], rtol=1e-3)
    for name, g, rg in zip([e["name"] for e in eng.layout.entries], eng.get_grads(), r["grads"]):
        np.testing.assert_allclose(g, rg, rtol=1e-3, atol=1e-4 * float(np.abs(rg).max()) + 1e-12, err_msg=name)


def test_wide_output_layer_930_classes(problem):
    """TwitterWorld runs at bucket 2400 = 930 classes with 900 hidden units (reference README.md:177-181): the output
    row is wider than one 512-column softmax pass."""
    A, X, Y, tr, dev, te, cfg = problem
    C = 930
    hid = [96, 96]
    from geographconv_b200.gcnmodel import GraphConv
    clf = GraphConv(cfg["f"], C, hid, regul_coef=0.0, drop_out=0.0, highway=True, shard=False)
    clf.build_model(A, seed=4)
    params = [p.copy() for p in clf.init_params]
    preds, probs = clf.predict(X, A, te)
    _, rprob64 = gcn_ref.predict(params, X, A, te, hid, True, dtype="float64")
    np.testing.assert_allclose(probs, rprob64, rtol=1e-3, atol=1e-8)
    np.testing.assert_allclose(probs.sum(1), 1.0, rtol=1e-5)
    assert_argmax_parity(preds, probs, rprob64)
    Yw = (np.arange(cfg["n"]) % C).astype(np.int32)
    out = clf.f_train(X, Yw[tr], Yw[dev], A, tr, dev, seed=1, update=False)
    r = gcn_ref.loss_and_grads(params, X, A, Yw, tr, hid, True, None, 0.0, dtype="float64", dev_idx=dev)
    np.testing.assert_allclose(out[0], r["train_loss"], rtol=1e-3)
    eng = clf._get_engine()
    for name, g, rg in zip([e["name"] for e in eng.layout.entries], eng.get_grads(), r["grads"]):
        np.testing.assert_allclose(g, rg, rtol=1e-3, atol=1e-4 * float(np.abs(rg).max()) + 1e-12, err_msg=name)
    # every SpMM engine takes the wide-softmax route
    from geographconv_b200 import layers
    q = np.random.RandomState(0).randn(cfg["n"], C).astype(np.float32)
    b = np.random.RandomState(1).randn(C).astype(np.float32)
    want = gcn_ref.softmax_rows((A.astype(np.float64) @ q.astype(np.float64)) + b[None, :])
    for variant in (0, 1, 2):
        got = layers.spmm(A, q, bias=b, softmax=True, variant=variant)
        np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-9, err_msg=str(variant))


def test_label_and_input_caches_follow_content(problem):
    """Fresh label vectors at a recycled address (Y[train] temporaries) and in-place edits of X must not be served from
    stale device copies; non-canonical X (duplicate entries, unsorted rows) means what SciPy says it means."""
    A, X, Y, tr, dev, te, cfg = problem
    hid = [64, 64]
    clf = _model(cfg, True, hid=hid, p=0.0)
    clf.build_model(A, seed=2)
    params = [p.copy() for p in clf.init_params]
    Y2 = ((Y.astype(np.int64) * 7 + 3) % cfg["classes"]).astype(np.int32)
    losses = []
    for labels in (Y, Y2):
        y_tr = labels[tr]  # temporary: the second one usually lands on the first one's address
        losses.append(clf.f_train(X, y_tr, labels[dev], A, tr, dev, seed=1, update=False)[0])
        del y_tr
    r1 = gcn_ref.loss_and_grads(params, X, A, Y, tr, hid, True, None, 0.0)["train_loss"]
    r2 = gcn_ref.loss_and_grads(params, X, A, Y2, tr, hid, True, None, 0.0)["train_loss"]
    np.testing.assert_allclose(losses, [r1, r2], rtol=1e-3)
    assert abs(r1 - r2) > 1e-4
    # in-place edit without invalidate_inputs(): the content fingerprint notices
    X2 = X.copy()
    l0 = clf.f_train(X2, Y[tr], Y[dev], A, tr, dev, seed=1, update=False)[0]
    X2.data *= np.float32(0.25)
    l1 = clf.f_train(X2, Y[tr], Y[dev], A, tr, dev, seed=1, update=False)[0]
    r = gcn_ref.loss_and_grads(params, X2, A, Y, tr, hid, True, None, 0.0)["train_loss"]
    assert l0 != l1
    np.testing.assert_allclose(l1, r, rtol=1e-3)
    # duplicates and unsorted rows: X3 stores every entry twice at half weight, rows reversed
    coo = X.tocoo()
    order = np.argsort(-(coo.row.astype(np.int64) * X.shape[1] + coo.col), kind="stable")
    row = np.concatenate([coo.row[order], coo.row[order]])
    col = np.concatenate([coo.col[order], coo.col[order]])
    val = np.concatenate([coo.data[order] * np.float32(0.5)] * 2)
    counts = np.bincount(row, minlength=X.shape[0])
    o2 = np.argsort(row, kind="stable")
    indptr = np.zeros(X.shape[0] + 1, dtype=np.int32)
    np.cumsum(counts, out=indptr[1:])
    X3 = sp.csr_matrix((val[o2], col[o2].astype(np.int32), indptr), shape=X.shape)
    assert not X3.has_canonical_format
    p3, pr3 = clf.predict(X3, A, te)
    p1, pr1 = clf.predict(X, A, te)
    np.testing.assert_allclose(pr3, pr1, rtol=1e-5, atol=1e-9)
