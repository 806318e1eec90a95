"""Host logic of the drop-in surface on CPU (the oracle stands in for the device engine), including a run of the
UNCHANGED reference driver gcnmain.main on top of dropin/gcnmodel.py when /root/reference is present."""
import gzip
import importlib
import os
import pickle
import sys

import numpy as np
import pytest

from geographconv_b200 import synth
from oracle import gcn_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get("GEOGRAPHCONV_REFERENCE", "/root/reference")
SMALL = dict(n=300, deg=6, f=120, xnnz=12, hid=[16, 16, 16], classes=6)


@pytest.fixture()
def oracle_engine(monkeypatch):
    from geographconv_b200 import gcnmodel
    from fake_engine import OracleEngine

    def _get_engine(self):
        if self._engine is None:
            self._engine = OracleEngine(self.layout, self.drop_out, self.regul_coef, self.nonlinearity)
            self._engine.set_params(self._host_params)
        return self._engine

    monkeypatch.setattr(gcnmodel.GraphConv, "_get_engine", _get_engine)
    return OracleEngine


def test_fit_early_stopping_and_best_weights(oracle_engine):
    """fit = full-batch loop, early stop on dev LOSS after max_down bad epochs and n > 2*max_down, best weights
    restored (gcnmodel.py:418-450)."""
    from geographconv_b200.gcnmodel import GraphConv
    A, X, Y, tr, dev, te, cfg = synth.synthetic_problem(SMALL)
    clf = GraphConv(cfg["f"], cfg["classes"], cfg["hid"], 0.0, 0.0, highway=True)
    clf.build_model(A, seed=7)
    clf.fit(X, A, Y, tr, dev, n_epochs=60, max_down=3, verbose=False)
    assert clf.fitted
    # replay with the oracle directly
    params = gcn_ref.init_params(cfg["f"], cfg["hid"], cfg["classes"], True, 7)
    state = gcn_ref.AdamState(params)
    best, best_loss, down, steps = None, float("inf"), 0, 0
    for n in range(60):
        before = params
        params, r = gcn_ref.train_step(params, state, X, A, Y, tr, dev, cfg["hid"], True, None)
        steps += 1
        if r["dev_loss"] < best_loss:
            # the reference snapshots get_all_param_values AFTER f_train applied its update (gcnmodel.py:430-438)
            best_loss, best, down = r["dev_loss"], params, 0
        else:
            down += 1
        if down > 3 and n > 6:
            break
    eng = clf._get_engine()
    assert sum(1 for c in eng.calls if c[0] == "train_step") == steps
    for a, b in zip(clf.best_params, best):
        np.testing.assert_allclose(a, b, rtol=1e-6, atol=1e-7)
    for a, b in zip(clf.get_all_param_values(), best):
        np.testing.assert_allclose(a, b, rtol=1e-6, atol=1e-7)


def test_surface_signatures_match_reference_calls(oracle_engine):
    """The exact keyword calls gcnmain.main makes (gcnmain.py:191-192,221,226) bind to our signatures."""
    import inspect
    from geographconv_b200.gcnmodel import GraphConv
    inspect.signature(GraphConv.__init__).bind(None, input_size=10, output_size=np.int64(3), hid_size_list=[4],
                                                regul_coef=0.0, drop_out=0.5, batchnorm=False, highway=True)
    inspect.signature(GraphConv.build_model).bind(None, "A", use_text=True, use_labels=False, seed=77)
    inspect.signature(GraphConv.fit).bind(None, "X", "A", "Y", train_indices=1, val_indices=2, n_epochs=10000,
                                           batch_size=500, max_down=10, verbose=True, seed=77)
    inspect.signature(GraphConv.predict).bind(None, "X", "A", "idx")
    params = list(inspect.signature(GraphConv.__init__).parameters)[1:9]
    assert params == ["input_size", "output_size", "hid_size_list", "regul_coef", "drop_out", "dtype", "batchnorm",
                      "highway"]  # gcnmodel.py:321
    fit_params = list(inspect.signature(GraphConv.fit).parameters)[1:]
    assert fit_params == ["X", "H", "Y", "train_indices", "val_indices", "n_epochs", "batch_size", "max_down",
                          "pseudolikelihood_thresh", "verbose", "seed"]  # gcnmodel.py:418


def test_unfitted_save_warns_and_load_roundtrip(oracle_engine, tmp_path, caplog):
    from geographconv_b200.gcnmodel import GraphConv
    A, X, Y, tr, dev, te, cfg = synth.synthetic_problem(SMALL)
    clf = GraphConv(cfg["f"], cfg["classes"], cfg["hid"], 0.0, 0.5, highway=False)
    clf.build_model(A)
    called = []
    clf.save(lambda obj, fn: called.append(fn), "x.pkl")
    assert called == []  # gcnmodel.py:463-464: only warns
    clf.fit(X, A, Y, tr, dev, n_epochs=3, verbose=False)
    fn = str(tmp_path / "m.pkl")

    def dump(obj, filename):
        with gzip.open(filename, "wb") as f:
            pickle.dump(obj, f, -1)

    def load(filename):
        with gzip.open(filename, "rb") as f:
            return pickle.load(f)

    clf.save(dump, fn)
    other = GraphConv(cfg["f"], cfg["classes"], cfg["hid"], 0.0, 0.5, highway=False)
    other.build_model(A, seed=1)
    other.load(load, fn)
    p1, pr1 = clf.predict(X, A, te)
    p2, pr2 = other.predict(X, A, te)
    np.testing.assert_array_equal(p1, p2)
    np.testing.assert_array_equal(pr1, pr2)


@pytest.mark.skipif(not os.path.exists(os.path.join(REFERENCE, "gcnmain.py")), reason="reference checkout not present")
def test_unchanged_reference_driver_runs_on_the_dropin(oracle_engine, tmp_path, monkeypatch, caplog):
    """gcnmain.main (imported unchanged from the reference) drives dropin/gcnmodel.GraphConv end to end."""
    for p in (REFERENCE, os.path.join(ROOT, "tests", "shims"), os.path.join(ROOT, "dropin")):
        monkeypatch.syspath_prepend(p)
    for m in ("gcnmodel", "gcnmain", "data", "haversine", "matplotlib", "matplotlib.collections", "kdtree"):
        monkeypatch.delitem(sys.modules, m, raising=False)
    gcnmain = importlib.import_module("gcnmain")
    import gcnmodel as dropin_module
    assert dropin_module.__file__.startswith(os.path.join(ROOT, "dropin"))
    assert gcnmain.GraphConv is dropin_module.GraphConv
    args = gcnmain.parse_args(["-hid", "16", "16", "16", "-highway", "-dropout", "0.5", "-reg", "0.0", "-maxdown", "2",
                               "-silent", "-save"])
    gcnmain.model_args = args  # the script sets this global in __main__ (gcnmain.py:305)
    data = synth.synthetic_dump(SMALL)
    monkeypatch.chdir(tmp_path)
    os.makedirs("data")
    import logging
    with caplog.at_level(logging.INFO):
        gcnmain.main(data, args, batch=args.batch, hidden=args.hid, regularization=args.regularization,
                     dropout=args.dropout, percent=args.percent)
    text = caplog.text
    assert "dev results:" in text and "test results:" in text and "Acc@161" in text  # gcnmain.py:61,225,230
    assert os.path.exists("gcn_1.0_percent_pred_%d.pkl" % SMALL["classes"])          # gcnmain.py:228
    model_file = "./data/model-%d-1.0.pkl" % SMALL["n"]                              # gcnmain.py:196
    assert os.path.exists(model_file)
    with gzip.open(model_file, "rb") as f:
        weights = pickle.load(f)
    assert [w.shape for w in weights] == [(120, 16), (16,), (16, 16), (16,), (16, 16), (16,), (16, 16), (16,),
                                          (16, 16), (16,), (16, 6), (6,)]


def test_predict_classes_equals_predict_without_probabilities(oracle_engine):
    """GraphConv.predict_classes: same argmax as predict (gcnmodel.py:452-454), no probability rows returned."""
    from geographconv_b200.gcnmodel import GraphConv
    A, X, Y, tr, dev, te, cfg = synth.synthetic_problem(SMALL)
    clf = GraphConv(cfg["f"], cfg["classes"], cfg["hid"], 0.0, 0.5)
    clf.build_model(A, seed=1)
    preds, probs = clf.predict(X, A, te)
    p2, handle = clf.predict_classes(X, A, te)
    np.testing.assert_array_equal(preds, p2)
    assert p2.dtype == np.int64 and len(handle) == len(te)
    with pytest.raises(ValueError, match="must be sparse"):
        clf.predict_classes(X.toarray(), A, te)


def test_dump_and_weight_pickles_round_trip(tmp_path):
    """geographconv_b200.io: dump.pkl in the reference's gzip+pickle format (data.py:28-34) and the input assembly of
    gcnmain.main (gcnmain.py:153-212); a Python-2 style pickle (latin1 bytes) loads through the fallback."""
    import gzip
    from geographconv_b200 import io as gio
    data = synth.synthetic_dump(SMALL)
    fn = str(tmp_path / "dump.pkl")
    gio.dump_obj(data, fn)
    with gzip.open(fn, "rb") as f:                      # what the reference's load_obj does
        back = pickle.load(f)
    assert len(back) == 13 and (back[0] != data[0]).nnz == 0 and (back[2] == data[2]).all()
    got = gio.assemble(gio.load_obj(fn))
    A, X_tr, Y_tr, X_dev, Y_dev, X_te, Y_te = data[:7]
    n_tr, n_dev, n_te = X_tr.shape[0], X_dev.shape[0], X_te.shape[0]
    assert got["X"].format == "csr" and got["X"].dtype == np.float32 and got["A"].dtype == np.float32
    assert got["X"].shape == (n_tr + n_dev + n_te, X_tr.shape[1]) and got["Y"].dtype == np.int32
    assert (got["X"][n_tr:n_tr + n_dev] != X_dev.astype(np.float32)).nnz == 0
    np.testing.assert_array_equal(got["Y"], np.concatenate([Y_tr, Y_dev, Y_te]))
    np.testing.assert_array_equal(got["dev_indices"], np.arange(n_tr, n_tr + n_dev))
    np.testing.assert_array_equal(got["test_indices"], np.arange(n_tr + n_dev, n_tr + n_dev + n_te))
    assert got["train_indices"].dtype == np.int32 and got["output_size"] == int(max(Y_tr.max(), Y_dev.max(), Y_te.max())) + 1
    with pytest.raises(ValueError):
        gio.assemble(data[:12])
    # a pickle holding non-ASCII byte strings the way Python 2 wrote them: protocol 2, str opcode with raw bytes
    py2 = b"\x80\x02U\x04caf\xe9q\x00."                 # pickle.dumps('caf\xe9') under Python 2
    fn2 = str(tmp_path / "py2.pkl")
    with gzip.open(fn2, "wb") as f:
        f.write(py2)
    assert gio.load_obj(fn2) == "caf\xe9"
