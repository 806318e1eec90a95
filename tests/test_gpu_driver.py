"""The call sequence of the reference driver (gcnmain.main, gcnmain.py:162-232) on the real B200 engine.

/root/reference does not travel to the GPU box, so the sequence is restated here line by line; the UNCHANGED
gcnmain.main itself is run against the same drop-in surface in tests/test_driver_dropin.py (CPU)."""
import gzip
import pickle

import numpy as np
import pytest
import scipy as sp
import scipy.sparse

from geographconv_b200 import synth

pytestmark = pytest.mark.gpu


def test_driver_sequence_on_gpu(tmp_path):
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dropin"))
    sys.modules.pop("gcnmodel", None)
    from gcnmodel import GraphConv  # gcnmain.py:34
    cfg = dict(n=4000, deg=8, f=800, xnnz=30, hid=[64, 64, 64], classes=12)
    data = synth.synthetic_dump(cfg)
    A, X_train, Y_train, X_dev, Y_dev, X_test, Y_test, U_train, U_dev, U_test, clat, clon, uloc = data
    X = sp.sparse.vstack([X_train, X_dev, X_test])                     # gcnmain.py:172
    Y = np.hstack((Y_train, Y_dev, Y_test)).astype("int32")             # :173-177
    X = X.astype("float32"); A = A.astype("float32")                    # :178-179
    input_size, output_size = X.shape[1], np.max(Y) + 1                 # :184-185
    all_train_indices = np.asarray(range(0, X_train.shape[0])).astype("int32")
    clf = GraphConv(input_size=input_size, output_size=output_size, hid_size_list=cfg["hid"], regul_coef=0.0,
                    drop_out=0.5, batchnorm=False, highway=True)       # :191
    clf.build_model(A, use_text=True, use_labels=False, seed=77)        # :192
    np.random.seed(77)
    for percentile in [0.5, 1.0]:                                       # :194 (two label fractions -> reset path)
        selection_size = min(int(percentile * X.shape[0]), all_train_indices.shape[0])
        train_indices = np.random.choice(all_train_indices, size=selection_size, replace=False).astype("int32")  # :206-207
        dev_indices = np.asarray(range(X_train.shape[0], X_train.shape[0] + X_dev.shape[0])).astype("int32")
        test_indices = np.asarray(range(X_train.shape[0] + X_dev.shape[0], X.shape[0])).astype("int32")
        if clf.fitted:
            clf.reset()                                                 # :219-220
        clf.fit(X, A, Y, train_indices=train_indices, val_indices=dev_indices, n_epochs=10000, batch_size=500,
                max_down=3, verbose=False, seed=77)                     # :221
        model_file = str(tmp_path / ("model-%d-%s.pkl" % (A.shape[0], percentile)))

        def dump_obj(obj, filename, protocol=-1):                       # data.py:28-30
            with gzip.open(filename, "wb") as fout:
                pickle.dump(obj, fout, protocol)
        clf.save(dump_obj, model_file)                                  # :223
        y_pred, probs = clf.predict(X, A, dev_indices)                  # :226
        assert len(y_pred) == len(U_dev) and y_pred.dtype == np.int64   # geo_eval's assert, :44
        assert probs.shape == (len(dev_indices), output_size)
        assert all(str(p) in clat for p in y_pred)                      # :53-54 lookups succeed
        y_pred_t, _ = clf.predict(X, A, test_indices)                   # :231
        assert len(y_pred_t) == len(U_test)
        with gzip.open(model_file, "rb") as f:
            assert len(pickle.load(f)) == 12
