"""On-disk formats either side of the hot path: the reference's ``dump.pkl`` and model pickles.

* ``dump_obj`` / ``load_obj`` -- gzip + pickle exactly like data.py:28-34 (these are the callables gcnmain.py hands to
  ``GraphConv.save`` / ``load``, gcnmain.py:216,223).  ``load_obj`` also reads dumps written by the reference's original
  Python 2 runs (README.md:26-46 publishes such files): NumPy arrays inside a Python 2 pickle need ``encoding='latin1'``.
* ``assemble`` -- the input assembly at the top of ``gcnmain.main`` (gcnmain.py:153-212): the 13-tuple
  ``(A, X_train, Y_train, X_dev, Y_dev, X_test, Y_test, U_train, U_dev, U_test, classLatMedian, classLonMedian,
  userLocation)`` becomes the stacked float32 CSR ``X``, float32 CSR ``A``, int32 ``Y`` and the int32 index ranges that
  ``GraphConv.fit`` / ``predict`` take.  Host-side only: the device copies are made by ``Engine.bind`` on first use.
"""
from __future__ import annotations

import gzip
import pickle

import numpy as np
import scipy.sparse as sp

DUMP_FIELDS = ("A", "X_train", "Y_train", "X_dev", "Y_dev", "X_test", "Y_test", "U_train", "U_dev", "U_test",
               "classLatMedian", "classLonMedian", "userLocation")


def dump_obj(obj, filename, protocol=-1, serializer=pickle):
    """data.py:28-30."""
    with gzip.open(filename, 'wb') as fout:
        serializer.dump(obj, fout, protocol)


def load_obj(filename, serializer=pickle):
    """data.py:31-34; falls back to ``encoding='latin1'`` for pickles written under Python 2."""
    with gzip.open(filename, 'rb') as fin:
        try:
            return serializer.load(fin)
        except UnicodeDecodeError:
            fin.seek(0)
            return serializer.load(fin, encoding='latin1')


def assemble(data, dtype='float32', dtypeint='int32'):
    """gcnmain.py:153-212 without the model: returns a dict with X, A, Y, train_indices (all training rows, the
    ``-lblfraction 1.0`` case), dev_indices, test_indices, input_size, output_size and the evaluation tables."""
    if len(data) != len(DUMP_FIELDS):
        raise ValueError("dump.pkl holds %d fields, expected the 13-tuple of gcnmain.py:153" % len(data))
    d = dict(zip(DUMP_FIELDS, data))
    X = sp.vstack([d["X_train"], d["X_dev"], d["X_test"]])                       # gcnmain.py:164
    Y_train, Y_dev, Y_test = (np.asarray(d[k]) for k in ("Y_train", "Y_dev", "Y_test"))
    Y = np.hstack((Y_train, Y_dev, Y_test)) if Y_train.ndim == 1 else np.vstack((Y_train, Y_dev, Y_test))
    Y = Y.astype(dtypeint)                                                       # gcnmain.py:169
    X = X.astype(dtype).tocsr()                                                  # vstack of CSR blocks is CSR
    A = sp.csr_matrix(d["A"]).astype(dtype)
    n_tr, n_dev, n_te = d["X_train"].shape[0], d["X_dev"].shape[0], d["X_test"].shape[0]
    if A.shape != (n_tr + n_dev + n_te,) * 2:
        raise ValueError("A is %s but the feature blocks hold %d rows" % (A.shape, n_tr + n_dev + n_te))
    return {
        "X": X, "A": A, "Y": Y,
        "input_size": X.shape[1], "output_size": int(np.max(Y)) + 1,                       # gcnmain.py:177-178
        "train_indices": np.asarray(range(0, n_tr)).astype(dtypeint),                      # gcnmain.py:182
        "dev_indices": np.asarray(range(n_tr, n_tr + n_dev)).astype(dtypeint),             # gcnmain.py:211
        "test_indices": np.asarray(range(n_tr + n_dev, n_tr + n_dev + n_te)).astype(dtypeint),  # gcnmain.py:212
        "Y_dev": Y_dev, "Y_test": Y_test, "U_dev": d["U_dev"], "U_test": d["U_test"], "U_train": d["U_train"],
        "classLatMedian": d["classLatMedian"], "classLonMedian": d["classLonMedian"], "userLocation": d["userLocation"],
    }
