"""TEST INFRASTRUCTURE: an Engine look-alike that evaluates the oracle on the host.

It exists so that the host logic of ``geographconv_b200.gcnmodel.GraphConv`` (fit loop, early stopping,
save/load/reset, argument handling) and the unchanged reference driver ``gcnmain.main`` can be exercised
on a machine without a GPU.  It is never importable from the product package.
"""
import numpy as np
import scipy.sparse as sp

from oracle import gcn_ref


class OracleEngine:
    def __init__(self, layout, drop_out=0.0, regul_coef=0.0, nonlin="tanh", device=None, group=None):
        self.layout, self.drop_out, self.regul_coef, self.nonlin = layout, float(drop_out), float(regul_coef), nonlin
        self.state = None
        self.calls = []

    def set_params(self, params):
        self.params = [np.asarray(p, dtype=np.float32).copy() for p in params]
        if self.state is None:
            self.state = gcn_ref.AdamState(self.params)

    def get_params(self):
        return [p.copy() for p in self.params]

    def bind(self, X, A, need_backward=True, force_upload=False, assume_symmetric=None):
        if not sp.issparse(X) or not sp.issparse(A):
            raise ValueError("Input for this layer must be sparse")
        self.X, self.A, self.n = X.tocsr(), A.tocsr(), X.shape[0]
        self.calls.append(("bind", X.shape, A.shape))

    def index_arrays(self, idx, labels=None, force_upload=False):
        return (np.asarray(idx), None if labels is None else np.asarray(labels), len(idx))

    def train_step(self, tr, dv, n_train, n_dev, seed, update=True):
        hid, hw = self.layout.hid, self.layout.highway
        keep = gcn_ref.dropout_keep_mask(int(seed) & (2**64 - 1), self.n, hid[0], self.drop_out)
        scale = keep.astype(np.float32) / np.float32(1 - self.drop_out) if self.drop_out > 0 else None
        Y = np.zeros(self.n, dtype=np.int64)
        Y[dv[0]] = dv[1]
        Y[tr[0]] = tr[1]
        new, r = gcn_ref.train_step(self.params, self.state, self.X, self.A, Y, tr[0], dv[0], hid, hw, scale,
                                    self.regul_coef, self.nonlin)
        if update:
            self.params = new
        self.P = r["probs"]
        self._m = (r["train_loss"], r["train_acc"], r["dev_loss"], r["dev_acc"])
        self.calls.append(("train_step", n_train, n_dev))

    def read_metrics(self):
        return self._m

    def forward(self, train=False, seed=0, want_gates=False):
        f = gcn_ref.forward(self.params, self.X, self.A, self.layout.hid, self.layout.highway, None, self.nonlin)
        self.P, self._gates = f["probs"], f["gates"]

    def gather_predictions(self, idx, want_probs=True):
        rows = self.P[np.asarray(idx, dtype=np.int64)]
        preds = rows.argmax(-1).astype(np.int64)
        return (preds, rows.astype(np.float32)) if want_probs else (preds, preds)  # no device here: host array twice

    def read_matrix(self, buf, rows, cols):
        return np.asarray(buf)[:rows, :cols]

    def gates(self):
        return self._gates
