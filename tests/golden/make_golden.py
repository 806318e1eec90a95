"""Generate tests/golden/*.npz from the oracle (oracle/gcn_ref.py).

PARITY UNPINNED: the reference ships no golden vectors and its runtime (Theano/Lasagne) cannot be
imported here, so these fixtures freeze the oracle's outputs (checked in tests/test_oracle.py
against SciPy, finite differences, closed forms and Philox known answers) rather than outputs of
the reference itself.  They guard against drift of either side: tests/test_golden.py checks the
oracle still reproduces them (CPU) and that the CUDA path matches them (GPU).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from geographconv_b200 import synth  # noqa: E402
from oracle import gcn_ref  # noqa: E402

CASES = {
    "hw3": dict(cfg=dict(n=96, deg=6, f=64, xnnz=10, hid=[20, 20, 20], classes=7), highway=True, reg=0.0, p=0.5),
    "plain": dict(cfg=dict(n=80, deg=5, f=48, xnnz=8, hid=[24, 12, 16], classes=5), highway=False, reg=1e-3, p=0.25),
}


def build(name):
    c = CASES[name]
    cfg = c["cfg"]
    A, X, Y, tr, dev, te, _ = synth.synthetic_problem(cfg, seed=11)
    params = gcn_ref.init_params(cfg["f"], cfg["hid"], cfg["classes"], c["highway"], 11)
    seed = 0x5EED0000 + len(name)
    keep = gcn_ref.dropout_keep_mask(seed, cfg["n"], cfg["hid"][0], c["p"])
    scale = keep.astype(np.float32) / np.float32(1 - c["p"])
    det = gcn_ref.forward(params, X, A, cfg["hid"], c["highway"])
    state = gcn_ref.AdamState(params)
    new_params, r = gcn_ref.train_step(params, state, X, A, Y, tr, dev, cfg["hid"], c["highway"], scale, c["reg"])
    out = dict(A_indptr=A.indptr, A_indices=A.indices, A_data=A.data, X_indptr=X.indptr, X_indices=X.indices,
               X_data=X.data, Y=Y, tr=tr, dev=dev, te=te, seed=np.uint64(seed), keep=keep,
               det_probs=det["probs"], det_logits=det["logits"], train_probs=r["probs"],
               metrics=np.array([r["train_loss"], r["train_acc"], r["dev_loss"], r["dev_acc"]], dtype=np.float64))
    for i, (p, g, q) in enumerate(zip(params, r["grads"], new_params)):
        out["param_%d" % i], out["grad_%d" % i], out["new_param_%d" % i] = p, g, q
    for i, g in enumerate(det["gates"]):
        out["gate_%d" % i] = g
    return out


if __name__ == "__main__":
    for name in CASES:
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **build(name))
        print("wrote", name)
