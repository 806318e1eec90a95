"""CPU oracle for the geographconv GCN hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  The product path
(``geographconv_b200``) never imports it and fails loudly without its CUDA library.

PARITY UNPINNED.  The reference (afshinrahimi/geographconv @ 52c6fd3f) has no tests, no
golden vectors and no fixtures, and its arithmetic lives in
un-vendored third-party code -- Theano 1.0.x and Lasagne master (requirements.txt:5,9) --
that cannot be installed or imported here (Python 3.12 / NumPy 2.3, no network).  This file
therefore restates, op by op, what the reference's call sites ask Theano/Lasagne to
compute, in NumPy/SciPy float32 (float64 on request as an arbiter):

* ``theano.sparse.structured_dot(csr, dense)`` (gcnmodel.py:39,130,153).  Theano's
  ``StructuredDot.perform`` is literally ``a * b`` on SciPy matrices and its C kernel
  ``sd_csr`` is the same row-sequential fp32 axpy loop as SciPy's ``csr_matvecs`` --
  ``structured_dot`` below is that call.  (``oracle/sd_csr.c`` is a plain-C
  restatement of the same loop, checked against SciPy in tests/.)
* ``T.dot`` -> BLAS sgemm == ``numpy.dot`` (gcnmodel.py:126,149).
* Lasagne ``DenseLayer`` / ``DropoutLayer`` / ``get_all_param_values`` / ``adam`` /
  ``GlorotUniform`` / ``Orthogonal`` / ``categorical_crossentropy`` /
  ``regularize_network_params``: published algorithms restated; each function cites the
  reference call site that uses it.

The only vectors that pin it are derived ones: the 4-node normalised-adjacency known
answer, finite-difference gradient checks in float64, and the committed
``tests/golden/*.npz`` fixtures produced by ``tests/golden/make_golden.py`` from this file.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

# --------------------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------------------


def normalize_adjacency(adj, dtype="float32"):
    """A_hat = D^-1/2 (Adj - diag + I) D^-1/2 as CSR (gcnmain.py:115-128).

    ``adj`` is any scipy sparse (or dense) square 0/1 adjacency.  The self loop value is
    1 (gcnmain.py:119-120); rows whose degree is 0 cannot occur after that, but the
    reference still zeroes infinities (gcnmain.py:123-125) and so do we.
    """
    adj = sp.lil_matrix(adj, dtype="float64")
    adj.setdiag(0)
    adj.setdiag(1)
    adj = adj.tocsr()
    m, n = adj.shape
    diags = np.asarray(adj.sum(axis=1)).reshape(-1)
    with np.errstate(divide="ignore"):
        d = 1.0 / np.sqrt(diags)
    d[np.isinf(d)] = 0
    D = sp.spdiags(d, [0], m, n, format="csr")
    A = (D @ adj @ D).tocsr().astype(dtype)
    A.sort_indices()
    return A


def structured_dot(A, B):
    """CSR x dense -> dense; row-sequential fp32 accumulate (gcnmodel.py:39,130,153)."""
    return np.asarray(A @ B)


def _act(name):
    if name == "tanh":  # gcnmodel.py:347 -- the live nonlinearity
        return np.tanh
    if name == "relu":  # gcnmodel.py:345 -- commented-out variant
        return lambda z: np.maximum(z, 0)
    if name == "sigmoid":
        return sigmoid
    if name == "linear":
        return lambda z: z
    raise ValueError(name)


def _act_grad_from_out(name, out):
    """d act / d pre-activation expressed through the activation output."""
    if name == "tanh":
        return 1 - out * out
    if name == "relu":
        return (out > 0).astype(out.dtype)
    if name == "sigmoid":
        return out * (1 - out)
    if name == "linear":
        return np.ones_like(out)
    raise ValueError(name)


def sigmoid(z):
    return 1.0 / (1.0 + np.exp(-z))


def softmax_rows(z):
    """Row-wise softmax as theano.tensor.nnet.softmax computes it (gcnmodel.py:374)."""
    e = np.exp(z - z.max(axis=1, keepdims=True))
    return e / e.sum(axis=1, keepdims=True)


# --------------------------------------------------------------------------------------
# parameters (Lasagne get_all_param_values order)
# --------------------------------------------------------------------------------------


def _glorot_uniform(rng, shape):
    """lasagne.init.GlorotUniform(gain=1): U(-a, a), a = sqrt(6 / (fan_in + fan_out))."""
    a = np.sqrt(6.0 / (shape[0] + shape[1]))
    return rng.uniform(low=-a, high=a, size=shape).astype("float32")


def _orthogonal(rng, shape):
    """lasagne.init.Orthogonal(gain=1): SVD of a standard-normal matrix (gcnmodel.py:359)."""
    a = rng.normal(0.0, 1.0, shape)
    u, _, v = np.linalg.svd(a, full_matrices=False)
    q = u if u.shape == shape else v
    return q.reshape(shape).astype("float32")


def init_params(input_size, hid_size_list, output_size, highway=True, seed=77):
    """Initial weights in the order ``lasagne.layers.get_all_param_values(l_out)`` yields.

    Layer creation order follows gcnmodel.py:351-374: first layer W0 (Glorot), the dropout
    layer (which draws one ``randint`` for its own stream seed), then per hidden layer
    ``l_h`` (Wh Glorot, bh 0) BEFORE ``l_t`` (Wt Orthogonal, bt -4) (gcnmodel.py:281-286),
    then the output layer.  The *value list* order puts the gate before the conv branch
    because MultiplicativeGatingLayer's incomings are [gate, input1, input2]
    (gcnmodel.py:258).  ``np.random.seed(seed)`` is gcnmodel.py:336.
    """
    rng = np.random.RandomState(seed)
    Hd = hid_size_list[0]
    W0 = _glorot_uniform(rng, (input_size, Hd))
    b0 = np.zeros(Hd, "float32")
    rng.randint(1, 2147462579)  # DropoutLayer's RandomStreams seed draw (gcnmodel.py:357)
    params = [W0, b0]
    prev = Hd
    for i, hid in enumerate(hid_size_list):
        if i == 0:
            continue
        if highway:
            Wh = _glorot_uniform(rng, (prev, prev))
            bh = np.zeros(prev, "float32")
            Wt = _orthogonal(rng, (prev, prev))
            bt = np.full(prev, -4.0, "float32")  # gcnmodel.py:274
            params += [Wt, bt, Wh, bh]
        else:
            W = _glorot_uniform(rng, (prev, hid))
            params += [W, np.zeros(hid, "float32")]
            prev = hid
    Wout = _glorot_uniform(rng, (prev, output_size))
    params += [Wout, np.zeros(output_size, "float32")]
    return params


def n_hidden_conv_layers(hid_size_list):
    return max(len(hid_size_list) - 1, 0)


# --------------------------------------------------------------------------------------
# forward / backward
# --------------------------------------------------------------------------------------


def forward(params, X, A, hid_size_list, highway=True, drop_scale=None, nonlin="tanh",
            dtype="float32", keep=False):
    """The network of gcnmodel.py:351-375 on CSR ``X`` (N x F) and CSR ``A`` (N x N).

    ``drop_scale`` is None for the deterministic output (gcnmodel.py:392) or an N x Hd
    array holding mask/(1-p) for the training output (gcnmodel.py:357,375).
    Returns a dict: ``probs``, ``logits`` and, with ``keep=True``, everything backward needs.
    """
    act = _act(nonlin)
    P = [np.asarray(p, dtype=dtype) for p in params]
    X = X.astype(dtype)
    A = A.astype(dtype)
    cache = {"layers": []}
    W0, b0 = P[0], P[1]
    # SparseInputDenseLayer, gcnmodel.py:39-42
    a0 = act(structured_dot(X, W0) + b0[None, :])
    x = a0 if drop_scale is None else a0 * drop_scale.astype(dtype)
    cache["a0"] = a0
    k = 2
    for _ in range(n_hidden_conv_layers(hid_size_list)):
        if highway:
            Wt, bt, Wh, bh = P[k:k + 4]
            k += 4
            # ConvolutionDenseLayer2, gcnmodel.py:126-136: dot, then A, then bias
            h = act(structured_dot(A, x @ Wh) + bh[None, :])
            # gate DenseLayer, gcnmodel.py:285-286 (not convolved)
            t = sigmoid(x @ Wt + bt[None, :])
            y = t * h + (1.0 - t) * x  # gcnmodel.py:266
            cache["layers"].append(("hw", x, h, t))
        else:
            W, b = P[k:k + 2]
            k += 2
            y = act(structured_dot(A, x @ W) + b[None, :])  # gcnmodel.py:372
            cache["layers"].append(("gc", x, y))
        x = y
    Wout, bout = P[k], P[k + 1]
    logits = structured_dot(A, x @ Wout) + bout[None, :]  # gcnmodel.py:149-156
    probs = softmax_rows(logits)
    out = {"probs": probs, "logits": logits, "gates": [l[3] for l in cache["layers"] if l[0] == "hw"]}
    if keep:
        cache["x_last"] = x
        out["cache"] = cache
    return out


def cross_entropy(probs_rows, y):
    """lasagne categorical_crossentropy(...).mean() with integer targets (gcnmodel.py:382)."""
    return float(np.mean(-np.log(probs_rows[np.arange(len(y)), y])))


def loss_and_grads(params, X, A, Y, train_idx, hid_size_list, highway=True, drop_scale=None,
                   regul_coef=0.0, nonlin="tanh", dtype="float32", dev_idx=None):
    """One evaluation of ``f_train``'s outputs and d(train_loss)/d(params).

    Outputs follow gcnmodel.py:375-389,409: train loss/acc and dev loss/acc are all taken
    from the *dropout* output.  The gradient is hand-derived (SURVEY.md 8a "Backward");
    the reference obtains it from ``theano.grad`` inside ``lasagne.updates.adam``
    (gcnmodel.py:407).  Regularisation: ``regul_coef * (L1 + L2)`` over every W, biases
    excluded (gcnmodel.py:383-387).
    """
    P = [np.asarray(p, dtype=dtype) for p in params]
    Xd = X.astype(dtype)
    Ad = A.astype(dtype)
    f = forward(P, Xd, Ad, hid_size_list, highway, drop_scale, nonlin, dtype, keep=True)
    probs, cache = f["probs"], f["cache"]
    train_idx = np.asarray(train_idx, dtype=np.int64)
    y_tr = np.asarray(Y)[train_idx].astype(np.int64)
    n = len(train_idx)
    loss = cross_entropy(probs[train_idx], y_tr)
    acc = float(np.mean(probs[train_idx].argmax(-1) == y_tr))
    res = {"train_loss": loss, "train_acc": acc, "probs": probs, "logits": f["logits"]}
    if dev_idx is not None:
        dev_idx = np.asarray(dev_idx, dtype=np.int64)
        y_dev = np.asarray(Y)[dev_idx].astype(np.int64)
        res["dev_loss"] = cross_entropy(probs[dev_idx], y_dev)
        res["dev_acc"] = float(np.mean(probs[dev_idx].argmax(-1) == y_dev))

    # d loss / d logits: softmax + mean NLL over the gathered rows
    G = np.zeros_like(probs)
    onehot = np.zeros((n, probs.shape[1]), dtype=dtype)
    onehot[np.arange(n), y_tr] = 1
    np.add.at(G, train_idx, (probs[train_idx] - onehot) / n)

    At = Ad.T.tocsr()
    grads = [None] * len(P)
    k = len(P) - 2
    Wout = P[k]
    x = cache["x_last"]
    dq = structured_dot(At, G)
    grads[k] = x.T @ dq
    grads[k + 1] = G.sum(axis=0)
    dx = dq @ Wout.T
    for layer in reversed(cache["layers"]):
        if layer[0] == "hw":
            _, x, h, t = layer
            k -= 4
            Wt, bt, Wh, bh = P[k:k + 4]
            dh_pre = dx * t * _act_grad_from_out(nonlin, h)
            dt_pre = dx * (h - x) * t * (1 - t)
            dx_new = dx * (1 - t)
            du = structured_dot(At, dh_pre)
            grads[k + 2] = x.T @ du
            grads[k + 3] = dh_pre.sum(axis=0)
            dx_new = dx_new + du @ Wh.T
            grads[k] = x.T @ dt_pre
            grads[k + 1] = dt_pre.sum(axis=0)
            dx = dx_new + dt_pre @ Wt.T
        else:
            _, x, y = layer
            k -= 2
            W, b = P[k:k + 2]
            dpre = dx * _act_grad_from_out(nonlin, y)
            du = structured_dot(At, dpre)
            grads[k] = x.T @ du
            grads[k + 1] = dpre.sum(axis=0)
            dx = du @ W.T
    a0 = cache["a0"]
    da0 = dx if drop_scale is None else dx * drop_scale.astype(dtype)
    dz0 = da0 * _act_grad_from_out(nonlin, a0)
    grads[0] = np.asarray(Xd.T.tocsr() @ dz0)
    grads[1] = dz0.sum(axis=0)

    if regul_coef > 0:
        reg = 0.0
        for i, p in enumerate(P):
            if p.ndim == 2:
                reg += np.abs(p).sum() + (p * p).sum()
                grads[i] = grads[i] + regul_coef * (np.sign(p) + 2 * p)
        res["train_loss"] = loss + regul_coef * float(reg)
    res["grads"] = [np.asarray(g, dtype=dtype) for g in grads]
    return res


# --------------------------------------------------------------------------------------
# optimiser and the train/predict entry points
# --------------------------------------------------------------------------------------


class AdamState:
    """State of lasagne.updates.adam (gcnmodel.py:407): shared step count t, m and v per param."""

    def __init__(self, params):
        self.t = 0
        self.m = [np.zeros_like(p) for p in params]
        self.v = [np.zeros_like(p) for p in params]


def adam_update(params, grads, state, lr=2e-3, beta1=0.9, beta2=0.999, eps=1e-8):
    """lasagne.updates.adam restated: a_t = lr*sqrt(1-b2^t)/(1-b1^t); theta -= a_t*m/(sqrt(v)+eps)."""
    dt = params[0].dtype.type
    state.t += 1
    t = dt(state.t)
    one = dt(1)
    a_t = dt(lr) * np.sqrt(one - dt(beta2) ** t) / (one - dt(beta1) ** t)
    out = []
    for i, (p, g) in enumerate(zip(params, grads)):
        m = dt(beta1) * state.m[i] + (one - dt(beta1)) * g
        v = dt(beta2) * state.v[i] + (one - dt(beta2)) * g * g
        state.m[i], state.v[i] = m, v
        out.append((p - a_t * m / (np.sqrt(v) + dt(eps))).astype(p.dtype))
    return out


def train_step(params, state, X, A, Y, train_idx, dev_idx, hid_size_list, highway=True,
               drop_scale=None, regul_coef=0.0, nonlin="tanh", dtype="float32"):
    """``f_train`` (gcnmodel.py:409-410): metrics from the pre-update weights, then Adam."""
    r = loss_and_grads(params, X, A, Y, train_idx, hid_size_list, highway, drop_scale,
                       regul_coef, nonlin, dtype, dev_idx)
    new_params = adam_update([np.asarray(p, dtype=dtype) for p in params], r["grads"], state)
    return new_params, r


def predict(params, X, A, test_idx, hid_size_list, highway=True, nonlin="tanh", dtype="float32"):
    """``f_val`` (gcnmodel.py:411,452-454): deterministic forward, row gather, argmax."""
    f = forward(params, X, A, hid_size_list, highway, None, nonlin, dtype)
    rows = f["probs"][np.asarray(test_idx, dtype=np.int64)]
    return rows.argmax(-1).astype(np.int64), rows.astype("float32")


def get_gates(params, X, A, hid_size_list, highway=True, nonlin="tanh", dtype="float32"):
    """``f_gates`` (gcnmodel.py:396-401,472-477): deterministic gate activations per layer."""
    return forward(params, X, A, hid_size_list, highway, None, nonlin, dtype)["gates"]


# --------------------------------------------------------------------------------------
# Philox4x32-10 -- restated so tests can check the GPU dropout mask bit-for-bit.
# (The reference's MRG31k3p stream is unreproducible without Theano; SURVEY.md 7 item 9.
#  Parity of a training step is defined on an explicit mask, which this regenerates.)
# --------------------------------------------------------------------------------------

_PHILOX_M0 = np.uint64(0xD2511F53)
_PHILOX_M1 = np.uint64(0xCD9E8D57)
_PHILOX_W0 = np.uint32(0x9E3779B9)
_PHILOX_W1 = np.uint32(0xBB67AE85)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32 with 10 rounds; all inputs uint32 arrays/scalars."""
    c0 = np.asarray(c0, np.uint32).copy()
    c1 = np.asarray(c1, np.uint32).copy()
    c2 = np.asarray(c2, np.uint32).copy()
    c3 = np.asarray(c3, np.uint32).copy()
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    mask = np.uint64(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _PHILOX_M0 * c0.astype(np.uint64)
            p1 = _PHILOX_M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & mask).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & mask).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(_PHILOX_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_PHILOX_W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def dropout_keep_mask(seed, n_rows, n_cols, p, row0=0):
    """Keep mask (uint8, n_rows x n_cols) exactly as the CUDA kernels draw it.

    Element (r, c): counter = (global_row, c // 4, 0, 0), key = (seed_lo, seed_hi); the
    four outputs of one Philox call serve columns 4*(c//4) .. +3; keep iff the 32-bit
    draw is < floor((1-p) * 2^32) (all-keep when p == 0).
    """
    if p <= 0:
        return np.ones((n_rows, n_cols), np.uint8)
    thresh = np.uint64(min(int((1.0 - float(p)) * 4294967296.0), 4294967295))
    ngrp = (n_cols + 3) // 4
    rows = (np.arange(n_rows, dtype=np.uint64) + np.uint64(row0)).astype(np.uint32)
    c0 = np.repeat(rows, ngrp)
    c1 = np.tile(np.arange(ngrp, dtype=np.uint32), n_rows)
    z = np.zeros_like(c0)
    r = philox4x32_10(c0, c1, z, z, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    draws = np.stack(r, axis=1).reshape(n_rows, ngrp * 4)[:, :n_cols]
    return (draws.astype(np.uint64) < thresh).astype(np.uint8)
