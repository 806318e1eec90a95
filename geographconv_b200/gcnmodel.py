"""B200 drop-in for the reference ``gcnmodel.GraphConv`` (reference gcnmodel.py:316-477).

Same constructor, ``build_model`` / ``fit`` / ``predict`` / ``reset`` / ``save`` / ``load`` /
``get_gates`` signatures, argument meaning, logging lines and error behaviour, so the reference
driver ``gcnmain.py`` runs on it unchanged (``from gcnmodel import GraphConv``, gcnmain.py:34 --
put ``dropin/`` on ``sys.path``).  The arithmetic runs in hand-written sm_100a CUDA kernels behind
the C ABI of ``libgcnb200.so`` (``include/gcnb200.h``); there is no CPU fallback: constructing the
device engine without the library or without a B200 raises ``capi.GcnbError``.

What differs from the reference, deliberately:
* Theano's MRG31k3p dropout stream cannot be reproduced without Theano; the mask comes from
  Philox4x32-10 keyed on (seed, epoch), replayable on the host (``oracle.gcn_ref.dropout_keep_mask``).
* ``f_train`` returns the full N x C output to the host every epoch (gcnmodel.py:409-410,430,
  ``all_probs``, never used); ``fit`` skips that copy.
* Multi-GPU: when ``torch.distributed`` is initialised with world_size > 1 and ``shard=True`` the
  graph is 1-D row partitioned over the ranks; every rank passes the same
  full ``X`` / ``A`` / indices and gets the same results back.
"""
from __future__ import annotations

import logging
import sys

import numpy as np
import scipy.sparse as sp

from .partition import ParamLayout

logging.basicConfig(format="%(asctime)s %(message)s", datefmt="%m/%d/%Y %I:%M:%S %p", level=logging.INFO)


def _glorot_uniform(shape):
    """lasagne.init.GlorotUniform(gain=1) on the global NumPy stream (gcnmodel.py:336,348)."""
    a = np.sqrt(6.0 / (shape[0] + shape[1]))
    return np.random.uniform(low=-a, high=a, size=shape).astype("float32")


def _orthogonal(shape):
    """lasagne.init.Orthogonal(gain=1) (gcnmodel.py:359): SVD of a standard-normal draw."""
    a = np.random.normal(0.0, 1.0, shape)
    u, _, v = np.linalg.svd(a, full_matrices=False)
    q = u if u.shape == shape else v
    return q.reshape(shape).astype("float32")


def initial_parameters(input_size, hid_size_list, output_size, highway, seed):
    """Initial weights in ``lasagne.layers.get_all_param_values`` order.

    Draw order follows layer creation order (gcnmodel.py:351-374): W0; the DropoutLayer's
    RandomStreams seed; per hidden layer Wh (Glorot) then Wt (Orthogonal) (gcnmodel.py:281-286);
    Wout.  List order puts the gate first because MultiplicativeGatingLayer's incomings are
    [gate, input1, input2] (gcnmodel.py:258).  ``np.random.seed(seed)`` is gcnmodel.py:336.
    """
    np.random.seed(seed)
    hd = hid_size_list[0]
    params = [_glorot_uniform((input_size, hd)), np.zeros(hd, "float32")]
    np.random.randint(1, 2147462579)  # lasagne DropoutLayer -> RandomStreams(get_rng().randint(...))
    prev = hd
    for i, hid in enumerate(hid_size_list):
        if i == 0:
            continue
        if highway:
            Wh = _glorot_uniform((prev, prev))
            Wt = _orthogonal((prev, prev))
            params += [Wt, np.full(prev, -4.0, "float32"), Wh, np.zeros(prev, "float32")]  # bt=-4: gcnmodel.py:274
        else:
            params += [_glorot_uniform((prev, hid)), np.zeros(hid, "float32")]
            prev = hid
    params += [_glorot_uniform((prev, output_size)), np.zeros(output_size, "float32")]
    return params


class GraphConv():
    '''
    Graph convolutional network of reference gcnmodel.py:316-477 on B200 kernels.
    Note that the input is assumed to be sparse (as in BoW model of text).
    '''

    def __init__(self, input_size, output_size, hid_size_list, regul_coef, drop_out, dtype='float32',
                 batchnorm=False, highway=True, device=None, shard=True, nonlinearity='tanh'):
        if dtype != 'float32':
            raise ValueError("the B200 path computes in float32 (gcnmain.py:167 fixes dtype to float32)")
        if len(hid_size_list) < 1:
            raise ValueError("hid_size_list needs at least one hidden size")
        self.input_size = input_size
        self.output_size = output_size
        self.hid_size_list = hid_size_list
        self.regul_coef = regul_coef
        self.drop_out = drop_out
        self.dtype = dtype
        self.dtypeint = 'int64' if self.dtype == 'float64' else 'int32'
        self.fitted = False
        self.batchnorm = batchnorm  # stored, never used (gcnmodel.py:331)
        self.highway = highway
        self.nonlinearity = nonlinearity  # tanh is the live choice (gcnmodel.py:347); relu the commented one
        self._device = device
        self._shard = shard
        self._engine = None
        self._epoch_counter = 0
        # X and A arrive as host SciPy matrices on every call (gcnmain.py:221,226,231); their device
        # copies are cached on identity.  False = copy host->device on every call (bench.py's e2e leg).
        self.cache_device_inputs = True
        logging.info('highway is {}'.format(self.highway))

    # ------------------------------------------------------------------ model construction
    def build_model(self, A, use_text=True, use_labels=True, seed=77):
        """Create the weights (gcnmodel.py:335-416).  ``A``, ``use_text``, ``use_labels`` are accepted
        and unused, like the reference (A is a call-time input there, gcnmodel.py:342)."""
        logging.info('Graphconv model input size {}, output size {} and hidden layers {} regul {} dropout {}.'.format(
            self.input_size, self.output_size, str(self.hid_size_list), self.regul_coef, self.drop_out))
        logging.info('{} gconv layers'.format(len(self.hid_size_list)))
        self.layout = ParamLayout(self.input_size, self.hid_size_list, self.output_size, self.highway)
        self.init_params = initial_parameters(self.input_size, list(self.hid_size_list), self.output_size,
                                              self.highway, seed)
        self._seed = int(seed)
        self._host_params = [p.copy() for p in self.init_params]
        if self._engine is not None:
            self._engine.set_params(self._host_params)
        self.l_out = self  # handle the caller may keep (gcnmain.py:192 ignores it)
        return self.l_out

    def _get_engine(self):
        if self._engine is None:
            import torch
            from .engine import Engine
            group = None
            if self._shard and torch.distributed.is_available() and torch.distributed.is_initialized() \
                    and torch.distributed.get_world_size() > 1:
                group = torch.distributed.group.WORLD
            self._engine = Engine(self.layout, self.drop_out, self.regul_coef, self.nonlinearity, self._device, group)
            self._engine.set_params(self._host_params)
        return self._engine

    def close(self):
        """Release the device engine (HBM buffers, NVLink peer arena).  Collective when the graph is sharded over
        several GPUs: every rank calls it at the same point.  The model can still be used afterwards (a new engine
        is created on demand from the host copy of the weights)."""
        if self._engine is not None:
            self._host_params = self._engine.get_params()
            self._engine.close()
            self._engine = None

    def invalidate_inputs(self):
        """Forget the prepared copies of X / A (pinned host staging + device CSR).  They are keyed on the identity,
        shape and nnz of the SciPy objects, so a matrix edited IN PLACE between calls needs this; new objects are
        picked up on their own."""
        if self._engine is not None:
            self._engine.unbind()

    # Lasagne get/set_all_param_values equivalents
    def get_all_param_values(self):
        if self._engine is None:
            return [p.copy() for p in self._host_params]
        return self._engine.get_params()

    def set_all_param_values(self, values):
        self.layout.check(values)
        self._host_params = [np.asarray(v, dtype=np.float32).copy() for v in values]
        if self._engine is not None:
            self._engine.set_params(self._host_params)

    # ------------------------------------------------------------------ training
    def f_train(self, X, y_train, y_dev, A, train_indices, dev_indices, seed=None, update=True):
        """One full-batch step (gcnmodel.py:409): returns [train_loss, train_acc, dev_loss, dev_acc].

        The reference also returns the N x C dropout output; read it with ``last_output()``."""
        eng = self._get_engine()
        force = not self.cache_device_inputs
        eng.bind(X, A, need_backward=True, force_upload=force)
        tr = eng.index_arrays(train_indices, y_train, force_upload=force)
        dv = eng.index_arrays(dev_indices, y_dev, force_upload=force)
        if seed is None:
            seed = (self._seed << 32) ^ self._epoch_counter
        self._epoch_counter += 1
        self._last_dropout_seed = seed
        eng.train_step(tr, dv, len(train_indices), len(dev_indices), seed, update=update)
        return list(eng.read_metrics())

    def fit(self, X, H, Y, train_indices, val_indices, n_epochs=10000, batch_size=1000, max_down=10,
            pseudolikelihood_thresh=0.2, verbose=True, seed=77):
        """Full-batch training with early stopping on the dev loss (gcnmodel.py:418-450).
        ``batch_size`` and ``pseudolikelihood_thresh`` are accepted and ignored, like the reference."""
        np.random.seed(seed)
        logging.info('training for {} epochs with batch size {}'.format(n_epochs, batch_size))
        if not sp.issparse(X):
            raise ValueError("Input for this layer must be sparse")  # gcnmodel.py:34-36
        eng = self._get_engine()
        eng.bind(X, H, need_backward=True)
        best_params = None
        best_val_loss = sys.maxsize
        best_val_acc = 0.0
        n_validation_down = 0
        report_k_epoch = 1
        Y = np.asarray(Y)
        train_indices = np.asarray(train_indices)
        val_indices = np.asarray(val_indices)
        y_train, y_dev = Y[train_indices], Y[val_indices]  # gcnmodel.py:427-428
        tr = eng.index_arrays(train_indices, y_train)
        dv = eng.index_arrays(val_indices, y_dev)
        for n in range(n_epochs):
            step_seed = (int(seed) << 32) ^ self._epoch_counter
            self._epoch_counter += 1
            eng.train_step(tr, dv, len(train_indices), len(val_indices), step_seed)
            l_train, acc_train, l_val, acc_val = eng.read_metrics()
            if l_val < best_val_loss:
                best_val_loss = l_val
                best_val_acc = acc_val
                best_params = eng.get_params()
                n_validation_down = 0
            else:
                # early stopping
                n_validation_down += 1
            if verbose:
                if n % report_k_epoch == 0:
                    logging.info('epoch {} train loss {:.2f} train acc {:.2f} val loss {:.2f} val acc {:.2f} best val acc {:.2f} maxdown {}'.format(
                        n, l_train, acc_train, l_val, acc_val, best_val_acc, n_validation_down))
            if n_validation_down > max_down and n > 2 * report_k_epoch * max_down:
                logging.info('validation results went down. early stopping ...')
                break
        self.best_params = best_params
        self.set_all_param_values(best_params)
        self.fitted = True

    # ------------------------------------------------------------------ inference
    def f_val(self, X, A, test_indices):
        eng = self._get_engine()
        if not sp.issparse(X):
            raise ValueError("Input for this layer must be sparse")
        eng.bind(X, A, need_backward=False, force_upload=not self.cache_device_inputs)
        eng.forward(train=False)
        return eng.gather_predictions(test_indices)

    def predict(self, X, A, test_indices):
        """Deterministic forward, row gather, argmax (gcnmodel.py:452-454) -> (int64[m], float32[m, C])."""
        preds_test, prob_test = self.f_val(X, A, test_indices)
        return preds_test, prob_test

    def predict_classes(self, X, A, test_indices):
        """``predict`` without the probability rows: (int64[m] on the host, the same predictions as a device int64
        tensor for ``geo.geo_eval(..., preds_device=...)``).  Nothing of size m x C crosses PCIe."""
        eng = self._get_engine()
        if not sp.issparse(X):
            raise ValueError("Input for this layer must be sparse")
        eng.bind(X, A, need_backward=False, force_upload=not self.cache_device_inputs)
        eng.forward(train=False)
        return eng.gather_predictions(test_indices, want_probs=False)

    def last_output(self):
        """N x C output of the most recent forward (the 5th output of the reference's f_train)."""
        eng = self._get_engine()
        return eng.read_matrix(eng.P, eng.n, self.output_size)

    def reset(self):
        """Restore the initial weights (gcnmodel.py:456-457).  Like the reference, the Adam
        moments and step count are shared state of the update rule and are NOT reset."""
        self.set_all_param_values(self.init_params)

    def save(self, dumper, filename='./model.pkl'):
        if self.fitted:
            logging.info('dumping model params in {}'.format(filename))
            dumper(self.best_params, filename)
        else:
            logging.warning('The model is not trained yet!')

    def load(self, loader, filename):
        logging.info('loading the model from {}'.format(filename))
        self.best_params = loader(filename)
        self.set_all_param_values(self.best_params)
        self.fitted = True

    def get_gates(self, X, A):
        """Deterministic gate activations, one N x Hd array per highway layer (gcnmodel.py:396-401,472-477)."""
        eng = self._get_engine()
        eng.bind(X, A, need_backward=False)
        eng.forward(train=False)
        return eng.gates()


# ----------------------------------------------------------------------------------------------
# Layer-level surface.  north_star names "SparseConvolutionLayer"; the reference's closest classes
# are SparseConvolutionDenseLayer(2) (gcnmodel.py:72-92,224-249), neither used by GraphConv.
# They are re-compositions of the same two kernels, exposed here as functional shims over host
# arrays for callers that want a single layer; each runs on the GPU through the C ABI.
# ----------------------------------------------------------------------------------------------
class _SingleLayer:
    def __init__(self, num_units, nonlinearity='tanh', device=None):
        self.num_units = int(num_units)
        self.nonlinearity = nonlinearity
        self._device = device
        self.W = None
        self.b = None

    def _init(self, n_in):
        if self.W is None:
            self.W = _glorot_uniform((n_in, self.num_units))
            self.b = np.zeros(self.num_units, "float32")


class SparseInputDenseLayer(_SingleLayer):
    """act(X.W + b) for sparse X (gcnmodel.py:29-42)."""

    def get_output_for(self, input, **kwargs):
        if not sp.issparse(input):
            raise ValueError("Input for this layer must be sparse")
        from .layers import sparse_dense
        self._init(input.shape[1])
        return sparse_dense(input, self.W, self.b, self.nonlinearity, device=self._device)


def _take_rows(out, target_indices, use=True):
    """``activation[target_indices, :]`` (gcnmodel.py:110,199,218): row gather of the host result."""
    if use and target_indices is not None and len(target_indices):
        return out[np.asarray(target_indices)]
    return out


class SparseInputDropoutLayer:
    """Dropout on a sparse input (gcnmodel.py:44-70): stored entries are kept with probability 1-p and rescaled by
    1/(1-p); ``deterministic`` or p == 0 returns the input.  The keep mask is drawn on the device with the engine's
    Philox stream (one draw per stored entry), not Theano's MRG stream."""

    def __init__(self, p=0.5, rescale=True, device=None):
        self.p, self.rescale, self._device = float(p), bool(rescale), device

    def get_output_for(self, input, deterministic=False, seed=0, **kwargs):
        if not sp.issparse(input):
            raise ValueError("Input for this layer must be sparse")
        if deterministic or self.p == 0:
            return input
        from .layers import dropout_keep_mask
        X = sp.csr_matrix(input, dtype=np.float32, copy=True)
        keep = dropout_keep_mask(1, X.nnz, self.p, seed, device=self._device)[0]
        scale = np.float32(1.0 / (1.0 - self.p)) if self.rescale else np.float32(1.0)
        X.data = X.data * scale * keep.astype(np.float32)
        return X


class SparseConvolutionDenseLayer(_SingleLayer):
    """act(A.(X.W) + b) for sparse X and sparse A (gcnmodel.py:72-92); ``A`` is a constructor argument there and may be
    given at either place here.  Without any A the convolution is skipped (gcnmodel.py:242)."""

    def __init__(self, num_units, nonlinearity='tanh', device=None, A=None):
        super().__init__(num_units, nonlinearity, device)
        self.A = A

    def get_output_for(self, input, A=None, **kwargs):
        if not sp.issparse(input):
            raise ValueError("Input for this layer must be sparse")
        from .layers import sparse_dense, graph_conv_dense
        self._init(input.shape[1])
        A = self.A if A is None else A
        if A is None:
            return sparse_dense(input, self.W, self.b, self.nonlinearity, device=self._device)
        return graph_conv_dense(A, sparse_dense(input, self.W, None, 'linear', device=self._device), None, self.b,
                                self.nonlinearity, device=self._device)


SparseConvolutionLayer = SparseConvolutionDenseLayer  # the name BASELINE.json's north_star uses


class SparseConvolutionDenseLayer2(SparseConvolutionDenseLayer):
    """The call-time-A variant (gcnmodel.py:224-249): ``get_output_for(input, A=...)``; falsy A skips the convolution."""

    def __init__(self, num_units, nonlinearity='tanh', device=None, use_target_indices=False):
        super().__init__(num_units, nonlinearity, device)
        self.use_target_indices = use_target_indices


class ConvolutionDenseLayer2(_SingleLayer):
    """act((A.(x.W) + b)[target_indices, :]) for dense x (gcnmodel.py:114-136).  The row gather only happens when the
    layer was built with ``use_target_indices=True`` AND ``target_indices`` is passed (gcnmodel.py:121-123,134-135); the
    activations are element-wise or row-wise, so gathering the rows of the activated matrix is the same thing."""

    def __init__(self, num_units, nonlinearity='tanh', device=None, use_target_indices=False):
        super().__init__(num_units, nonlinearity, device)
        self.use_target_indices = use_target_indices

    def get_output_for(self, input, A=None, target_indices=None, **kwargs):
        from .layers import graph_conv_dense
        self._init(np.shape(input)[1])
        out = graph_conv_dense(A, np.asarray(input, dtype=np.float32), self.W, self.b, self.nonlinearity,
                               device=self._device)
        return _take_rows(out, target_indices, self.use_target_indices)


class ConvolutionDenseLayer3(ConvolutionDenseLayer2):
    """softmax(A.(x.W) + b) (gcnmodel.py:138-157)."""

    def __init__(self, num_units, nonlinearity='softmax', device=None):
        super().__init__(num_units, nonlinearity, device)

    def get_output_for(self, input, A=None, **kwargs):  # no target_indices in the reference's signature (gcnmodel.py:148)
        return super().get_output_for(input, A=A)


class ConvolutionDenseLayer_zero(ConvolutionDenseLayer2):
    """act(A.(x.W) + b) with A a constructor argument (gcnmodel.py:159-179)."""

    def __init__(self, num_units, nonlinearity='tanh', device=None, A=None):
        super().__init__(num_units, nonlinearity, device)
        self.A = A

    def get_output_for(self, input, **kwargs):
        return super().get_output_for(input, A=self.A)


class ConvolutionDenseLayer(ConvolutionDenseLayer_zero):
    """act((A.(x.W) + b)[target_indices, :]) (gcnmodel.py:94-112).  The activation is element-wise, so gathering the
    rows of the activated matrix is the same thing."""

    def get_output_for(self, input, target_indices=None, **kwargs):
        return _take_rows(super().get_output_for(input), target_indices)


class ConvolutionLayer:
    """act((A.x)[target_indices, :]): a graph convolution without weights (gcnmodel.py:181-201)."""

    def __init__(self, use_target_indices=False, A=None, nonlinearity='linear', device=None):
        self.use_target_indices, self.A, self.nonlinearity, self._device = use_target_indices, A, nonlinearity, device

    def get_output_for(self, input, target_indices=None, **kwargs):
        from .layers import spmm
        out = spmm(self.A, np.asarray(input, dtype=np.float32), act=self.nonlinearity or 'linear', device=self._device)
        return _take_rows(out, target_indices, self.use_target_indices)


class DenseLayer2(_SingleLayer):
    """act((x.W + b)[target_indices, :]) (gcnmodel.py:203-221)."""

    def __init__(self, num_units, nonlinearity='tanh', device=None, use_target_indices=False):
        super().__init__(num_units, nonlinearity, device)
        self.use_target_indices = use_target_indices

    def get_output_for(self, input, target_indices=None, **kwargs):
        from .layers import gemm
        self._init(np.shape(input)[1])
        out = gemm(np.asarray(input, dtype=np.float32), self.W, bias=self.b, act=self.nonlinearity or 'linear',
                   device=self._device)
        return _take_rows(out, target_indices, self.use_target_indices)


class MultiplicativeGatingLayer:
    """y = t*h1 + (1-t)*h2 for inputs [t, h1, h2] of equal shape (gcnmodel.py:252-266)."""

    def __init__(self, device=None):
        self._device = device

    def get_output_for(self, inputs, **kwargs):
        t, h1, h2 = (np.asarray(v, dtype=np.float32) for v in inputs)
        assert t.shape == h1.shape == h2.shape  # gcnmodel.py:260
        from .layers import gate_mix
        return gate_mix(t, h1, h2, device=self._device)


def highway_dense(x, A=None, gconv=False, Wh=None, bh=None, Wt=None, bt=None, nonlinearity='sigmoid', device=None):
    """Highway layer on host arrays (gcnmodel.py:268-288): h = act(conv(x.Wh) + bh) with conv = A. when ``gconv``,
    t = sigmoid(x.Wt + bt), returns (t*h + (1-t)*x, t).  Default initialisers like the reference: Wh, Wt Glorot
    uniform, bh = 0, bt = -4 (gcnmodel.py:271-274); Wh is drawn before Wt.  Note the reference's default activation of
    the h branch here is the sigmoid; ``GraphConv`` passes its own (gcnmodel.py:361)."""
    from .layers import highway, spmm
    x = np.asarray(x, dtype=np.float32)
    n_in = x.shape[1]
    Wh = _glorot_uniform((n_in, n_in)) if Wh is None else Wh
    Wt = _glorot_uniform((n_in, n_in)) if Wt is None else Wt
    bh = np.zeros(n_in, "float32") if bh is None else bh
    bt = np.full(n_in, -4.0, "float32") if bt is None else bt
    S = spmm(A, x, device=device) if gconv else x
    y, _, t = highway(S, x, Wh, bh, Wt, bt, act=nonlinearity, device=device)
    return y, t


def residual_dense(x, A, W=None, b=None, nonlinearity='selu', device=None):
    """nonlinearity(A.(x.W) + b + x) (gcnmodel.py:290-294): one GEMM, then one SpMM whose epilogue adds the residual
    and applies the SELU."""
    from .layers import gemm, spmm
    x = np.asarray(x, dtype=np.float32)
    n_in = x.shape[1]
    W = _glorot_uniform((n_in, n_in)) if W is None else W
    b = np.zeros(n_in, "float32") if b is None else b
    q = gemm(x, W, device=device)
    return spmm(A, q, bias=b, act=nonlinearity or 'linear', accumulate_into=x, accumulate_mode=2, device=device)


def np_softmax(x):
    """Softmax over ALL entries of ``x`` jointly, which is what gcnmodel.py:298-301 computes (no axis argument)."""
    x = np.asarray(x)
    shifted = np.exp(x - x.max())
    return shifted / shifted.sum()


def iterate_minibatches(inputs, targets, batchsize, shuffle=False):
    """Yield (inputs, targets) slices of ``batchsize`` rows; a trailing partial batch is dropped and ``shuffle`` draws
    one permutation from the global NumPy stream (behaviour of gcnmodel.py:303-313; ``GraphConv`` trains full-batch and
    never calls it)."""
    n = inputs.shape[0]
    assert n == targets.shape[0]
    order = None
    if shuffle:
        order = np.arange(n)
        np.random.shuffle(order)
    for lo in range(0, n - batchsize + 1, batchsize):
        sel = slice(lo, lo + batchsize) if order is None else order[lo:lo + batchsize]
        yield inputs[sel], targets[sel]
