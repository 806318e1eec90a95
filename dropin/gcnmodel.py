"""Drop-in replacement for the reference's ``gcnmodel`` module.

Put this directory on ``sys.path`` ahead of the reference checkout and the unchanged
``gcnmain.py`` (``from gcnmodel import GraphConv``, gcnmain.py:34) runs the GCN hot path on the
B200 kernels of ``geographconv_b200``.  See INTEGRATION.md.
"""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from geographconv_b200.gcnmodel import (  # noqa: E402,F401
    GraphConv, SparseInputDenseLayer, SparseInputDropoutLayer, SparseConvolutionDenseLayer, SparseConvolutionLayer,
    SparseConvolutionDenseLayer2, ConvolutionDenseLayer, ConvolutionDenseLayer2, ConvolutionDenseLayer3,
    ConvolutionDenseLayer_zero, ConvolutionLayer, DenseLayer2, MultiplicativeGatingLayer, highway_dense, residual_dense,
    np_softmax, iterate_minibatches, initial_parameters)
