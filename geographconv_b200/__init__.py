"""geographconv_b200 -- the GCN forward/backward hot path of afshinrahimi/geographconv on B200.

Python host (this package) -> ctypes -> ``libgcnb200.so`` (hand-written sm_100a CUDA behind the C
ABI of ``include/gcnb200.h``).  ``gcnmodel.GraphConv`` mirrors the reference class of the same
name; ``dropin/gcnmodel.py`` re-exports it under the reference's module name.  Importing the
package needs neither a GPU nor the built library; using it does, and fails loudly otherwise.
"""
from .capi import GcnbError, load_library, LIB_PATH  # noqa: F401
from .partition import ParamLayout  # noqa: F401

__all__ = ["GcnbError", "load_library", "LIB_PATH", "ParamLayout", "GraphConv"]


def __getattr__(name):
    if name == "GraphConv":
        from .gcnmodel import GraphConv
        return GraphConv
    raise AttributeError(name)
