"""Test stand-in: data.py:21 imports matplotlib.collections at module level; nothing on the GCN path uses it."""
