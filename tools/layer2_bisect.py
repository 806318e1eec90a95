#!/usr/bin/env python
"""Where does the gemm_v 2 forward start to depend on its context?  One GPU, workload C3.

With GCNB_GEMM_V=2 the forward checksum of a 1-GPU run differs from the 2- and 8-GPU runs from the second highway
layer on (Y1 / H1 / T1 equal, Y2 / H2 / T2 not; profiles/r2n_*.json) although every kernel involved is deterministic and
row-count independent when called alone (tools/gemm_rows_check.py).  This tool takes the buffers of a real forward and

* recomputes S2 = A_hat . Y1 with each SpMM engine (bit patterns must agree),
* re-runs the fused highway kernel of layer 2 on (S2, Y1) alone -- all rows, and the first half of the rows the way rank
  0 of a 2-GPU run sees them -- and compares H2 / T2 / Y2 with what the forward left in the layer buffers.

    GCNB_GEMM_V=2 python tools/layer2_bisect.py
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from geographconv_b200 import adjacency, synth  # noqa: E402
from geographconv_b200.engine import _ptr  # noqa: E402
from geographconv_b200.gcnmodel import GraphConv  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:  # under torchrun: the same checks on the row-partitioned run (checksums are summed over the ranks)
    import torch.distributed as dist
    from geographconv_b200.partition import row_blocks
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank != 0:
        sys.stdout = open(os.devnull, "w")
cfg = dict(synth.CONFIGS["C3"])
rr = row_blocks(cfg["n"], world)[1][rank] if world > 1 else None
A, X, Y, tr, dev, te, _ = synth.synthetic_problem(
    cfg, seed=77, row_range=rr,
    graph_builder=lambda u, v, n: adjacency.normalized_adjacency_from_edges(u, v, n, device=local))
clf = GraphConv(cfg["f"], cfg["classes"], cfg["hid"], regul_coef=0.0, drop_out=0.5, highway=True, device=local, shard=True)
clf.build_model(A, seed=77)
eng = clf._get_engine()
eng.bind(X, A, need_backward=True, assume_symmetric=True)
print("gemm_v", eng.ctx.get_option("gemm_v"), "world", world, "exchange", eng.exchange, flush=True)
hd = cfg["hid"][0]
n, ld = eng.n_loc, eng.ldh[0]


def total(count):
    """Sum a per-rank count over the ranks."""
    t = torch.tensor([int(count)], dtype=torch.int64, device=eng.dev)
    if world > 1:
        dist.all_reduce(t)
    return int(t.item())


eng.forward(train=True, seed=424242)
eng.ctx.sync()
lay2 = eng.lay[1]
Y1 = eng.lay[0]["Y"]
bits = lambda t, rows=n: t[:rows, :hd].contiguous().view(torch.int32)
ref = {k: bits(lay2[k]).clone() for k in ("Y", "H", "T")}
print("forward checksums: Y1 %012x  Y2 %012x H2 %012x T2 %012x" % (
    eng.checksum(Y1, n, hd), eng.checksum(lay2["Y"], n, hd), eng.checksum(lay2["H"], n, hd), eng.checksum(lay2["T"], n, hd)),
    flush=True)

# S2 with every SpMM engine
S = eng.S.view(-1)[: eng.nbuf * ld].view(eng.nbuf, ld)
s_bits = {}
for engine in ((-2, 0, 1, 2) if world == 1 else (-2,)):
    eng.A.struct.engine = engine
    eng._conv(Y1, eng.A, S, ld, hd)
    eng.ctx.sync()
    s_bits[engine] = bits(S).clone()
    print("S2 with SpMM engine %2d: checksum %012x, elements differing from the default engine: %d" % (
        engine, eng.checksum(S, n, hd), total((s_bits[engine] != s_bits[-2]).sum())), flush=True)
eng.A.struct.engine = -2
eng._conv(Y1, eng.A, S, ld, hd)

# the layer-2 highway kernel alone
k = eng.layout.layers[1]["i"]
Wh, ldwh = eng._pptr("Wh%d" % k)
bh, _ = eng._pptr("bh%d" % k)
Wt, ldwt = eng._pptr("Wt%d" % k)
bt, _ = eng._pptr("bt%d" % k)

for rows in ((n, n // 2, n // 8) if world == 1 else (n,)):  # n = rows of this rank
    out = [torch.zeros(rows, ld, device=eng.dev) for _ in range(3)]
    eng._fence()
    eng.ctx.call("gcnb_highway_fwd_f32", rows, hd, _ptr(S), ld, _ptr(Y1), ld, Wh, ldwh, bh, Wt, ldwt, bt, eng.act,
                 _ptr(out[0]), ld, _ptr(out[1]), ld, _ptr(out[2]), ld)
    eng.ctx.sync()
    for name, t in zip(("Y", "H", "T"), out):
        d = total((bits(t, rows) != ref[name][:rows]).sum())
        print("highway layer 2 alone on rows [0, %d): %s2 differs from the forward's in %d elements" % (rows, name, d), flush=True)
# and the forward once more: is it reproducible inside one process?
eng.forward(train=True, seed=424242)
eng.ctx.sync()
for name in ("Y", "H", "T"):
    print("second forward: %s2 differs from the first in %d elements" % (name, total((bits(lay2[name]) != ref[name]).sum())))
