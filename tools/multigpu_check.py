#!/usr/bin/env python
"""Row-partitioned (world_size > 1) parity check, launched by torchrun on one box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/multigpu_check.py

Every rank passes the same full X / A / indices to GraphConv (shard=True).  For both exchange designs (feature-sliced
over NVLink peer memory, all-gather over NCCL) and a highway and a plain network, rank 0 checks

* the forward (probabilities of every row, predictions) BIT-IDENTICAL to a single-GPU run of the same library
  (same hot columns on every rank, every row summed in CSR order);
* predict and one training step (losses, every gradient) against the NumPy oracle at 1e-3 and against the single-GPU
  gradients to fp32 rounding (weight gradients are sums over nodes: their association differs by design).

Prints MULTIGPU_CHECK PASS / FAIL (kept in gpurun_out/ as the driver cannot run a 2-GPU pytest).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from geographconv_b200 import synth
    from geographconv_b200.gcnmodel import GraphConv
    from oracle import gcn_ref

    cfg = dict(n=3001, deg=9, f=700, xnnz=30, hid=[96, 96, 96], classes=37)
    A, X, Y, tr, dev, te, _ = synth.synthetic_problem(cfg, seed=5)
    # a few hub rows longer than one row item (1024 nonzeros) exercise the partial-sum / fix-up path of the sliced product
    import scipy.sparse as sp
    Ah = A.tolil()
    rng = np.random.RandomState(0)
    for hub in (17, 2500):
        cols = rng.choice(cfg["n"], size=1500, replace=False)
        Ah[hub, cols] = 0.01
        Ah[cols, hub] = 0.01
    A = sp.csr_matrix(Ah, dtype=np.float32)
    A.sort_indices()
    all_rows = np.arange(cfg["n"], dtype=np.int32)
    ok = True
    modes = os.environ.get("GCNB_CHECK_MODES", "slice,gather").split(",")
    for mode in modes:
        os.environ["GCNB_EXCHANGE"] = mode
        for highway, hid in [(True, [96, 96, 96]), (False, [96, 64, 80])]:
            clf = GraphConv(cfg["f"], cfg["classes"], hid, 0.0, 0.5, highway=highway, device=local, shard=True)
            clf.build_model(A, seed=3)
            params = [p.copy() for p in clf.init_params]
            preds, probs = clf.predict(X, A, all_rows)
            seed = 99
            out = clf.f_train(X, Y[tr], Y[dev], A, tr, dev, seed=seed, update=False)
            eng = clf._get_engine()
            assert eng.world == world
            used = eng.exchange
            train_probs = clf.last_output()
            grads = eng.get_grads()
            if rank == 0:
                one = GraphConv(cfg["f"], cfg["classes"], hid, 0.0, 0.5, highway=highway, device=local, shard=False)
                one.build_model(A, seed=3)
                p1, pr1 = one.predict(X, A, all_rows)
                out1 = one.f_train(X, Y[tr], Y[dev], A, tr, dev, seed=seed, update=False)
                tp1 = one.last_output()
                g1 = one._get_engine().get_grads()
                bit_equal = bool(np.array_equal(pr1, probs) and np.array_equal(p1, preds) and np.array_equal(tp1, train_probs))
                max_diff = float(np.abs(pr1.astype(np.float64) - probs).max())
                rp, rprob = gcn_ref.predict(params, X, A, all_rows, hid, highway, dtype="float64")
                np.testing.assert_allclose(probs, rprob, rtol=1e-3, atol=1e-7)
                keep = gcn_ref.dropout_keep_mask(seed, cfg["n"], hid[0], 0.5)
                r = gcn_ref.loss_and_grads(params, X, A, Y, tr, hid, highway, keep.astype(np.float32) / 0.5, 0.0,
                                           dtype="float64", dev_idx=dev)
                np.testing.assert_allclose(out[0], r["train_loss"], rtol=1e-3)
                np.testing.assert_allclose(out[2], r["dev_loss"], rtol=1e-3)
                for g, rg, gs in zip(grads, r["grads"], g1):
                    np.testing.assert_allclose(g, rg, rtol=1e-3, atol=1e-4 * float(np.abs(rg).max()) + 1e-12)
                    np.testing.assert_allclose(g, gs, rtol=1e-4, atol=1e-5 * float(np.abs(gs).max()) + 1e-12)
                print("world=%d exchange=%s(%s) highway=%s: forward bit-identical to 1 GPU: %s (max |diff| %.3g); oracle parity "
                      "ok; loss %.6f vs 1-GPU %.6f" % (world, mode, used, highway, bit_equal, max_diff, out[0], out1[0]),
                      flush=True)
                ok = ok and bit_equal and used == mode
                del one
            del eng
            dist.barrier()
            clf.close()
            dist.barrier()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTIGPU_CHECK", "PASS" if ok else "FAIL", flush=True)
        sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
