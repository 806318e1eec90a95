#!/usr/bin/env python
"""bench.py -- GCN fwd+bwd+Adam step throughput (nodes/s) on synthetic graphs of BASELINE.json's shapes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C3] [--impl reference]

One "step" = one ``f_train``-equivalent call (reference gcnmodel.py:409,429-430): full-graph forward
with dropout, cross-entropy metrics, backward, Adam.  Prints ONE JSON line (rank 0).

* ``value``    nodes/s with X, A_hat, labels and weights already resident in HBM (device-timed,
               CUDA events on the stream the kernels run on, max over ranks).
* ``e2e``      the same step through the public ``GraphConv.f_train`` call with HOST buffers: every
               step copies X, X^T, A_hat (CSR arrays + row plans) and the index / label arrays from
               pinned host memory to the device and reads the four metrics back.
* ``roofline`` the dominant kernel, the A_hat.H CSR SpMM at the hidden width: algorithmic bytes
               B_touch / average launch duration from CUDA events recorded around
               every such launch inside the timed region, against MEASURED_PEAKS.json's HBM copy rate.
* ``cpu_baseline`` the oracle port (SciPy csr_matvecs single-threaded like Theano's sd_csr + BLAS
               on all cores) on a bounded sample: a smaller graph with the same per-node statistics
               (the whole graph when the workload is small enough: C1), plus one full-size A_hat.H
               product and one full-size dense product timed on the host and the op-count
               extrapolation to the full configuration.
* ``parity``   in-run proof that THIS run computes the reference's numbers, at every N: a forward
               pass with the initial weights, 2000 sampled rows of one A_hat.H product and of the
               final probabilities recomputed in float64 with SciPy from the device's own layer
               inputs (max relative error, argmax mismatches), and a checksum of the bit patterns
               of all probabilities that is identical at 1/2/4/8 GPUs iff the row-partitioned
               forward is bit-identical to the single-GPU one.

``--impl reference`` times that CPU port alone (rank 0 only); Theano/Lasagne cannot be installed
here, so ``kind`` is "port".  N > 1: launched by torchrun, one rank per GPU, rows of
A_hat / X / activations partitioned, one all-gather of the dense operand per graph convolution.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "gcn_fwd_bwd_nodes_per_sec"
UNIT = "nodes/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def workload_description(name, cfg):
    return ("%s: synthetic N=%d nodes, avg-degree %d A_hat, %d BoW feats (%d nnz/row), hidden %s %s, %d classes, "
            "dropout 0.5, Adam" % (name, cfg["n"], cfg["deg"], cfg["f"], cfg["xnnz"], "x".join(map(str, cfg["hid"])),
                                   "highway", cfg["classes"]))


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
                pw.append(float(parts[3]))
            except ValueError:
                continue
            for nme, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm), "power_w_max": float(max(pw))}
        return out


# ------------------------------------------------------------------------------------------------
# CPU port (oracle) timing -- cpu_baseline leg and --impl reference
# ------------------------------------------------------------------------------------------------
def cpu_sample_problem(cfg, n_sample):
    from geographconv_b200 import synth
    c = dict(cfg)
    c["n"] = int(n_sample)
    return synth.synthetic_problem(c, seed=77)


def cpu_step_seconds(problem, steps, warmup):
    """Seconds per f_train-equivalent step of the oracle port on the host cores."""
    from oracle import gcn_ref
    A, X, Y, tr, dev, te, cfg = problem
    params = gcn_ref.init_params(cfg["f"], cfg["hid"], cfg["classes"], True, 77)
    state = gcn_ref.AdamState(params)
    rng = np.random.RandomState(0)
    times = []
    for i in range(warmup + steps):
        scale = (rng.random_sample((cfg["n"], cfg["hid"][0])) < 0.5).astype(np.float32) / np.float32(0.5)
        t0 = time.perf_counter()
        params, _ = gcn_ref.train_step(params, state, X, A, Y, tr, dev, cfg["hid"], True, scale)
        t1 = time.perf_counter()
        if i >= warmup:
            times.append(t1 - t0)
    return times


def cpu_full_size_ops(cfg, A):
    """One A_hat.H product (SciPy csr_matvecs, one thread = Theano's sd_csr loop) and one N x Hd x Hd dense product
    (BLAS, all cores) at the workload's FULL size, and the op-count extrapolation of a whole training step from them
    (SURVEY.md 8d).  Sparse products scale with nnz x columns, dense ones with flops."""
    n, hd, C = cfg["n"], cfg["hid"][0], cfg["classes"]
    L = len(cfg["hid"]) - 1
    rng = np.random.RandomState(1)
    H = rng.standard_normal((n, hd)).astype(np.float32)
    W = rng.standard_normal((hd, hd)).astype(np.float32)
    t0 = time.perf_counter()
    S = A @ H
    t_spmm = time.perf_counter() - t0
    t0 = time.perf_counter()
    H @ W
    t_gemm = time.perf_counter() - t0
    del S
    per_nnz_col = t_spmm / (A.nnz * hd)
    per_flop = t_gemm / (2.0 * n * hd * hd)
    nnz_x = n * cfg["xnnz"]
    # forward: X.W0, L A-products at Hd, one at C; backward: the same A-products again and X^T.dz
    sparse = per_nnz_col * (2 * nnz_x * hd + 2 * L * A.nnz * hd + 2 * A.nnz * C)
    # dense: forward 2 products per highway layer + output; backward twice that
    dense = per_flop * 3 * (L * 2 * 2.0 * n * hd * hd + 2.0 * n * hd * C)
    return {"spmm_a_full_s": t_spmm, "gemm_full_s": t_gemm, "extrapolated_step_s": sparse + dense,
            "extrapolated_nodes_per_s": n / (sparse + dense),
            "how": "one A_hat.H (N=%d, nnz=%d, K=%d) and one N x %d x %d product timed on the host; step = op counts x those "
                   "rates (sparse ~ nnz x columns, dense ~ flops)" % (n, A.nnz, hd, hd, hd)}


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        return os.cpu_count() or 1


def run_reference(args, cfg, wname):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_s = min(args.cpu_sample_nodes, cfg["n"])
    prob = cpu_sample_problem(cfg, n_s)
    times = cpu_step_seconds(prob, args.steps, args.warmup)
    sec = float(np.mean(times))
    val = n_s / sec
    cores = blas_threads()
    sample = ("oracle port (SciPy csr_matvecs 1 thread = Theano sd_csr; OpenBLAS sgemm %d threads) on a %d-node "
              "graph with the workload's per-node statistics; one step = fwd+bwd+Adam" % (cores, n_s))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_description(wname, cfg), "timing": "host perf_counter",
                       "sample_nodes": n_s, "same_config": n_s == cfg["n"]},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "host_cpus": len(os.sched_getaffinity(0))},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu(args, cfg, wname):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (there is no CPU fallback); use --impl reference for the CPU port")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from geographconv_b200 import synth
    from geographconv_b200.gcnmodel import GraphConv

    from geographconv_b200.partition import row_blocks
    t_gen = time.time()
    # each rank generates the rows of X it owns (rows outside its block stay empty)
    rr = row_blocks(cfg["n"], world)[1][rank] if world > 1 else None
    # A_hat comes out of the library's GPU builder (csrc/adjacency.cu; bit-identical to the NumPy restatement of
    # gcnmain.py:115-128, tests/test_gpu_adjacency.py) -- the host np.unique over 2 x 63M edge keys of C4 takes minutes
    from geographconv_b200 import adjacency
    A, X, Y, tr, dev, te, _ = synth.synthetic_problem(
        cfg, seed=77, alpha=args.alpha, row_range=rr,
        graph_builder=None if args.host_graph else
        (lambda u, v, n: adjacency.normalized_adjacency_from_edges(u, v, n, device=local_rank)))
    t_gen = time.time() - t_gen
    N = cfg["n"]
    clf = GraphConv(cfg["f"], cfg["classes"], cfg["hid"], regul_coef=0.0, drop_out=0.5, highway=True,
                    device=local_rank, shard=True)
    clf.build_model(A, seed=77)
    eng = clf._get_engine()
    if args.spmm_variant is not None:
        eng.ctx.set_option("spmm_variant", args.spmm_variant)
    if args.gemm_tc is not None:
        eng.ctx.set_option("gemm_tc", args.gemm_tc)
        eng.side_ctx.set_option("gemm_tc", args.gemm_tc)
    y_tr, y_dev = Y[tr], Y[dev]
    t_bind = time.time()
    eng.bind(X, A, need_backward=True, assume_symmetric=True)  # symmetric by construction (synth.synthetic_graph)
    d_tr = eng.index_arrays(tr, y_tr)
    d_dev = eng.index_arrays(dev, y_dev)
    eng.ctx.sync()
    t_bind = time.time() - t_bind

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(i):
        eng.train_step(d_tr, d_dev, len(tr), len(dev), seed=1000 + i)

    hd = cfg["hid"][0]
    # ---- in-run parity block (initial weights; before anything is trained) ----
    barrier()  # host preparation takes a different time on every rank; the device barriers of the step are short-fused
    parity = parity_block(eng, clf, A, cfg, rank, args.parity_rows) if args.parity_rows > 0 else None

    # ---- device-resident leg (value) ----
    # clocks / throttle reasons are sampled every 100 ms from the warm-up to the end of the timed steps (a C1 step is
    # 1 ms: the timed region alone is shorter than one sample period)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        if ms_estimate_short(cfg):
            for i in range(200):  # keep the GPU under load for a few sample periods
                step_resident(-1 - i)
    for i in range(args.warmup):
        step_resident(i)
    eng.read_metrics()
    barrier()
    # inside the timed region only the dominant kernel (the A_hat.H product at the hidden width) carries CUDA events;
    # the full per-op split comes from an instrumented pass of the same steps afterwards (two event records around every
    # one of the ~80 ops of a step are measurable at 7 ms per step)
    eng.prof_mask(1 << capi_tag("spmm_a"))
    eng.prof_enable(True)
    eng.prof_reset()
    launches0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_host0 = time.perf_counter()
    e0.record(eng.stream)  # the stream every kernel of the step is launched on
    for i in range(args.steps):
        step_resident(args.warmup + i)
    e1.record(eng.stream)
    t_issue = time.perf_counter() - t_host0  # host time to enqueue the steps (must stay below the device time)
    barrier()
    clocks = sampler.stop() if sampler else None
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - launches0
    prof = eng.prof_collect()
    metrics = eng.read_metrics()
    tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    ms_per_step = ms / args.steps
    value = N / (ms_per_step * 1e-3)

    # roofline of the dominant kernel: A_hat.H SpMM at the hidden width (tag spmm_a)
    spmm_ms, spmm_ops = prof["spmm_a"]
    peak, peak_kind = load_peaks()
    b_touch = eng.conv_touched_bytes(hd)
    w_mine = int(eng._slice_plan(hd)[1][eng.rank]) if eng.exchange == "slice" else hd
    n_rows_prod = eng.A.shape[0]
    b_min = eng.A.nnz * 8 + (n_rows_prod + 1) * 4 + (eng.A.shape[1] + n_rows_prod) * w_mine * 4
    pieces = len(eng._panels(hd, eng.ldh[0], False)) if eng.exchange == "gather" else 1  # SpMM launches per convolution
    roof = None
    kname = "spmm_panel_kernel" if eng.exchange == "slice" else \
        ["spmm_ldg_kernel", "spmm_bulk_kernel", "spmm_panel_kernel"][eng.A.engine_for(eng, eng.ldh[0], hd)]
    traffic, l2_hit, traffic_src = args.ncu_traffic_bytes, None, "--ncu-traffic-bytes" if args.ncu_traffic_bytes else None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if traffic is None and world == 1 and os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("workload") == wname and args.alpha is None and kname in tj.get("kernel", ""):
            traffic, l2_hit = tj["dram_bytes_per_launch"], tj.get("l2_hit_pct")  # committed ncu --set full capture
            traffic_src = tj.get("source")
    if spmm_ops:
        t_launch = spmm_ms / (spmm_ops / pieces) * 1e-3  # all pieces of one product
        ach = b_touch / t_launch / 1e9
        tms2 = torch.tensor([ach], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tms2, op=dist.ReduceOp.MIN)  # slowest rank's kernel
        scale = float(tms2.item()) / ach
        ach = float(tms2.item())
        t_launch = t_launch / scale
        roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic,
                "kernel": "%s (A_hat.H, K=%d%s)" % (kname, hd, ", column slice of %d" % w_mine if eng.exchange == "slice" else ""),
                "algorithmic_bytes_per_launch": b_touch, "compulsory_bytes_per_launch": b_min,
                "launches_timed": spmm_ops, "launches_per_product": pieces,
                "avg_launch_ms": t_launch * 1e3, "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)",
                "per_rank": world > 1,
                # the physical side of the same launch: DRAM bytes ncu counted / the live launch time
                "frac_dram": (traffic / t_launch / 1e9 / peak) if traffic else None,
                "frac_compulsory": b_min / t_launch / 1e9 / peak,
                "l2_hit_pct": l2_hit, "traffic_source": traffic_src}
        if kname == "spmm_panel_kernel":
            roof["note"] = ("L2-resident column panels: B leaves HBM once per product and the per-nonzero gathers are L2 hits, so "
                            "`frac` (algorithmic B_touch / t over the HBM peak, the contract's definition) may exceed 1; "
                            "`frac_dram` is ncu's DRAM bytes over the same time, `frac_compulsory` the perfect-reuse bytes")

    # ---- instrumented pass: per-op split (every tag timed), NCCL collectives timed with their own events ----
    split_steps = max(1, min(args.steps, 5))
    eng.prof_mask(-1)
    eng.prof_reset()
    eng.time_nccl = True
    eng.nccl_ms()
    barrier()
    for i in range(split_steps):
        step_resident(args.warmup + args.steps + i)
    barrier()
    prof_all = eng.prof_collect()
    eng.prof_enable(False)
    split = {k: round(v[0] / split_steps, 4) for k, v in prof_all.items()}
    split["nccl"] = round(eng.nccl_ms() / split_steps, 4)
    eng.time_nccl = False
    split["comm_total"] = round(split.get("comm", 0.0) + split.get("sync", 0.0) + split["nccl"], 4)

    # ---- end-to-end leg through GraphConv.f_train with host buffers ----
    e2e_steps = max(0, min(args.steps, args.e2e_steps))
    e2e = None
    if e2e_steps > 0:
        clf.cache_device_inputs = False
        clf.f_train(X, y_tr, y_dev, A, tr, dev, seed=1)  # warm (pinned buffers exist, device buffers reused)
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            clf.f_train(X, y_tr, y_dev, A, tr, dev, seed=2000 + i)
        barrier()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        te2e = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te2e, op=dist.ReduceOp.MAX)
        e2e_s = float(te2e.item())
        h2d = eng.host.nbytes + sum(a.nbytes for a in (tr, dev)) + 4 * (len(y_tr) + len(y_dev))
        hb = torch.tensor([float(h2d)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(hb)
        e2e = {"value": N / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(hb.item()),
               "d2h_bytes_per_step": 32 * world, "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
               "api": "GraphConv.f_train(X, y_train, y_dev, A, train_idx, dev_idx), host SciPy/NumPy inputs"}
    clf.cache_device_inputs = True
    mem_peak = torch.tensor([float(torch.cuda.max_memory_allocated() + (eng.arena.nbytes if eng.arena is not None else 0))],
                            dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(mem_peak, op=dist.ReduceOp.MAX)

    line = None
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_description(wname, cfg), "parallelism": "rows/%d" % world,
                           "exchange": {"slice": "feature-sliced A_hat.H over NVLink peer memory (A_hat replicated)",
                                        "gather": "NCCL all-gather of the dense operand", "none": "single GPU"}[eng.exchange],
                           "l2": "inputs larger than L2 (each N x Hd activation is %.0f MB)" % (N / world * eng.ldh[0] * 4 / 1e6),
                           "spmm_variant": eng.ctx.get_option("spmm_variant"), "gemm_tc": eng.ctx.get_option("gemm_tc"),
                           "power_law_alpha": args.alpha, "nnz_A": int(A.nnz), "nnz_X": int(X.nnz)},
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": e2e,
                "roofline": roof, "split_ms_per_step": split,
                "split_note": "per-op CUDA events from an instrumented pass of %d further steps (rank 0); `sync` = peer "
                              "barriers incl. waiting for the slowest rank, `comm` = slice pushes not fused into a producer, "
                              "`nccl` = all-gather / all-reduce collectives.  Weight-gradient GEMMs run on a second stream "
                              "next to the SpMM / element-wise kernels (side_stream: %s), so the entries overlap in time and "
                              "may sum to more than ms_per_step" % (split_steps, eng.use_side),
                "host_issue_ms_per_step": t_issue / args.steps * 1e3,
                "peak_device_bytes_per_rank": int(mem_peak.item()),
                "exchange_bytes_per_product": eng.conv_exchange_bytes(hd),
                "parity": parity,
                "last_metrics": {"train_loss": metrics[0], "train_acc": metrics[1], "dev_loss": metrics[2],
                                 "dev_acc": metrics[3]},
                "setup_s": {"generate": round(t_gen, 1), "prepare_and_upload": round(t_bind, 1)}}
        if not args.no_cpu_baseline and world == 1:
            n_s = min(args.cpu_sample_nodes, cfg["n"])
            times = cpu_step_seconds(cpu_sample_problem(cfg, n_s), 1, 1)
            cores = blas_threads()
            line["cpu_baseline"] = {
                "value": n_s / float(np.mean(times)), "unit": UNIT, "cores": cores, "kind": "port",
                "sample": "oracle port (SciPy csr_matvecs 1 thread like Theano sd_csr; OpenBLAS %d threads) on %s, 1 warm + 1 "
                          "timed step" % (cores, "the whole workload" if n_s == cfg["n"] else
                                          "a %d-node graph with the workload's per-node statistics" % n_s),
                "same_config": n_s == cfg["n"],
                "host_cpus": len(os.sched_getaffinity(0))}
            if n_s != cfg["n"] and not args.no_cpu_full_ops:
                line["cpu_baseline"]["full_size_ops"] = cpu_full_size_ops(cfg, A)
        else:
            line["cpu_baseline"] = None
        emit(line)
    if world > 1:
        dist.barrier()
    clf.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def ms_estimate_short(cfg):
    """True for workloads whose whole timed region would fit inside one nvidia-smi sample period."""
    return cfg["n"] * cfg["deg"] < 2_000_000


def capi_tag(name):
    from geographconv_b200 import capi
    return capi.TAGS.index(name)


def parity_block(eng, clf, A, cfg, rank, n_rows):
    """In-run parity (every N; collective): one forward pass with the initial weights and a fixed dropout seed, then
    * ``forward_checksum``: checksum of the bit patterns of all N x C probabilities (identical at every GPU count iff the
      row-partitioned forward is bit-identical to the single-GPU forward); ``spmm_checksum``: the same over all of
      S = A_hat.H0 (identical across SpMM engines and exchange designs iff they agree bit for bit);
    * ``spmm``: ``n_rows`` sampled rows of S = A_hat.H0 recomputed in float64 with SciPy from the device's own H0 rows;
    * ``probs``: the same rows of P = softmax(A_hat.(Y.Wout) + bout) recomputed in float64 from the device's own last
      hidden layer Y; argmax compared on every sampled row.
    Errors are relative to the largest reference magnitude (products) / element-wise relative (probabilities)."""
    import scipy.sparse as sp
    n, hd, C = cfg["n"], cfg["hid"][0], cfg["classes"]
    seed = 424242
    eng.forward(train=True, seed=seed)
    checksum = eng.checksum(eng.P, eng.n_loc, C)
    # per-layer checksums of the same forward (H0, then Y / H / T of every highway layer): where a difference between two
    # runs (GPU counts, kernel versions) first appears
    layer_sums = {"H0": "%012x" % eng.checksum(eng.H0, eng.n_loc, hd)}
    for i, b in enumerate(eng.lay):
        for name in ("Y", "H", "T"):
            if b.get(name) is not None:
                layer_sums["%s%d" % (name, i + 1)] = "%012x" % eng.checksum(b[name], eng.n_loc, eng.layout.layers[i]["n_out"])
    rng = np.random.RandomState(12345)
    rows = np.sort(rng.choice(n, size=min(n_rows, n), replace=False))
    sub = A[rows].tocsr()
    nb = np.unique(sub.indices)
    remap = np.full(n, -1, dtype=np.int64)
    remap[nb] = np.arange(len(nb))
    sub64 = sp.csr_matrix((sub.data.astype(np.float64), remap[sub.indices], sub.indptr), shape=(len(rows), len(nb)))
    # (1) one A_hat.H product through the engine's exchange path: S = A_hat.H0
    S = eng.S.view(-1)[: eng.nbuf * eng.ldh[0]].view(eng.nbuf, eng.ldh[0])
    eng._conv(eng.H0, eng.A, S, eng.ldh[0], hd)
    spmm_checksum = eng.checksum(S, eng.n_loc, hd)
    s_gpu = eng.read_rows(S, rows, hd)
    h0_nb = eng.read_rows(eng.H0, nb, hd)
    # (2) final probabilities from the device's own last hidden layer
    p_gpu = eng.read_rows(eng.P, rows, C)
    y_nb = eng.read_rows(eng.x_last, nb, eng.w_last)
    out = None
    if rank == 0:
        s_ref = sub64 @ h0_nb.astype(np.float64)
        spmm_err = float(np.abs(s_gpu - s_ref).max() / max(np.abs(s_ref).max(), 1e-30))
        params = clf.init_params
        Wout, bout = params[-2].astype(np.float64), params[-1].astype(np.float64)
        logits = sub64 @ (y_nb.astype(np.float64) @ Wout) + bout[None, :]
        logits -= logits.max(axis=1, keepdims=True)
        p_ref = np.exp(logits)
        p_ref /= p_ref.sum(axis=1, keepdims=True)
        probs_err = float((np.abs(p_gpu - p_ref) / p_ref).max())
        mism = np.nonzero(p_gpu.argmax(1) != p_ref.argmax(1))[0]
        top2 = np.sort(p_ref[mism], axis=1)[:, -2:] if len(mism) else np.zeros((0, 2))
        out = {"rows_sampled": int(len(rows)), "forward_checksum": "%012x" % checksum,
               "spmm_checksum": "%012x" % spmm_checksum, "layer_checksums": layer_sums,
               "spmm_max_rel": spmm_err, "probs_max_rel": probs_err, "max_rel": max(spmm_err, probs_err),
               "argmax_mismatch": int(len(mism)),
               "argmax_mismatch_top2_rel_gap": [float((b - a) / b) for a, b in top2],
               "tolerance": 1e-3,
               "how": "float64 SciPy on rank 0 from the device's own layer inputs; seed %d, initial weights" % seed}
        assert out["max_rel"] <= 1e-3, "in-run parity failed: %r" % out
    return out


_JSON_FD = None


def emit(line):
    """The one JSON line goes to the real stdout; everything else this process (or NCCL, which prints its version
    banner on stdout) writes lands on stderr, so stdout carries exactly one line."""
    data = (json.dumps(line) + "\n").encode()
    os.write(_JSON_FD if _JSON_FD is not None else 1, data)


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C3", help="C1|C2|C3|C4|tiny (geographconv_b200.synth.CONFIGS)")
    ap.add_argument("--alpha", type=float, default=None, help="power-law degree exponent (BASELINE configs[4])")
    ap.add_argument("--spmm-variant", type=int, default=None)
    ap.add_argument("--gemm-tc", type=int, default=None)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-sample-nodes", type=int, default=32768)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--host-graph", action="store_true", help="build A_hat with the NumPy restatement instead of the GPU builder")
    ap.add_argument("--no-cpu-full-ops", action="store_true", help="skip the full-size host SpMM / GEMM timing")
    ap.add_argument("--parity-rows", type=int, default=2000, help="rows of the in-run parity block (0 = skip)")
    ap.add_argument("--ncu-traffic-bytes", type=float, default=None,
                    help="dram bytes per launch of the dominant kernel from the committed ncu capture")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    from geographconv_b200 import synth
    cfg = dict(synth.CONFIGS[args.workload])
    if args.impl == "reference":
        run_reference(args, cfg, args.workload)
    else:
        run_gpu(args, cfg, args.workload)


if __name__ == "__main__":
    main()
