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
               on all cores) on a bounded sample: a smaller graph with the same per-node statistics.

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
    n_s = args.cpu_sample_nodes
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
                       "sample_nodes": n_s},
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
    A, X, Y, tr, dev, te, _ = synth.synthetic_problem(cfg, seed=77, alpha=args.alpha, row_range=rr)
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
    y_tr, y_dev = Y[tr], Y[dev]
    t_bind = time.time()
    eng.bind(X, A, need_backward=True)
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

    # ---- device-resident leg (value) ----
    for i in range(args.warmup):
        step_resident(i)
    eng.read_metrics()
    barrier()
    eng.ctx.prof_enable(True)
    eng.ctx.prof_reset()
    launches0 = eng.ctx.launch_count()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(eng.stream)  # the stream every kernel of the step is launched on
    for i in range(args.steps):
        step_resident(args.warmup + i)
    e1.record(eng.stream)
    barrier()
    clocks = sampler.stop() if sampler else None
    ms = e0.elapsed_time(e1)
    launches = eng.ctx.launch_count() - launches0
    prof = eng.ctx.prof_collect()
    eng.ctx.prof_enable(False)
    metrics = eng.read_metrics()
    tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    ms_per_step = ms / args.steps
    value = N / (ms_per_step * 1e-3)

    # roofline of the dominant kernel: A_hat.H SpMM at the hidden width (tag spmm_a)
    hd = cfg["hid"][0]
    spmm_ms, spmm_ops = prof["spmm_a"]
    peak, peak_kind = load_peaks()
    b_touch = eng.conv_touched_bytes(hd)
    pieces = len(eng._panels(hd, eng.ldh[0], False)) if world > 1 else 1  # SpMM launches per graph convolution
    roof = None
    traffic = args.ncu_traffic_bytes
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    kname = ["spmm_ldg_kernel", "spmm_bulk_kernel", "spmm_panel_kernel"][eng.A.engine_for(eng, eng.ldh[0], hd)]
    if traffic is None and world == 1 and os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("workload") == wname and args.alpha is None and tj.get("kernel", "").startswith(kname):
            traffic = tj["dram_bytes_per_launch"]  # from the committed ncu --set full capture of this kernel
    if spmm_ops:
        t_launch = spmm_ms / (spmm_ops / pieces) * 1e-3  # all pieces of one product
        ach = b_touch / t_launch / 1e9
        tms2 = torch.tensor([ach], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tms2, op=dist.ReduceOp.MIN)  # slowest rank's kernel
        ach = float(tms2.item())
        roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic,
                "kernel": "%s (A_hat.H, K=%d)" % (kname, hd),
                "algorithmic_bytes_per_launch": b_touch, "launches_timed": spmm_ops, "launches_per_product": pieces,
                "avg_launch_ms": spmm_ms / spmm_ops * pieces, "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)",
                "per_rank": world > 1}
        if kname == "spmm_panel_kernel":
            roof["note"] = ("L2-resident column panels: B leaves HBM once per product, the per-nonzero gathers are L2 hits, so "
                            "algorithmic B_touch / t may exceed the HBM peak; `traffic` is the DRAM bytes ncu counted")
    split = {k: round(v[0] / args.steps, 4) for k, v in prof.items()}

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

    line = None
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_description(wname, cfg), "parallelism": "rows/%d" % world,
                           "l2": "inputs larger than L2 (each N x Hd activation is %.0f MB)" % (N / world * 320 * 4 / 1e6),
                           "spmm_variant": eng.ctx.get_option("spmm_variant"), "gemm_tc": eng.ctx.get_option("gemm_tc"),
                           "power_law_alpha": args.alpha, "nnz_A": int(A.nnz), "nnz_X": int(X.nnz)},
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": e2e,
                "roofline": roof, "split_ms_per_step": split,
                "last_metrics": {"train_loss": metrics[0], "train_acc": metrics[1], "dev_loss": metrics[2],
                                 "dev_acc": metrics[3]},
                "setup_s": {"generate": round(t_gen, 1), "prepare_and_upload": round(t_bind, 1)}}
        if not args.no_cpu_baseline and world == 1:
            n_s = args.cpu_sample_nodes
            times = cpu_step_seconds(cpu_sample_problem(cfg, n_s), 1, 1)
            cores = blas_threads()
            line["cpu_baseline"] = {
                "value": n_s / float(np.mean(times)), "unit": UNIT, "cores": cores, "kind": "port",
                "sample": "oracle port (SciPy csr_matvecs 1 thread like Theano sd_csr; OpenBLAS %d threads) on a "
                          "%d-node graph with the workload's per-node statistics, 1 warm + 1 timed step" % (cores, n_s),
                "host_cpus": len(os.sched_getaffinity(0))}
        else:
            line["cpu_baseline"] = None
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


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
