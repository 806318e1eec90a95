#!/usr/bin/env python
"""Turn ncu CSV exports into the markdown summaries committed under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches_r1.csv  > profiles/r1_launches.md
    python tools/summarize_ncu.py full     gpurun_out/prof_r1_raw.csv  > profiles/r1_ncu_full.md
    python tools/summarize_ncu.py traffic  profiles/r2c_ncu_full_raw.csv C3 spmm_panel_kernel 3+4,6+7,18+19,24+25 > profiles/roofline_traffic.json

``traffic`` regenerates the file bench.py reads ``roofline.traffic`` from: DRAM bytes per launch (dram__bytes_read.sum +
dram__bytes_write.sum) and the L2 hit rate of the dominant kernel, per A_hat.H product at the hidden width (a product is
two launches of the panel kernel: the 32-column panels and the remainder columns).
"""
import collections
import csv
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    kn, mv, gs, bs = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Block Size")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        name = r[kn].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        t = float(r[mv].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | total ms | share | avg ms |")
    print("|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.3f | %.1f%% | %.3f |" % (k[:80], v[0], v[1] / 1e6, 100 * v[1] / tot, v[1] / v[0] / 1e6))
    print("\ntotal device time in the list: %.1f ms over %d launches" % (tot / 1e6, sum(v[0] for v in agg.values())))


def full(path):
    rows = list(csv.reader(open(path)))
    h, units = rows[0], rows[1]
    want = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("gpu__time_duration.sum", "ms"),
            ("dram__bytes_read.sum", "dram rd GB"), ("dram__bytes_write.sum", "dram wr GB"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
            ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
            ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
            ("launch__registers_per_thread", "regs"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %")]
    idx = [(h.index(a), b) for a, b in want if a in h]
    print("| # | " + " | ".join(b for _, b in idx) + " |")
    print("|---|" + "---|" * len(idx))
    for n, r in enumerate(rows[2:]):
        cells = []
        for i, b in idx:
            v = r[i]
            if b == "kernel":
                v = "`" + v.split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:40] + "`"
            elif b == "ms":
                f = float(v.replace(",", ""))
                u = units[i]
                v = "%.3f" % (f / 1e6 if u in ("ns", "nsecond") else f / 1e3 if u in ("us", "usecond") else f)
            elif b.endswith("GB"):
                f = float(v.replace(",", ""))
                u = units[i]
                scale = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}.get(u, 1e-9)
                v = "%.3f" % (f * scale)
            else:
                try:
                    v = "%.1f" % float(v.replace(",", ""))
                except ValueError:
                    pass
            cells.append(v)
        print("| %d | " % n + " | ".join(cells) + " |")


def _num(v):
    return float(v.replace(",", ""))


def traffic(path, workload, kernel, groups):
    """groups: comma-separated products, each the launch numbers (rows of the CSV, 0-based) that make up ONE product joined
    with '+', e.g. 3+4,6+7,18+19,24+25 (the 32-column panels and the remainder-column launch of each A_hat.H product).
    DRAM bytes are summed inside a product and averaged over the products; the L2 hit rate is weighted by duration."""
    import json
    rows = list(csv.reader(open(path)))
    h, units = rows[0], rows[1]
    col = {name: h.index(name) for name in ("Kernel Name", "Grid Size", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                            "lts__t_sector_hit_rate.pct", "gpu__time_duration.sum")}
    scale = lambda i: {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[i], 1.0)
    data = rows[2:]
    prods = []
    names = []
    for g in groups.split(","):
        rd = wr = dur = hit = 0.0
        for n in (int(x) for x in g.split("+")):
            r = data[n]
            if kernel not in r[col["Kernel Name"]]:
                raise SystemExit("launch %d is %s, not %s" % (n, r[col["Kernel Name"]][:60], kernel))
            names.append(r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "") + " grid " +
                         r[col["Grid Size"]])
            t = _num(r[col["gpu__time_duration.sum"]])
            rd += _num(r[col["dram__bytes_read.sum"]]) * scale(col["dram__bytes_read.sum"])
            wr += _num(r[col["dram__bytes_write.sum"]]) * scale(col["dram__bytes_write.sum"])
            hit += _num(r[col["lts__t_sector_hit_rate.pct"]]) * t
            dur += t
        prods.append((rd, wr, hit / dur, dur))
    k = len(prods)
    out = {"workload": workload, "kernel": kernel,
           "dram_bytes_per_launch": int(sum(p[0] + p[1] for p in prods) / k),
           "dram_bytes_read": int(sum(p[0] for p in prods) / k), "dram_bytes_write": int(sum(p[1] for p in prods) / k),
           "l2_hit_pct": round(sum(p[2] for p in prods) / k, 2),
           "launches_of_one_product": sorted(set(names)),
           "source": "%s, products %s: dram__bytes_read.sum + dram__bytes_write.sum summed inside a product (all its "
                     "launches), mean over %d products" % (path, groups, k)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
