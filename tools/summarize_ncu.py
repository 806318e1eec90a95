#!/usr/bin/env python
"""Turn ncu CSV exports into the markdown summaries committed under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches_r1.csv  > profiles/r1_launches.md
    python tools/summarize_ncu.py full     gpurun_out/prof_r1_raw.csv  > profiles/r1_ncu_full.md
    python tools/summarize_ncu.py traffic  profiles/r2_ncu_full_raw.csv C3 spmm_panel_kernel 0 - 2,4,14,19 > profiles/roofline_traffic.json

``traffic`` regenerates the file bench.py reads ``roofline.traffic`` from: DRAM bytes per launch (dram__bytes_read.sum +
dram__bytes_write.sum) and the L2 hit rate of the dominant kernel, averaged over the launches of that kernel whose grid
has at least the given number of CTAs in x (the full-width A_hat.H products; narrower launches are other products).
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


def traffic(path, workload, kernel, min_grid_x="0", grid_y=None, ids=None):
    """ids: comma-separated launch numbers (rows of the CSV, 0-based) when the grid alone does not single out the product"""
    import json
    rows = list(csv.reader(open(path)))
    h, units = rows[0], rows[1]
    col = {name: h.index(name) for name in ("Kernel Name", "Grid Size", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                            "lts__t_sector_hit_rate.pct", "gpu__time_duration.sum")}
    scale = lambda i: {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[i], 1.0)
    picked = []
    for n, r in enumerate(rows[2:]):
        if kernel not in r[col["Kernel Name"]]:
            continue
        grid = [int(x) for x in r[col["Grid Size"]].replace("(", "").replace(")", "").split(",")]
        if grid[0] < int(min_grid_x) or (grid_y not in (None, "-") and grid[1] != int(grid_y)):
            continue
        if ids and n not in [int(x) for x in ids.split(",")]:
            continue
        rd = _num(r[col["dram__bytes_read.sum"]]) * scale(col["dram__bytes_read.sum"])
        wr = _num(r[col["dram__bytes_write.sum"]]) * scale(col["dram__bytes_write.sum"])
        picked.append((n, r[col["Kernel Name"]].split("(")[0], grid, rd, wr, _num(r[col["lts__t_sector_hit_rate.pct"]])))
    if not picked:
        raise SystemExit("no launch of %s with grid.x >= %s in %s" % (kernel, min_grid_x, path))
    rd = sum(p[3] for p in picked) / len(picked)
    wr = sum(p[4] for p in picked) / len(picked)
    out = {"workload": workload, "kernel": picked[0][1].replace("void ", "").replace("(anonymous namespace)::", ""),
           "dram_bytes_per_launch": int(rd + wr), "dram_bytes_read": int(rd), "dram_bytes_write": int(wr),
           "l2_hit_pct": round(sum(p[5] for p in picked) / len(picked), 2),
           "launches": [p[0] for p in picked], "grid": picked[0][2],
           "source": "%s launches %s: dram__bytes_read.sum + dram__bytes_write.sum, mean over %d launches" % (
               path, "/".join("#%d" % p[0] for p in picked), len(picked))}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
