#!/bin/bash
# Round-2 session I: end-to-end leg with deferred, grouped uploads (GCNB_X_GROUPS sweep); C1 line.
set -u
mkdir -p gpurun_out
T=${1:-r2i}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_c3.json 2> gpurun_out/${T}_bench_c3.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/${T}_bench_c3.json; tail -2 gpurun_out/${T}_bench_c3.err
for g in 1 4 16; do
  GCNB_X_GROUPS=$g timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --parity-rows 0 > gpurun_out/${T}_bench_c3_xg$g.json 2> gpurun_out/${T}_bench_c3_xg$g.err
  echo "bench xg$g rc=$?"
done
timeout 300 python bench.py --workload C1 --steps 20 --warmup 3 > gpurun_out/${T}_bench_c1.json 2> gpurun_out/${T}_bench_c1.err; echo "bench c1 rc=$?"; cut -c1-250 gpurun_out/${T}_bench_c1.json
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${T}_bench_c*.json")):
    try:
        d = json.load(open(f))
        print(f, "ms/step %.2f" % d["ms_per_step"], "gemm %.2f" % d["split_ms_per_step"]["gemm"], "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"], 2), d["e2e"] and d["e2e"]["h2d_bytes_per_step"])
    except Exception as e:
        print(f, "unreadable", e)
PY
