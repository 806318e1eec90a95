#!/bin/bash
# Round-2 session H: gemm_v 2 with the weights' residual tile derived in shared memory (gemm_blo2), copy-only copy stream.
set -u
mkdir -p gpurun_out
T=${1:-r2h}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/${T}_pytest.log
timeout 300 python tools/gemm_bench.py > gpurun_out/${T}_gemm_bench.txt 2>&1; echo "gemm_bench rc=$?"; cat gpurun_out/${T}_gemm_bench.txt
GCNB_GEMM_BLO2=0 timeout 300 python tools/gemm_bench.py > gpurun_out/${T}_gemm_bench_blo0.txt 2>&1; echo "gemm_bench blo2=0 rc=$?"; cat gpurun_out/${T}_gemm_bench_blo0.txt
timeout 300 python tools/gemm_phases.py 2 > gpurun_out/${T}_gemm_phases_v2.txt 2>&1; echo "phases rc=$?"; cat gpurun_out/${T}_gemm_phases_v2.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_c3.json 2> gpurun_out/${T}_bench_c3.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/${T}_bench_c3.json; tail -2 gpurun_out/${T}_bench_c3.err
GCNB_X_GROUPS=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --parity-rows 0 > gpurun_out/${T}_bench_c3_xg1.json 2> gpurun_out/${T}_bench_c3_xg1.err
echo "bench xg1 rc=$?"; cut -c1-300 gpurun_out/${T}_bench_c3_xg1.json
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${T}_bench_c3*.json")):
    try:
        d = json.load(open(f))
        print(f, "ms/step %.2f" % d["ms_per_step"], "gemm %.2f" % d["split_ms_per_step"]["gemm"], "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"], 2))
    except Exception as e:
        print(f, "unreadable", e)
PY
