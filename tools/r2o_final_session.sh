#!/bin/bash
# Round-2 final single-GPU validation of the committed defaults: GPU test suite, smoke, one C3 bench line.
set -u
mkdir -p gpurun_out
T=${1:-r2o}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${T}_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_c3_n1.json 2> gpurun_out/${T}_bench_c3_n1.err
echo "bench rc=$?"; python - <<PY
import json
d = json.load(open("gpurun_out/${T}_bench_c3_n1.json"))
print("ms/step %.2f" % d["ms_per_step"], "e2e %.2f" % d["e2e"]["ms_per_step"], d["split_ms_per_step"], d["parity"]["forward_checksum"], d["parity"]["max_rel"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"])
PY
