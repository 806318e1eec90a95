#!/bin/bash
# Round-2 session L: does a row's result depend on how many rows the call has?  (forward checksum differs 1 vs N GPUs)
set -u
mkdir -p gpurun_out
T=${1:-r2l}
timeout 300 python tools/gemm_rows_check.py 2 > gpurun_out/${T}_rows_check_v2.txt 2>&1; echo "rows v2 rc=$?"; cat gpurun_out/${T}_rows_check_v2.txt
timeout 300 python tools/gemm_rows_check.py 1 > gpurun_out/${T}_rows_check_v1.txt 2>&1; echo "rows v1 rc=$?"; cat gpurun_out/${T}_rows_check_v1.txt
GCNB_GEMM_V=1 timeout 300 python bench.py --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/${T}_bench_v1.json 2> gpurun_out/${T}_bench_v1.err
python -c "import json; d=json.load(open('gpurun_out/${T}_bench_v1.json')); print('v1', d['parity']['forward_checksum'], d['parity']['spmm_checksum'])"
