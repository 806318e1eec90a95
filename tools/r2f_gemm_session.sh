#!/bin/bash
# Round-2 session F: the two-CTAs-per-SM tcgen05 GEMM (gemm_v 2) against the one-tile-per-SM kernel (gemm_v 1).
# Every step runs under its own timeout so a hung kernel cannot hold the box.
set -u
mkdir -p gpurun_out
T=${1:-r2f}
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm or highway or wgrad" > gpurun_out/${T}_pytest_gemm.log 2>&1
echo "pytest gemm rc=$?"; tail -5 gpurun_out/${T}_pytest_gemm.log
timeout 300 python tools/gemm_bench.py > gpurun_out/${T}_gemm_bench.txt 2>&1; echo "gemm_bench rc=$?"; cat gpurun_out/${T}_gemm_bench.txt
timeout 300 python tools/gemm_phases.py 2 > gpurun_out/${T}_gemm_phases_v2.txt 2>&1; echo "phases rc=$?"; cat gpurun_out/${T}_gemm_phases_v2.txt
for v in 2 1; do
  GCNB_GEMM_V=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_c3_v$v.json 2> gpurun_out/${T}_bench_c3_v$v.err
  echo "bench v$v rc=$?"; cut -c1-400 gpurun_out/${T}_bench_c3_v$v.json; tail -2 gpurun_out/${T}_bench_c3_v$v.err
done
for h in 1536; do
  GCNB_HOT_MAX=$h GCNB_HOT_DENSITY=0.02 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 --parity-rows 0 > gpurun_out/${T}_bench_c3_hot$h.json 2> gpurun_out/${T}_bench_c3_hot$h.err
  echo "bench hot$h rc=$?"; cut -c1-300 gpurun_out/${T}_bench_c3_hot$h.json
done
