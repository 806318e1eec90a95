#!/bin/bash
# Round-2 session J (final code): ncu launch list + --set full capture of one C3 step on one B200, plus the live bench
# line the shares are checked against.  Numbers printed by runs under ncu are never bench values.
set -u
mkdir -p gpurun_out
T=${1:-r2j}
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_c3_n1.json 2> gpurun_out/${T}_bench_c3_n1.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/${T}_bench_c3_n1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline --parity-rows 0 > gpurun_out/${T}_ncu1.log 2>&1
echo "ncu launches rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:"spmm_panel_kernel|spmm_fixup_kernel|row_softmax_kernel|gemm_tc2_kernel|gemm_tc_kernel|wgrad_tc_kernel|highway_bwd_colsum_kernel|act_bwd_colsum_kernel|xent_grad_dense_kernel" \
    -c 27 -o /tmp/prof_${T} python bench.py --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline --parity-rows 0 > gpurun_out/${T}_ncu2.log 2>&1
echo "ncu full rc=$?"
ncu -i /tmp/prof_${T}.ncu-rep --page raw --csv > gpurun_out/${T}_ncu_full_raw.csv 2>> gpurun_out/${T}_ncu2.log
ls -la gpurun_out | tail -12
