#!/bin/bash
# Round-2 single-GPU measurement session (run under gpurun on one B200):
#   tests, A_hat builder timings, C1 / C3 / C5 bench lines, ncu launch list and --set full capture.
# Numbers printed by runs under ncu are never bench values.
set -u
mkdir -p gpurun_out
T=${1:-r2}
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
for a in none 2.0 1.5; do
  if [ $a = none ]; then python tools/adj_bench.py; else python tools/adj_bench.py --alpha $a; fi
done > gpurun_out/${T}_adjacency_build.txt 2>&1; cat gpurun_out/${T}_adjacency_build.txt
python bench.py --workload C1 --steps 20 --warmup 3 > gpurun_out/${T}_bench_c1_n1.json 2> gpurun_out/${T}_bench_c1_n1.err; cut -c1-250 gpurun_out/${T}_bench_c1_n1.json
python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_c3_n1.json 2> gpurun_out/${T}_bench_c3_n1.err; cut -c1-250 gpurun_out/${T}_bench_c3_n1.json
for a in 1.5 2.0 2.5 3.0; do
  python bench.py --alpha $a --steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/${T}_bench_c5_alpha$a.json 2> gpurun_out/${T}_bench_c5_alpha$a.err
  cut -c1-200 gpurun_out/${T}_bench_c5_alpha$a.json
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline --parity-rows 0 > gpurun_out/${T}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:"spmm_panel_kernel|spmm_fixup_kernel|row_softmax_kernel|gemm_tc_kernel|wgrad_tc_kernel|highway_bwd_colsum_kernel|act_bwd_colsum_kernel|xent_grad_dense_kernel" \
    -c 27 -o /tmp/prof_${T} python bench.py --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline --parity-rows 0 > gpurun_out/${T}_ncu2.log 2>&1
ncu -i /tmp/prof_${T}.ncu-rep --page raw --csv > gpurun_out/${T}_ncu_full_raw.csv 2>> gpurun_out/${T}_ncu2.log
ncu --set full --clock-control none -k regex:"spmm_panel_kernel|spmm_fixup_kernel" -c 8 -o /tmp/prof_${T}_a15 \
    python bench.py --alpha 1.5 --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline --parity-rows 0 > gpurun_out/${T}_ncu3.log 2>&1
ncu -i /tmp/prof_${T}_a15.ncu-rep --page raw --csv > gpurun_out/${T}_ncu_alpha1.5_raw.csv 2>> gpurun_out/${T}_ncu3.log
ls -la gpurun_out | tail -20
