#!/bin/bash
# Round-2 session K: N-GPU check of the final code (bit-equality tool + one bench line), N = number of visible GPUs.
set -u
mkdir -p gpurun_out
N=$(python -c "import torch; print(torch.cuda.device_count())")
T=${1:-r2k}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    tools/multigpu_check.py > gpurun_out/${T}_multigpu_check_n$N.log 2>&1
echo "multigpu_check rc=$?"; grep -a "MULTIGPU_CHECK\|bit-identical\|PASS\|FAIL" gpurun_out/${T}_multigpu_check_n$N.log | tail -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_c3_n$N.json 2> gpurun_out/${T}_bench_c3_n$N.err
echo "bench rc=$?"; cut -c1-400 gpurun_out/${T}_bench_c3_n$N.json; tail -3 gpurun_out/${T}_bench_c3_n$N.err
