#!/bin/bash
# Round-2 session N (2 GPUs): per-layer checksums of the parity forward at 1 and at 2 GPUs, gemm_v 2.
set -u
mkdir -p gpurun_out
T=${1:-r2n}
show() { python -c "import json; d=json.load(open('gpurun_out/${T}_$1.json')); print('$1', d['parity']['forward_checksum'], d['parity']['spmm_checksum'], d['parity']['layer_checksums'])"; }
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus 2 --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline > gpurun_out/${T}_n2.json 2> gpurun_out/${T}_n2.err &
CUDA_VISIBLE_DEVICES=0 true
wait
show n2
CUDA_VISIBLE_DEVICES=0 timeout 400 python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline > gpurun_out/${T}_n1.json 2> gpurun_out/${T}_n1.err
show n1
