#!/bin/bash
# Round-2 session M (2 GPUs): which part makes the forward checksum differ between 1 and N GPUs with gemm_v 2?
set -u
mkdir -p gpurun_out
T=${1:-r2m}
N=$(python -c "import torch; print(torch.cuda.device_count())")
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
      bench.py --gpus $N --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/${T}_$name.json 2> gpurun_out/${T}_$name.err
  python -c "import json; d=json.load(open('gpurun_out/${T}_$name.json')); print('$name', d['parity']['forward_checksum'], d['parity']['spmm_checksum'], d['config'].get('exchange'))"
}
run v1_slice GCNB_GEMM_V=1
run v2_gather GCNB_GEMM_V=2 GCNB_EXCHANGE=gather
