#!/bin/bash
# NCCL check of the sharded frame gather and the data-parallel training step on N GPUs
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 \
  tests/tools/multi_gpu_check.py > gpurun_out/multi_check_${N}gpu.log 2>&1
grep -E "MULTI|rror" gpurun_out/multi_check_${N}gpu.log | head; tail -3 gpurun_out/multi_check_${N}gpu.log | cut -c1-300
