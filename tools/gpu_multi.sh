#!/bin/bash
# N-GPU session (gpurun --gpus N): the driver's torchrun launch of bench.py (strong-scaling job with the NCCL all-gather
# inside the timed region; NCCL's own log goes to stderr), then the NCCL bit-exactness checks of the sharded paths.
N=${1:-2}
mkdir -p gpurun_out
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL NCCL_DEBUG_FILE=gpurun_out/nccl_${N}gpu.%h.%p.log timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
  --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 \
  tests/tools/multi_gpu_check.py > gpurun_out/multi_check_${N}gpu.log 2>&1
tail -n 1 gpurun_out/bench_${N}gpu.json | cut -c1-1500; tail -3 gpurun_out/bench_${N}gpu.err | cut -c1-300
cat gpurun_out/nccl_${N}gpu.*.log 2>/dev/null | grep -c "AllGather"; grep MULTI gpurun_out/multi_check_${N}gpu.log
