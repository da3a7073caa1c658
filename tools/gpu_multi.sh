#!/bin/bash
# multi-GPU check: smoke() on cuda:0, then the driver's torchrun launch of bench.py on N GPUs
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
cat gpurun_out/bench_${N}gpu.json | cut -c1-600; tail -3 gpurun_out/bench_${N}gpu.err
