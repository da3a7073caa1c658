#!/bin/bash
# 2-GPU session: world-size-2 bench (strong scaling job with the all-gather inside) + NCCL log + multi-GPU check
mkdir -p gpurun_out
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --views 8 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/tools/multi_gpu_check.py > gpurun_out/multi_check_2gpu.log 2>&1
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --views 8 --no-cpu-baseline > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
tail -1 gpurun_out/bench_2gpu.json | cut -c1-3000; grep -c "AllGather" gpurun_out/bench_2gpu.err; grep -m3 "AllGather" gpurun_out/bench_2gpu.err | cut -c1-300; tail -3 gpurun_out/bench_2gpu.err | cut -c1-300; grep MULTI gpurun_out/multi_check_2gpu.log; tail -1 gpurun_out/bench_1gpu.json | cut -c1-400
