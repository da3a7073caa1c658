#!/bin/bash
# One GPU session: bring-up diagnostics, GPU test-suite, bench, ncu evidence.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
export INRF_TC_CHECK=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/gpu.txt
timeout 600 python tests/tools/tc_debug.py > gpurun_out/tc_debug.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
unset INRF_TC_CHECK
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
[ "$1" == "profile" ] && bash tools/gpu_profile.sh > gpurun_out/profile.log 2>&1
tail -n 2 gpurun_out/tc_debug.log; tail -n 25 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
