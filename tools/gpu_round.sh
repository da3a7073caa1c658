#!/bin/bash
# One GPU session: bring-up diagnostics, GPU test-suite, bench.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
export INRF_TC_CHECK=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/gpu.txt
timeout 600 python tools/tc_debug.py > gpurun_out/tc_debug.log 2>&1
echo "tc_debug exit $?" >> gpurun_out/tc_debug.log
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_render.py > gpurun_out/pytest_stages.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_render.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_render.log 2>&1
unset INRF_TC_CHECK
if grep -q "TC_DEBUG PASS" gpurun_out/tc_debug.log; then
  timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
fi
bash tools/gpu_profile.sh > gpurun_out/profile.log 2>&1
tail -n 4 gpurun_out/tc_debug.log; tail -n 15 gpurun_out/pytest_stages.log; tail -n 30 gpurun_out/pytest_render.log; cat gpurun_out/bench.json 2>/dev/null; tail -5 gpurun_out/bench.err 2>/dev/null
