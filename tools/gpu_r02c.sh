#!/bin/bash
# round 2 session c: fused kernel tests first (wrapped in a short timeout: a new kernel), then the full suite, bench, reference arm
mkdir -p gpurun_out
export INRF_TC_WATCHDOG_CYCLES=400000000
timeout 600 python -m pytest tests/test_gpu_fused.py -x -q -p no:cacheprovider > gpurun_out/pytest_fused.log 2>&1
echo "fused rc=$?" >> gpurun_out/pytest_fused.log
unset INRF_TC_WATCHDOG_CYCLES
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_fused.py > gpurun_out/pytest_gpu.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 900 python bench.py --steps 3 --warmup 3 --views 8 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -n 30 gpurun_out/pytest_fused.log; tail -n 30 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
