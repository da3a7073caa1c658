#!/bin/bash
mkdir -p gpurun_out
export INRF_TC_WATCHDOG_CYCLES=400000000
timeout 300 python tests/tools/tc_perf.py 160000 > gpurun_out/tc_perf.log 2>&1
timeout 300 python tests/tools/tc_perf.py 160000 ssr >> gpurun_out/tc_perf.log 2>&1
timeout 300 python tests/tools/fused_perf.py > gpurun_out/fused_perf.log 2>&1
unset INRF_TC_WATCHDOG_CYCLES
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
grep TC_PERF gpurun_out/tc_perf.log; tail -2 gpurun_out/fused_perf.log; tail -8 gpurun_out/pytest_gpu.log | cut -c1-300
