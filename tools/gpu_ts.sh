#!/bin/bash
# A/B: activations in tensor memory (INRF_TC_TS=1/2) vs shared memory (default); ring depth in the TS variant
mkdir -p gpurun_out
: > gpurun_out/ts.log
export INRF_TC_WATCHDOG_CYCLES=400000000
for cfg in "0 0" "2 6" "1 6" "2 6" "0 0"; do
  set -- $cfg
  echo "TS=$1 NS=$2" >> gpurun_out/ts.log
  INRF_TC_TS=$1 INRF_TC_NS=$2 timeout 200 python tests/tools/tc_perf.py 160000 >> gpurun_out/ts.log 2>&1
done
echo "TS=2 tests" >> gpurun_out/ts.log
INRF_TC_TS=2 timeout 400 python -m pytest tests/test_gpu_fused.py tests/test_gpu_render.py tests/test_gpu_wide.py tests/test_gpu_status.py -q -p no:cacheprovider 2>&1 | tail -6 >> gpurun_out/ts.log
INRF_TC_TS=2 timeout 200 python tests/tools/fused_perf.py >> gpurun_out/ts.log 2>&1
grep -E "^TS|TC_PERF|FUSED_PERF|passed|failed|rror" gpurun_out/ts.log | cut -c1-330
