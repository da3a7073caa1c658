#!/bin/bash
# A/B of the epilogue drain schedules (INRF_TC_EPI = 0 default, 1 chunk per warp group, 2 the same with ordered loads)
mkdir -p gpurun_out
: > gpurun_out/epi.log
export INRF_TC_WATCHDOG_CYCLES=400000000
for m in 0 1 2 0 1 2; do
  echo "EPI=$m" >> gpurun_out/epi.log
  INRF_TC_EPI=$m timeout 200 python tests/tools/tc_perf.py 160000 >> gpurun_out/epi.log 2>&1
done
for m in 1 2; do
  echo "EPI=$m ssr" >> gpurun_out/epi.log
  INRF_TC_EPI=$m timeout 200 python tests/tools/tc_perf.py 160000 ssr >> gpurun_out/epi.log 2>&1
  INRF_TC_EPI=$m timeout 300 python -m pytest tests/test_gpu_fused.py tests/test_gpu_render.py -q -p no:cacheprovider -x 2>&1 | tail -2 >> gpurun_out/epi.log
done
grep -E "^EPI|TC_PERF|passed|failed|rror" gpurun_out/epi.log | cut -c1-220
