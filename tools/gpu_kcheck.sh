#!/bin/bash
# kernel change check: tc_debug (vs fp32 kernel, watchdog on), render + train tests, kernel-only timing (object + ssr)
mkdir -p gpurun_out
INRF_TC_CHECK=1 timeout 300 python tests/tools/tc_debug.py > gpurun_out/tc_debug.log 2>&1
grep -E "^\[|TC_DEBUG|rror" gpurun_out/tc_debug.log | cut -c1-160
INRF_TC_CHECK=1 timeout 900 python -m pytest tests/test_gpu_render.py tests/test_gpu_train_tc.py tests/test_gpu_stages.py -m gpu -q -p no:cacheprovider -x > gpurun_out/pytest_k.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_k.log; grep -E "^E  |FAILED" gpurun_out/pytest_k.log | cut -c1-300 | head -10
: > gpurun_out/kperf.log
timeout 300 python tests/tools/tc_perf.py 160000 >> gpurun_out/kperf.log 2>&1
timeout 300 python tests/tools/tc_perf.py 160000 ssr >> gpurun_out/kperf.log 2>&1
grep -E "TC_PERF|rror" gpurun_out/kperf.log | cut -c1-200
