#!/bin/bash
# fastest kernel check: tc_debug (vs fp32 kernel, watchdog on) + kernel-only timing
mkdir -p gpurun_out
INRF_TC_CHECK=1 timeout 300 python tests/tools/tc_debug.py > gpurun_out/tc_debug.log 2>&1
grep -E "TC_DEBUG|rror" gpurun_out/tc_debug.log | cut -c1-160
timeout 300 python tests/tools/tc_perf.py 160000 2>&1 | grep -E "TC_PERF|rror" | cut -c1-200
