#!/bin/bash
# quick kernel check: correctness vs the fp32 kernel + kernel-only timing
mkdir -p gpurun_out
: > gpurun_out/quick.log
INRF_TC_CHECK=1 timeout 300 python tests/tools/tc_debug.py > gpurun_out/tc_debug.log 2>&1
grep -E "^\[|TC_DEBUG" gpurun_out/tc_debug.log | cut -c1-160
timeout 300 python tests/tools/tc_perf.py 160000 >> gpurun_out/quick.log 2>&1
INRF_TC_NOWEIGHTS=1 timeout 300 python tests/tools/tc_perf.py 160000 >> gpurun_out/quick.log 2>&1
timeout 300 python tests/tools/tc_perf.py 160000 ssr >> gpurun_out/quick.log 2>&1
grep -E "TC_PERF|rror" gpurun_out/quick.log | cut -c1-200
