#!/bin/bash
# CTA-pair kernel: correctness vs the fp32 kernel, kernel-only timing (pair vs single-CTA kernel), wait-cycle profile
mkdir -p gpurun_out
: > gpurun_out/pair.log
INRF_TC_CHECK=1 timeout 300 python tests/tools/tc_debug.py > gpurun_out/tc_debug.log 2>&1
grep -E "^\[|TC_DEBUG" gpurun_out/tc_debug.log | cut -c1-200
timeout 300 python tests/tools/tc_perf.py 160000 >> gpurun_out/pair.log 2>&1
INRF_TC_PAIR=0 timeout 300 python tests/tools/tc_perf.py 160000 >> gpurun_out/pair.log 2>&1
timeout 300 python tests/tools/tc_perf.py 160000 ssr >> gpurun_out/pair.log 2>&1
grep -E "TC_PERF|rror" gpurun_out/pair.log
if [ -f intrinsicnerf_b200/csrc/libinrf_prof.so ]; then
  INRF_LIB=$PWD/intrinsicnerf_b200/csrc/libinrf_prof.so timeout 300 python tests/tools/tc_perf.py 160000 > gpurun_out/pairprof.log 2>&1
  grep TC2PROF gpurun_out/pairprof.log | tail -${PROF_LINES:-31}
fi
