#!/bin/bash
# CTA-pair kernel (INRF_TC_PAIR=1): correctness vs the fp32 kernel, kernel-only timing vs the single-CTA kernel
mkdir -p gpurun_out
: > gpurun_out/pair.log
INRF_TC_PAIR=1 INRF_TC_CHECK=1 timeout 300 python tests/tools/tc_debug.py > gpurun_out/tc_debug.log 2>&1
grep -E "^\[|TC_DEBUG" gpurun_out/tc_debug.log | cut -c1-160
for rep in 1 2; do
  echo "pair" >> gpurun_out/pair.log
  INRF_TC_PAIR=1 timeout 300 python tests/tools/tc_perf.py 160000 >> gpurun_out/pair.log 2>&1
  echo "single" >> gpurun_out/pair.log
  timeout 300 python tests/tools/tc_perf.py 160000 >> gpurun_out/pair.log 2>&1
done
echo "pair noweights" >> gpurun_out/pair.log
INRF_TC_PAIR=1 INRF_TC_NOWEIGHTS=1 timeout 300 python tests/tools/tc_perf.py 160000 >> gpurun_out/pair.log 2>&1
echo "pair ssr" >> gpurun_out/pair.log
INRF_TC_PAIR=1 timeout 300 python tests/tools/tc_perf.py 160000 ssr >> gpurun_out/pair.log 2>&1
grep -E "^pair|^single|TC_PERF|rror" gpurun_out/pair.log | cut -c1-112
