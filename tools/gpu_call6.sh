#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/prof6.log
for e in 0 1 2; do
INRF_TC_EXP=$e INRF_TC_NOWEIGHTS=1 INRF_TC_CLUSTER=1 timeout 300 python tools/tc_perf.py 160000 >> gpurun_out/prof6.log 2>&1
done
grep -E "TC_PERF|rror" gpurun_out/prof6.log
