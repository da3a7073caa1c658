#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/prof4.log
INRF_TC_PROF=1 INRF_TC_CLUSTER=1 timeout 300 python tools/tc_perf.py 160000 >> gpurun_out/prof4.log 2>&1
INRF_TC_PROF=1 INRF_TC_CLUSTER=2 timeout 300 python tools/tc_perf.py 160000 >> gpurun_out/prof4.log 2>&1
INRF_TC_NOWEIGHTS=1 INRF_TC_PROF=1 INRF_TC_CLUSTER=1 timeout 300 python tools/tc_perf.py 160000 >> gpurun_out/prof4.log 2>&1
grep -E "TC_PERF|TCPROF|rror" gpurun_out/prof4.log | awk '!seen[$0]++' | head -150
