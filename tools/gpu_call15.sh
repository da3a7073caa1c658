#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/prof15.log
INRF_TC_CHECK=1 timeout 300 python tools/tc_debug.py > gpurun_out/tc_debug.log 2>&1
tail -2 gpurun_out/tc_debug.log
timeout 300 python tools/tc_perf.py 160000 >> gpurun_out/prof15.log 2>&1
INRF_TC_CLUSTER=2 timeout 300 python tools/tc_perf.py 160000 >> gpurun_out/prof15.log 2>&1
INRF_TC_NOWEIGHTS=1 timeout 300 python tools/tc_perf.py 160000 >> gpurun_out/prof15.log 2>&1
timeout 300 python tools/tc_perf.py 160000 ssr >> gpurun_out/prof15.log 2>&1
INRF_TC_PROF=1 timeout 300 python tools/tc_perf.py 160000 >> gpurun_out/prof15.log 2>&1
grep -E "TC_PERF|rror" gpurun_out/prof15.log
grep -E "TCTRACE" gpurun_out/prof15.log | tail -11
