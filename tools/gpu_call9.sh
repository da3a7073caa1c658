#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/prof9.log
INRF_TC_PROF=1 timeout 300 python tools/tc_perf.py 160000 >> gpurun_out/prof9.log 2>&1
INRF_TC_PROF=1 INRF_TC_NOWEIGHTS=1 timeout 300 python tools/tc_perf.py 160000 >> gpurun_out/prof9.log 2>&1
grep -E "TC_PERF|rror" gpurun_out/prof9.log
grep -E "TCPROF" gpurun_out/prof9.log | grep -A12 "role=producer" | tail -28
