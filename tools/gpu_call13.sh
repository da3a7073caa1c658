#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/prof13.log
INRF_TC_PROF=1 timeout 300 python tools/tc_perf.py 160000 >> gpurun_out/prof13.log 2>&1
grep -E "TC_PERF|rror" gpurun_out/prof13.log
grep -E "TCTRACE" gpurun_out/prof13.log | tail -11
