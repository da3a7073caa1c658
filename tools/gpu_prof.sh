#!/bin/bash
mkdir -p gpurun_out
INRF_TC_PROF=1 timeout 300 python tests/tools/tc_perf.py 160000 > gpurun_out/tc_prof.log 2>&1
grep -E "TCPROF|TC_PERF" gpurun_out/tc_prof.log | tail -40
