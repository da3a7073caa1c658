#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/prof5.log
INRF_TC_CHECK=1 timeout 300 python tools/tc_perf.py 4096 >> gpurun_out/prof5.log 2>&1
INRF_TC_PROF=1 INRF_TC_CLUSTER=2 timeout 300 python tools/tc_perf.py 160000 >> gpurun_out/prof5.log 2>&1
INRF_TC_CLUSTER=1 timeout 300 python tools/tc_perf.py 160000 >> gpurun_out/prof5.log 2>&1
INRF_TC_NOWEIGHTS=1 INRF_TC_CLUSTER=1 timeout 300 python tools/tc_perf.py 160000 >> gpurun_out/prof5.log 2>&1
timeout 300 python tools/tc_perf.py 160000 ssr >> gpurun_out/prof5.log 2>&1
grep -E "TC_PERF|rror" gpurun_out/prof5.log
grep -E "TCPROF" gpurun_out/prof5.log | tail -45
bash tools/gpu_round.sh > gpurun_out/round.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_*.log; cat gpurun_out/bench.json
