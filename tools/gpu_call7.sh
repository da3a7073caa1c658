#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/prof7.log
INRF_TC_CHECK=1 timeout 300 python tools/tc_debug.py > gpurun_out/tc_debug.log 2>&1
tail -3 gpurun_out/tc_debug.log
INRF_TC_PROF=1 timeout 300 python tools/tc_perf.py 160000 >> gpurun_out/prof7.log 2>&1
INRF_TC_CLUSTER=2 timeout 300 python tools/tc_perf.py 160000 >> gpurun_out/prof7.log 2>&1
INRF_TC_NOWEIGHTS=1 timeout 300 python tools/tc_perf.py 160000 >> gpurun_out/prof7.log 2>&1
INRF_TC_BIASMMA=0 timeout 300 python tools/tc_perf.py 160000 >> gpurun_out/prof7.log 2>&1
timeout 300 python tools/tc_perf.py 160000 ssr >> gpurun_out/prof7.log 2>&1
grep -E "TC_PERF|rror" gpurun_out/prof7.log
grep -E "TCPROF" gpurun_out/prof7.log | tail -32
bash tools/gpu_round.sh > gpurun_out/round.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_*.log; cat gpurun_out/bench.json
