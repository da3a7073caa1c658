#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/variants.log
for cfg in "2 1" "1 1" "4 1" "2 0" "1 0"; do
  set -- $cfg
  INRF_TC_CLUSTER=$1 INRF_TC_BIASMMA=$2 INRF_TC_CHECK=1 timeout 300 python tools/tc_perf.py 4096 >> gpurun_out/variants.log 2>&1
  INRF_TC_CLUSTER=$1 INRF_TC_BIASMMA=$2 timeout 300 python tools/tc_perf.py 160000 >> gpurun_out/variants.log 2>&1
done
INRF_TC_CHECK=1 timeout 300 python tools/tc_perf.py 4096 ssr >> gpurun_out/variants.log 2>&1
timeout 300 python tools/tc_perf.py 160000 ssr >> gpurun_out/variants.log 2>&1
grep -E "TC_PERF|Error|error" gpurun_out/variants.log
