#!/bin/bash
# kernel-only experiments: cluster size, weights off
mkdir -p gpurun_out
: > gpurun_out/exp.log
for cl in 2 4 1; do
  echo "cluster $cl" >> gpurun_out/exp.log
  INRF_TC_CLUSTER=$cl timeout 300 python tests/tools/tc_perf.py 160000 >> gpurun_out/exp.log 2>&1
done
echo "noweights cl2" >> gpurun_out/exp.log
INRF_TC_NOWEIGHTS=1 timeout 300 python tests/tools/tc_perf.py 160000 >> gpurun_out/exp.log 2>&1
echo "noweights cl1" >> gpurun_out/exp.log
INRF_TC_CLUSTER=1 INRF_TC_NOWEIGHTS=1 timeout 300 python tests/tools/tc_perf.py 160000 >> gpurun_out/exp.log 2>&1
grep -E "^cluster|^noweights|TC_PERF|rror" gpurun_out/exp.log | cut -c1-190
