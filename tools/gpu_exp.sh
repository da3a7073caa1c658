#!/bin/bash
# kernel-only experiments: fused vs staged chunk, bare k_mlp_tc timing (object / SSR), weights-off ablation, barrier-wait profile
mkdir -p gpurun_out
timeout 300 python tests/tools/fused_perf.py > gpurun_out/fused_perf.log 2>&1
timeout 300 python tests/tools/tc_perf.py 160000 > gpurun_out/exp.log 2>&1
timeout 300 python tests/tools/tc_perf.py 160000 ssr >> gpurun_out/exp.log 2>&1
INRF_TC_NOWEIGHTS=1 timeout 300 python tests/tools/tc_perf.py 160000 >> gpurun_out/exp.log 2>&1
INRF_TC_PROF=1 timeout 300 python tests/tools/tc_perf.py 160000 > gpurun_out/tc_prof.log 2>&1
tail -1 gpurun_out/fused_perf.log; grep -E "TC_PERF|rror" gpurun_out/exp.log | cut -c1-200; grep TCPROF gpurun_out/tc_prof.log | tail -12
