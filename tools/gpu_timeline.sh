#!/bin/bash
# Development: cycles per tile (INRF_TC_PROF=1: clock64 totals of CTA 0) and time per tile of the issuer variants, alternating
# to average out clock drift; then the clock64 timeline of the issuer and two epilogue warps of CTA 0 over two steady-state
# tiles (library built with -DINRF_TC_TIMELINE as csrc/libinrf_tl.so)
mkdir -p gpurun_out
: > gpurun_out/exp4.log
for rep in 1 2; do
for cfg in "4 0" "4 64" "0 0" "0 64"; do
  set -- $cfg
  INRF_TC_PROF=1 INRF_TC_EXP=$1 INRF_TC_SPLIT=$2 timeout 200 python tests/tools/tc_perf.py 160000 > gpurun_out/tmp.log 2>&1
  cyc=$(grep "TCPROF role=issuer" gpurun_out/tmp.log | tail -1 | sed 's/.*total_cycles=\([0-9]*\) n_iter=\([0-9]*\)/\1 \2/')
  grep TC_PERF gpurun_out/tmp.log | sed "s/^TC_PERF/TC_PERF exp=$1 issuer_cycles_n_iter=($cyc)/" >> gpurun_out/exp4.log
done
done
grep -E "TC_PERF|rror" gpurun_out/exp4.log | cut -c1-200
export INRF_LIB=$PWD/intrinsicnerf_b200/csrc/libinrf_tl.so
for sp in 0 64; do
  INRF_TC_SPLIT=$sp timeout 200 python tests/tools/tc_perf.py 160000 > gpurun_out/timeline_lean$sp.log 2>&1
  grep -E "TC_PERF" gpurun_out/timeline_lean$sp.log | cut -c1-200
done
