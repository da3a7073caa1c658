#!/bin/bash
# pair kernel: wait-cycle profile (prof build) + one full ncu capture
mkdir -p gpurun_out
INRF_LIB=$PWD/intrinsicnerf_b200/csrc/libinrf_prof.so timeout 300 python tests/tools/tc_perf.py 160000 > gpurun_out/pairprof.log 2>&1
grep TC2PROF gpurun_out/pairprof.log | tail -${PROF_LINES:-40}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mlp_tc2 -s 4 -c 1 -f -o gpurun_out/prof_mlp_tc2 \
  python tests/tools/profile_target.py 40000 tc > gpurun_out/prof_target.log 2>&1
tail -3 gpurun_out/prof_target.log
