#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mlp_tc -s 7 -c 1 -o gpurun_out/prof_mlp_tc_v2 \
  python tools/profile_target.py 40000 tc > gpurun_out/prof_target.log 2>&1
ls -la gpurun_out/*.ncu-rep
