#!/bin/bash
# ncu evidence: launch list of the bench command + one full capture of the dominant kernel.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mlp_tc -s 7 -c 1 -o gpurun_out/prof_mlp_tc_r01b \
  python tests/tools/profile_target.py 40000 tc > gpurun_out/prof_target.log 2>&1
ls -la gpurun_out
