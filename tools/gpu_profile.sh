#!/bin/bash
# ncu evidence (summarise afterwards, here, with `python tools/summarize_profiles.py r02`): launch list of one bench step,
# one full capture of a FUSED fine-pass k_mlp_tc launch, launch list of three training steps.
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/launches*.csv
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --views 1 --steps 1 --warmup 3 --no-cpu-baseline --no-config5 > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mlp_tc -s 3 -c 1 -o gpurun_out/prof_mlp_tc_fused_fine \
  python tests/tools/profile_target.py 40000 tc > gpurun_out/prof_target.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_train.csv \
  python tests/tools/train_target.py > gpurun_out/train_under_ncu.log 2>&1
# the same three steps with warm caches (--cache-control none): per-kernel durations closer to the ones inside a running step
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 600 --csv --log-file gpurun_out/launches_train_warm.csv \
  python tests/tools/train_target.py > gpurun_out/train_under_ncu_warm.log 2>&1
ls -la gpurun_out | head -30
