#!/bin/bash
# End-of-session evidence with the final defaults: GPU test-suite, smoke, default bench, room0 bench, and the two ncu passes
# whose results bench.py / DESIGN quote (launch list of one bench step; one full capture of a fused fine-pass k_mlp_tc launch)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --steps 4 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python bench.py --workload room0 --steps 5 --warmup 3 > gpurun_out/bench_room0.json 2> gpurun_out/bench_room0.err
rm -f gpurun_out/*.ncu-rep gpurun_out/launches*.csv
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --views 1 --steps 1 --warmup 3 --no-cpu-baseline --no-config5 > gpurun_out/bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_mlp_tc -s 3 -c 1 -o gpurun_out/prof_mlp_tc_fused_fine \
  python tests/tools/profile_target.py 40000 tc > gpurun_out/prof_target.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log | cut -c1-200; tail -n 1 gpurun_out/smoke.log; tail -n 1 gpurun_out/bench.json | cut -c1-260; tail -n 1 gpurun_out/bench_room0.json | cut -c1-260
ls -la gpurun_out | grep -E "ncu-rep|launches"
