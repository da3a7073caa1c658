#!/bin/bash
# r01e session: GPU test-suite, section-8f / training timings, both bench arms, ncu evidence (bench launch list, full
# captures of k_mlp_tc, k_gemm_dx, k_gemm_dw, launch list of a training step)
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/launches*.csv
export INRF_TC_CHECK=1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu.log; grep -E "^E  |FAILED" gpurun_out/pytest_gpu.log | cut -c1-300 | head -30
unset INRF_TC_CHECK
timeout 300 python tests/tools/aux_bench.py > gpurun_out/aux_bench.log 2>&1; grep -E "AUX|rror" gpurun_out/aux_bench.log
timeout 300 python tests/tools/train_bench.py > gpurun_out/train_bench.log 2>&1; grep -E "TRAIN|rror" gpurun_out/train_bench.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref.json
if [ "$1" == "profile" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mlp_tc -s 7 -c 1 -f -o gpurun_out/prof_mlp_tc \
    python tests/tools/profile_target.py 40000 tc > gpurun_out/prof_target.log 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_train.csv \
    python tests/tools/train_target.py > gpurun_out/train_under_ncu.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_dx -s 28 -c 1 -f -o gpurun_out/prof_gemm_dx \
    python tests/tools/train_target.py > gpurun_out/prof_dx.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_dw -s 3 -c 1 -f -o gpurun_out/prof_gemm_dw \
    python tests/tools/train_target.py > gpurun_out/prof_dw.log 2>&1
  ls -la gpurun_out | tail -20
fi
