#!/bin/bash
# r01e session: GPU test-suite, section-8f timings, training-step timing, both bench arms
mkdir -p gpurun_out
export INRF_TC_CHECK=1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu.log; grep -E "^E  |FAILED" gpurun_out/pytest_gpu.log | cut -c1-300 | head -30
unset INRF_TC_CHECK
timeout 300 python tests/tools/aux_bench.py > gpurun_out/aux_bench.log 2>&1; grep -E "AUX|rror" gpurun_out/aux_bench.log
timeout 300 python tests/tools/train_bench.py > gpurun_out/train_bench.log 2>&1; grep -E "TRAIN|rror" gpurun_out/train_bench.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref.json
