#!/bin/bash
# tensor-core training path: parity tests + step timing
mkdir -p gpurun_out
export INRF_TC_CHECK=1
timeout 900 python -m pytest tests/test_gpu_train_tc.py tests/test_gpu_render.py tests/test_gpu_stages.py -m gpu -q -p no:cacheprovider ${1:--x} > gpurun_out/pytest_train.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_train.log; grep -E "^E  |FAILED" gpurun_out/pytest_train.log | cut -c1-400 | head -30
unset INRF_TC_CHECK
timeout 300 python tests/tools/train_bench.py > gpurun_out/train_bench.log 2>&1; grep -E "TRAIN|rror" gpurun_out/train_bench.log
