#!/bin/bash
# GPU test-suite (optionally a -k filter) + timings of the section-8f rows
mkdir -p gpurun_out
export INRF_TC_CHECK=1
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider $1 $2 > gpurun_out/pytest_gpu.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu.log; grep -E "^E  |FAILED" gpurun_out/pytest_gpu.log | cut -c1-300 | head -30
INRF_TC_CHECK=0 timeout 300 python tests/tools/aux_bench.py > gpurun_out/aux_bench.log 2>&1; grep -E "AUX|rror" gpurun_out/aux_bench.log
