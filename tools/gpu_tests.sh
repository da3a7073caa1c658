#!/bin/bash
# GPU test-suite only (optionally a -k filter)
mkdir -p gpurun_out
export INRF_TC_CHECK=1
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider $1 $2 > gpurun_out/pytest_gpu.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu.log; grep -E "^E  |FAILED" gpurun_out/pytest_gpu.log | cut -c1-300 | head -30
