#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tests/tools/fused_perf.py > gpurun_out/fused_perf.log 2>&1
timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_wide.py tests/test_gpu_train_tc.py tests/test_gpu_dropin.py tests/test_gpu_convergence.py -q -p no:cacheprovider -s > gpurun_out/pytest_d.log 2>&1
tail -3 gpurun_out/fused_perf.log; grep -E "passed|failed|FAILED|^object/|^ssr/|loss:" gpurun_out/pytest_d.log | cut -c1-900
