#!/bin/bash
# stage tests + bench (no cpu baseline) - quick regression check after a kernel change
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stages.py tests/test_gpu_render.py -m gpu -q -p no:cacheprovider -x > gpurun_out/pytest_quick.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_quick.log; grep -E "^E  |FAILED" gpurun_out/pytest_quick.log | cut -c1-300 | head -10
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; cut -c1-160 gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_quick.err
