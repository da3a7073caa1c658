#!/bin/bash
# GPU test-suite + a short default bench (development loop)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 --views 8 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
tail -n 6 gpurun_out/pytest_gpu.log | cut -c1-300; tail -n 1 gpurun_out/bench_quick.json | cut -c1-2600; tail -n 3 gpurun_out/bench_quick.err | cut -c1-300
