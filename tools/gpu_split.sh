#!/bin/bash
# Validation + timing of the current k_mlp_tc variants: whole GPU test-suite, bare launches (object TS / SSR hybrid /
# training forward), fused vs staged chunk, a short default bench.  INRF_TC_EXP=4 selects the first (pre-lean) issuers.
mkdir -p gpurun_out
: > gpurun_out/split.log
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log >> gpurun_out/split.log
timeout 200 python tests/tools/fused_perf.py >> gpurun_out/split.log 2>&1
timeout 200 python tests/tools/tc_perf.py 160000 >> gpurun_out/split.log 2>&1
timeout 200 python tests/tools/tc_perf.py 160000 ssr >> gpurun_out/split.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --views 8 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
grep -E "TC_PERF|STASH_PERF|FUSED_PERF|passed|failed|rror" gpurun_out/split.log | cut -c1-420; tail -n 1 gpurun_out/bench_quick.json | cut -c1-300
