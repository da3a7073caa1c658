#!/bin/bash
# Validation + timing of the current k_mlp_tc variants: whole GPU test-suite, bare launches (object TS / SSR hybrid /
# training forward), fused vs staged chunk, a short default bench.  INRF_TC_EXP=4 selects the first (pre-lean) issuers.
mkdir -p gpurun_out
: > gpurun_out/split.log
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log >> gpurun_out/split.log
for e in 4 0; do
  INRF_TC_EXP=$e timeout 200 python tests/tools/fused_perf.py 2>&1 | sed "s/^FUSED_PERF/FUSED_PERF exp=$e/" >> gpurun_out/split.log
done
timeout 200 python tests/tools/tc_perf.py 160000 >> gpurun_out/split.log 2>&1
timeout 200 python tests/tools/stash_perf.py >> gpurun_out/split.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --views 8 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
INRF_TC_EXP=4 timeout 600 python bench.py --steps 3 --warmup 3 --views 8 --no-cpu-baseline > gpurun_out/bench_quick_exp4.json 2> gpurun_out/bench_quick_exp4.err
grep -E "TC_PERF|STASH_PERF|FUSED_PERF|passed|failed|rror" gpurun_out/split.log | cut -c1-420; tail -n 1 gpurun_out/bench_quick.json | cut -c1-700;  tail -n 1 gpurun_out/bench_quick_exp4.json | cut -c1-300
