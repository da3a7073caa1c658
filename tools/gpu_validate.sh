#!/bin/bash
# Validation + timing of the final k_mlp_tc defaults: GPU test-suite, cycles per tile of the bare fine launch (INRF_TC_PROF=1),
# fused vs staged chunk, a short default bench
mkdir -p gpurun_out
: > gpurun_out/split.log
timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log >> gpurun_out/split.log
timeout 200 python tests/tools/fused_perf.py >> gpurun_out/split.log 2>&1
INRF_TC_PROF=1 timeout 200 python tests/tools/tc_perf.py 160000 > gpurun_out/tmp.log 2>&1
grep "TCPROF role=issuer" gpurun_out/tmp.log | tail -1 >> gpurun_out/split.log; grep TC_PERF gpurun_out/tmp.log >> gpurun_out/split.log
timeout 300 python bench.py --steps 2 --warmup 3 --views 8 --no-cpu-baseline --no-config5 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
grep -E "TC_PERF|TCPROF|FUSED_PERF|passed|failed|rror" gpurun_out/split.log | cut -c1-420; tail -n 1 gpurun_out/bench_quick.json | cut -c1-300
