#!/bin/bash
# round 2 session b: status + drop-in tests, reference arm, new bench (short job)
mkdir -p gpurun_out
python oracle/build_ref.py --check > gpurun_out/ref_check.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_status.py tests/test_gpu_dropin.py -q -p no:cacheprovider > gpurun_out/pytest_b.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 900 python bench.py --steps 3 --warmup 3 --views 8 > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/ref_check.txt; tail -n 40 gpurun_out/pytest_b.log; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
