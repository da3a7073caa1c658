#!/bin/bash
# One single-GPU session on the B200 box (run with gpurun): GPU test-suite, smoke(), the default bench, the reference
# arm and the two secondary workloads.  Everything lands in gpurun_out/.  `tools/gpu_round.sh profile` adds the ncu evidence.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/gpu.txt
python oracle/build_ref.py --check >> gpurun_out/gpu.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python bench.py --workload room0 --steps 5 --warmup 3 > gpurun_out/bench_room0.json 2> gpurun_out/bench_room0.err
timeout 600 python bench.py --workload room0_train --steps 20 --warmup 5 > gpurun_out/bench_room0_train.json 2> gpurun_out/bench_room0_train.err
[ "$1" == "profile" ] && bash tools/gpu_profile.sh > gpurun_out/profile.log 2>&1
tail -n 6 gpurun_out/pytest_gpu.log | cut -c1-300; tail -n 1 gpurun_out/smoke.log
for f in bench_ref bench bench_room0 bench_room0_train; do tail -n 1 gpurun_out/$f.json | cut -c1-700; tail -n 2 gpurun_out/$f.err | cut -c1-300; done
