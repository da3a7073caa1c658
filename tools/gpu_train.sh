#!/bin/bash
# training path: parity tests, the training-step bench line, and a warm-cache launch list of three steps
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_tc.py tests/test_gpu_stages.py tests/test_gpu_convergence.py tests/test_gpu_dropin.py tests/test_gpu_status.py -q -x -p no:cacheprovider > gpurun_out/pytest_train.log 2>&1
timeout 600 python bench.py --workload room0_train --steps 20 --warmup 5 > gpurun_out/bench_room0_train.json 2> gpurun_out/bench_room0_train.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 600 --csv --log-file gpurun_out/launches_train_warm.csv \
  python tests/tools/train_target.py > gpurun_out/train_under_ncu_warm.log 2>&1
tail -n 5 gpurun_out/pytest_train.log | cut -c1-300
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_room0_train.json").read().strip().splitlines()[-1])
print("TRAIN ms_per_step", d["ms_per_step"], "graph", d["config5"].get("ms_per_step_cuda_graph"), "launches/step", d["config5"].get("libinrf_launches_per_step"))
PY
tail -n 2 gpurun_out/bench_room0_train.err | cut -c1-300
