#!/bin/bash
bash tools/gpu_variants.sh
bash tools/gpu_round.sh > gpurun_out/round.log 2>&1
tail -40 gpurun_out/round.log
