#!/bin/bash
# bring-up of the tensor-core training path
mkdir -p gpurun_out
for v in "object 1000" "ssr 700" "ssr 300 ep"; do
  timeout 300 python tests/tools/tcbwd_debug.py $v > gpurun_out/tcbwd_$(echo $v | tr ' ' '_').log 2>&1
  grep -E "^\[fwd|^\[stash\] (H0|H7|PE)|TCBWD|rror|^\[grad\]|^\[dz\]" gpurun_out/tcbwd_$(echo $v | tr ' ' '_').log | cut -c1-150 | tail -32
done
