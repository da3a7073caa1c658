#!/bin/bash
# run an arbitrary python script on the GPU box, log to gpurun_out/cmd.log
mkdir -p gpurun_out
timeout 900 python "$@" > gpurun_out/cmd.log 2>&1
tail -40 gpurun_out/cmd.log
