#!/bin/bash
# Development: clock64 timeline (build csrc/libinrf_tl.so first, HERE: python -m intrinsicnerf_b200.build --timeline) of the issuer,
# two epilogue warps and back-end warp 12 of CTA 0 over two steady-state tiles - of the bare TS launch and of both fused launches
mkdir -p gpurun_out
export INRF_LIB=$PWD/intrinsicnerf_b200/csrc/libinrf_tl.so
timeout 200 python tests/tools/fused_timeline.py > gpurun_out/timeline_fused.log 2>&1
grep -c TCTL gpurun_out/timeline_fused.log
timeout 200 python tests/tools/tc_perf.py 160000 > gpurun_out/timeline_bare.log 2>&1
grep -E "TC_PERF" gpurun_out/timeline_bare.log | cut -c1-200
