#!/bin/bash
# compute-sanitizer memcheck over the new paths at small sizes (tensor-core training, frame kernels, merge fast path)
mkdir -p gpurun_out
export INRF_TC_CHECK=1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_fused.py -m gpu -q -p no:cacheprovider -x -k "object_fused_equals_staged and (127 or 129)" > gpurun_out/sanitize_fused.log 2>&1
echo "fused rc=$?"; grep -E "ERROR SUMMARY|Invalid|passed|failed" gpurun_out/sanitize_fused.log | head -8
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python tests/tools/tcbwd_debug.py ssr 300 ep > gpurun_out/sanitize_tcbwd.log 2>&1
echo "tcbwd rc=$?"; grep -E "ERROR SUMMARY|Invalid|TCBWD" gpurun_out/sanitize_tcbwd.log | head -8
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_frame.py -m gpu -q -p no:cacheprovider -x -k "not end_to_end and not 800" > gpurun_out/sanitize_frame.log 2>&1
echo "frame rc=$?"; grep -E "ERROR SUMMARY|Invalid|passed|failed" gpurun_out/sanitize_frame.log | head -8
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_stages.py -m gpu -q -p no:cacheprovider -x -k "merge_sorted" > gpurun_out/sanitize_merge.log 2>&1
echo "merge rc=$?"; grep -E "ERROR SUMMARY|Invalid|passed|failed" gpurun_out/sanitize_merge.log | head -8
