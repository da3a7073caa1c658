#!/bin/bash
# full ncu captures of the training GEMM kernels (fine network of the first step of tests/tools/train_target.py):
# forward with stash, the trunk-output dX GEMM, the chained trunk dX kernel, dW
mkdir -p gpurun_out
rm -f gpurun_out/prof_train_*.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mlp_tc -s 1 -c 1 -o gpurun_out/prof_train_fwd \
  python tests/tools/train_target.py > gpurun_out/prof_train.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_dx -s 3 -c 1 -o gpurun_out/prof_train_dx \
  python tests/tools/train_target.py >> gpurun_out/prof_train.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dx_chain -s 0 -c 1 -o gpurun_out/prof_train_dxchain \
  python tests/tools/train_target.py >> gpurun_out/prof_train.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_dw -s 0 -c 1 -o gpurun_out/prof_train_dw \
  python tests/tools/train_target.py >> gpurun_out/prof_train.log 2>&1
ls -la gpurun_out/*.ncu-rep
