#!/bin/bash
# builds intrinsicnerf_b200/csrc/libinrf_prof.so: same sources with -DTC2_PROF (wait-cycle counters in k_mlp_tc2)
set -e
cd "$(dirname "$0")/../intrinsicnerf_b200/csrc"
mkdir -p prof_obj
for f in pack stages mlp_fp32 mlp_bwd_fp32 mlp_tc mlp_tc2 train_tc cluster loss frame api; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -DTC2_PROF -c $f.cu -o prof_obj/$f.o &
done
wait
nvcc -shared -o libinrf_prof.so prof_obj/*.o -lcudart
rm -rf prof_obj
echo built libinrf_prof.so
