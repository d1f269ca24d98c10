#!/bin/bash
# A/B builds of the product library: scripts/build_variant.sh NAME "-DMACRO=..." -> lrr_b200/csrc/build/variants/NAME.so
# (select it with LRR_B200_LIB=lrr_b200/csrc/build/variants/NAME.so; development aid only)
set -e
cd "$(dirname "$0")/../lrr_b200/csrc"
name=$1; shift
out=build/variants/$name; mkdir -p $out
for f in lrr_grid lrr_icp lrr_tile lrr_select lrr_metrics lrr_objcov lrr_cdc lrr_batch lrr_regions lrr_capi; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcudafe --diag_suppress=550 -Xcudafe --diag_suppress=177 "$@" -c $f.cu -o $out/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/$name.so $out/*.o -lcudart
echo built build/variants/$name.so
