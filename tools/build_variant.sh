#!/bin/bash
# A/B builds: tools/build_variant.sh NAME "-DFOO=1 ..."  ->  ikd-tree_b200/variants/libikd_b200_NAME.so (use with IKD_LIB_PATH)
set -e
name=$1; flags=$2
cd "$(dirname "$0")/../ikd-tree_b200"
mkdir -p variants/obj_$name
for f in ikd_capi ikd_build ikd_knn ikd_range ikd_update ikd_plane; do
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -Xcompiler -O2 --expt-relaxed-constexpr $flags -c csrc/$f.cu -o variants/obj_$name/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/libikd_b200_$name.so variants/obj_$name/*.o -lcudart
rm -rf variants/obj_$name
ls -la variants/libikd_b200_$name.so
