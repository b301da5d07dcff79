#!/bin/bash
# Build kernel variants of np2_kernels.cu side by side (here, on the CPU box: nvcc cross-compiles) so that one GPU
# call can A/B them:   tools/build_variants.sh name1 "<-D flags>" name2 "<-D flags>" ...
# -> motionplanningtoolkit_b200/variants/libmptcc_<name>.so, selected at run time with MPTCC_LIB=<path>.
set -e
cd "$(dirname "$0")/../motionplanningtoolkit_b200/csrc"
make > /dev/null
mkdir -p ../variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  (
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC \
      -Xcompiler -ffp-contract=off -Xptxas -v $flags -c np2_kernels.cu -o ../variants/np2_$name.o 2> ../variants/np2_$name.log
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libmptcc_$name.so \
      bvh.o mptcc_api.o collide_kernel.o split_kernels.o ../variants/np2_$name.o self_kernel.o edge_kernels.o -lpthread
    echo "$name: $flags :: $(grep -A2 'narrowphase2_kernelILb0' ../variants/np2_$name.log | grep -oE 'Used [0-9]+ registers|[0-9]+ bytes spill stores' | tr '\n' ' ')"
  ) &
done
wait
