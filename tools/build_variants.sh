#!/bin/bash
# Build kernel variants of one translation unit (UNIT=np2_kernels by default, e.g. UNIT=np3_kernels) side by side (here, on the CPU box: nvcc cross-compiles) so that one GPU
# call can A/B them:   tools/build_variants.sh name1 "<-D flags>" name2 "<-D flags>" ...
# -> motionplanningtoolkit_b200/variants/libmptcc_<name>.so, selected at run time with MPTCC_LIB=<path>.
set -e
cd "$(dirname "$0")/../motionplanningtoolkit_b200/csrc"
make > /dev/null
UNIT=${UNIT:-np2_kernels}
KSYM=${KSYM:-narrowphase[23]_kernelILb0}
OBJS="bvh.o mptcc_api.o collide_kernel.o split_kernels.o np2_kernels.o np3_kernels.o self_kernel.o edge_kernels.o"
mkdir -p ../variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  (
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC \
      -Xcompiler -ffp-contract=off -Xptxas -v $flags -c $UNIT.cu -o ../variants/${UNIT}_$name.o 2> ../variants/${UNIT}_$name.log
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libmptcc_$name.so \
      ${OBJS/$UNIT.o/../variants/${UNIT}_$name.o} -lpthread
    echo "$name: $flags :: $(grep -E -A2 "$KSYM" ../variants/${UNIT}_$name.log | grep -oE 'Used [0-9]+ registers|[0-9]+ bytes spill stores' | tr '\n' ' ')"
  ) &
done
wait
