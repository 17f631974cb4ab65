#!/bin/bash
# Development helper: build a variant of the library with extra -D flags for nw_kernels.cu into tools/ab/NAME.so
# usage: tools/build_variant.sh NAME "-DDYNA_ROWS2_THREADS=384"
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
mkdir -p tools/ab/obj_$NAME
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -Iinclude -O3 -lineinfo -std=c++17 -ccbin /usr/bin/g++ \
  -Xcompiler -fPIC,-O2,-fvisibility=hidden -Xptxas -v "$@" -c dynaalign_b200/csrc/nw_kernels.cu -o tools/ab/obj_$NAME/nw_kernels.o 2> tools/ab/obj_$NAME/ptxas.log
OBJS=$(ls dynaalign_b200/csrc/build/*.o | grep -v nw_kernels.o)
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -Xcompiler -fPIC -o tools/ab/$NAME.so $OBJS tools/ab/obj_$NAME/nw_kernels.o
echo built tools/ab/$NAME.so
