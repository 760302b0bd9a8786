#!/bin/bash
# usage: mkvariant.sh <name> <file.cu> <extra nvcc flags...>   -> build_variants/lib_<name>.so with one object rebuilt
set -e
cd /root/repo/hystrath_b200/csrc
name=$1; src=$2; shift 2
NVF="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --fmad=false -std=c++17 -Xcompiler -fPIC,-fopenmp,-ffp-contract=off,-O3 -I../../include"
obj=/tmp/var_${name}.o
nvcc $NVF "$@" -Xptxas -v -c $src -o $obj 2>&1 | grep -A2 "moveKernel\|collideLane\|sampleKernel" | grep -E "spill|Used" | head -4
objs=""
for o in engine.o kernels_move.o kernels_sort.o kernels_collide.o kernels_init.o host_mesh.o; do
  if [ "$o" == "${src%.cu}.o" ]; then objs="$objs $obj"; else objs="$objs $o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o /root/repo/build_variants/lib_${name}.so $objs -Xcompiler -fopenmp -lgomp -ldl
echo built lib_${name}.so
