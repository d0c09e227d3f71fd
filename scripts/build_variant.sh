#!/bin/bash
# Development aid: build libnnpops_b200 with extra defines for ONE source file into nnpops_b200/variants/<name>.so (selected at run time
# with NNPOPS_LIB_PATH).  usage: build_variant.sh <name> <source.cu> <defines...>
set -e
cd "$(dirname "$0")/.."
name=$1; src=$2; shift 2
mkdir -p nnpops_b200/variants
obj=nnpops_b200/variants/$name.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr -diag-suppress 177 "$@" -c nnpops_b200/csrc/$src -o $obj
others=$(ls nnpops_b200/build/*.o | grep -v "/$(basename $src .cu).o")
/usr/local/cuda/bin/nvcc -shared -o nnpops_b200/variants/$name.so $obj $others -lcudart -lcuda -lcufft
echo nnpops_b200/variants/$name.so
