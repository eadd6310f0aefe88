#!/bin/bash
# usage: tools/build_variant.sh NAME FILE.cu "FLAGS"  ->  tools/variants/libfrisk_NAME.so
# (the library with ONE translation unit recompiled with extra -D flags; A/B runs load it through FRISK_B200_LIB)
set -e
cd "$(dirname "$0")/../frisk_b200/csrc"
make -s >/dev/null
name=$1; file=$2; flags=$3
obj=/tmp/variant_${name}.o
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC,-O3,-Wall -diag-suppress 186 $flags -c $file -o $obj
others=$(ls build/*.o | grep -v "build/$(basename $file .cu).o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/variants/libfrisk_${name}.so $others $obj -lpthread
echo built tools/variants/libfrisk_${name}.so
