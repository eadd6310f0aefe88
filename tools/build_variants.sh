#!/bin/bash
# A/B builds of the direct kernel: tools/variants/lib_<name>.so differ only in frisk_direct.cu's compile-time switches.
# usage: tools/build_variants.sh name1 "-DX=1 -DY=0" name2 "..." ...
set -e
cd "$(dirname "$0")/../frisk_b200/csrc"
make >/dev/null
mkdir -p ../../tools/variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
while [ $# -ge 2 ]; do
  name=$1; defs=$2; shift 2
  kern=build/frisk_kernels.o
  ( case "$defs" in *FRISK_SERIES_LOG2*|*FRISK_DIRECT_DEFAULT*)
      kern=build/kernels_$name.o
      nvcc $ARCH -std=c++17 -O3 -lineinfo -Xcompiler -fPIC,-O3,-Wall -diag-suppress 186 $defs -c frisk_kernels.cu -o $kern & ;;
    esac
    nvcc $ARCH -std=c++17 -O3 -lineinfo -Xcompiler -fPIC,-O3,-Wall -diag-suppress 186 $defs -c frisk_direct.cu -o build/direct_$name.o && wait &&
    nvcc $ARCH -shared -o ../../tools/variants/lib_$name.so $kern build/direct_$name.o build/frisk_general.o build/frisk_ingest.o build/frisk_features.o build/frisk_host.o -lpthread &&
    echo built $name ) &
done
wait
