#!/bin/bash
# builds an A/B variant of the library: tools/build_variant.sh <suffix> <extra nvcc flags...>   ->  rgk_b200/librgk_b200_<suffix>.so
# (selected at run time with RGK_B200_LIB=<path>, read by rgk_b200/abi.py -- the Python binding, not the library)
set -e
sfx=$1; shift
cd "$(dirname "$0")/.."
mkdir -p build/var_$sfx
for f in api.cu trace.cu render.cu probe.cu host_scene.cpp host_bvh.cpp; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off -Iinclude -Irgk_b200/csrc "$@" -x cu -c rgk_b200/csrc/$f -o build/var_$sfx/${f%.*}.o &
done; wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o rgk_b200/librgk_b200_$sfx.so build/var_$sfx/*.o -lcudart
