#!/bin/bash
# build_variant.sh NAME [-DFLAG=..]...: compile an A/B variant of the library into build_variants/libgx_NAME.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build_variants
env -u CC -u CXX nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -diag-suppress 20091 \
  -shared -Xcompiler -fPIC "$@" -o build_variants/libgx_$name.so galax_b200/csrc/gx_kernels.cu
