#!/bin/sh
# tools/build_variant.sh NAME [-DFLAG ...]  ->  restir_embree_b200/variants/NAME.so (same sources, extra nvcc flags)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; shift
mkdir -p "$ROOT/restir_embree_b200/variants"
cd "$ROOT/restir_embree_b200/csrc"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-O2 -shared "$@" \
  -o "$ROOT/restir_embree_b200/variants/$NAME.so" restir_b200.cu
echo "built variants/$NAME.so"
