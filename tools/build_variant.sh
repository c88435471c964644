#!/bin/bash
# usage: tools/build_variant.sh <name> [-DMACRO=...]...   -> raycastworlds.jl_b200/lib/variants/librcw_b200_<name>.so
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
name=$1; shift
mkdir -p "$ROOT/raycastworlds.jl_b200/lib/variants"
cd "$ROOT/raycastworlds.jl_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false \
  -Xcompiler -fPIC,-ffp-contract=off -I"$ROOT/include" -I. "$@" -Xptxas -v -shared \
  -o "$ROOT/raycastworlds.jl_b200/lib/variants/librcw_b200_$name.so" rcw_kernels.cu rcw_capi.cu 2>&1 \
  | grep -A2 "frame_kernelILi0ELi0" | grep registers | sed "s/^/$name: /"
