#!/bin/bash
# Builds the working tree's CUDA library under another name for same-box A/B runs:
#   profiles/build_variant.sh <tag> [nvcc defines...]   ->  picsp_b200/variants/libpicsp_b200_<tag>.so
# (git-ignored, travels to the GPU box; select it with PICSP_B200_LIB=<path> — bench.py / tests only.)
set -e
cd "$(dirname "$0")/.."
tag=$1; shift
mkdir -p picsp_b200/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function -ccbin g++ \
    -I include "$@" -shared -o picsp_b200/variants/libpicsp_b200_$tag.so picsp_b200/csrc/abi.cu \
    $(ls picsp_b200/csrc/host/*.cpp | grep -v main.cpp) -lcufft -ldl -Xlinker -rpath,/usr/local/cuda/lib64
echo picsp_b200/variants/libpicsp_b200_$tag.so
