#!/bin/bash
# Build libffvd_b200.so for sm_100a (in-tree, so it travels to the GPU box with gpurun).
set -e
cd "$(dirname "$0")"
mkdir -p ffvd_b200/lib
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared \
     -o ffvd_b200/lib/libffvd_b200.so ffvd_b200/csrc/capi.cu "$@"
