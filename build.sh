#!/bin/bash
# Build libffvd_b200.so for sm_100a (in-tree, so it travels to the GPU box with gpurun).
set -e
cd "$(dirname "$0")"
make -j"$(nproc)" "$@"
