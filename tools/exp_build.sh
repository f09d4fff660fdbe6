#!/bin/bash
# tools/exp_build.sh NAME [-DFLAG ...] : build a tuning variant of the CUDA library into exp/NAME.so
# (git-ignored; travels to the GPU box; load it with ZW_LIB_PATH=exp/NAME.so)
set -e
cd "$(dirname "$0")/../image_webp_b200/csrc"
name=$1; shift
mkdir -p ../../exp
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-ffp-contract=off -shared "$@" -o ../../exp/$name.so zw_capi.cu
echo built exp/$name.so
