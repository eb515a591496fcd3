#!/bin/bash
# Build an experimental variant of the library next to the product one:  tools/build_variant.sh <name> [-DKNOB=value ...]
# -> gym_simpletetris_b200/libst_<name>.so (git-ignored; select it with ST_B200_LIB=<path>)
set -e
name=$1; shift
cd "$(dirname "$0")/.."
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --use_fast_math \
  -Xcompiler -fPIC,-fvisibility=hidden -shared -cudart static "$@" \
  -o gym_simpletetris_b200/libst_$name.so gym_simpletetris_b200/csrc/st_kernels.cu gym_simpletetris_b200/csrc/st_abi.cu
echo built gym_simpletetris_b200/libst_$name.so
