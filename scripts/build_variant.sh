#!/bin/bash
# Builds a kernel-variant copy of libeskf_gpu.so for A/B runs on the GPU box:
#   scripts/build_variant.sh NAME -DESKF_ALIGN_FOLD=1 -DESKF_ALIGN_MINB=2 ...
# -> eskf_lio_b200/lib/variants/NAME.so ; select it with ESKF_GPU_LIB=<path>.
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p eskf_lio_b200/lib/variants
PATH=/usr/bin:$PATH /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
  -Xcompiler -fPIC -shared "$@" -o eskf_lio_b200/lib/variants/$name.so eskf_lio_b200/csrc/*.cu
echo built eskf_lio_b200/lib/variants/$name.so
