#!/bin/bash
# Builds libextfem_cuda.so in-tree for sm_100a (travels to the GPU box with the snapshot).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --extended-lambda \
      -split-compile 0 -Xcompiler -fPIC -shared -ldl ${EXTFEM_NVCC_EXTRA} -o libextfem_cuda.so extfem.cu
