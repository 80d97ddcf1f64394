#!/bin/bash
# Builds libextfem_cuda.so in-tree for sm_100a (travels to the GPU box with the snapshot).
# EXTFEM_SPLIT=1 halves the build time (-split-compile 0) for development, but the partitioned optimisation changes the
# code of unrelated kernels from build to build (measured: +-17 % on the template kernels): release builds do not use it.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --extended-lambda \
      ${EXTFEM_SPLIT:+-split-compile 0} -Xcompiler -fPIC -shared -ldl ${EXTFEM_NVCC_EXTRA} -o libextfem_cuda.so extfem.cu
