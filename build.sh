#!/usr/bin/env bash
# Build libsaev_b200.so in-tree for sm_100a (the .so travels to the GPU box with the repo snapshot).
set -euo pipefail
cd "$(dirname "$0")"
mkdir -p saev_b200/lib
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared ${NVCC_EXTRA:-} \
  saev_b200/csrc/encode_gemm.cu saev_b200/csrc/encode_gemm2.cu saev_b200/csrc/dense_gemm2.cu saev_b200/csrc/sparse_kernels.cu saev_b200/csrc/aux_kernels.cu saev_b200/csrc/dense_kernels.cu saev_b200/csrc/repair_kernels.cu saev_b200/csrc/batch_topk_kernels.cu saev_b200/csrc/api.cu saev_b200/csrc/shard_loader.cu \
  -o saev_b200/lib/libsaev_b200.so
