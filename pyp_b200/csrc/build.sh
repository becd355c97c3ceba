#!/bin/bash
# Build libcspb200.so in-tree for sm_100a (the .so travels to the GPU box with the snapshot).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../libcspb200.so
SRCS="capi.cu plan.cu fft.cu refine.cu search.cu recon.cu csp.cu pipeline.cu"
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
  -Xcompiler -fPIC,-Wall,-Wno-unused-function -shared $EXTRA_NVCC_FLAGS \
  -o $OUT $SRCS -L/usr/local/cuda/lib64 -Xlinker -rpath=/usr/local/cuda/lib64 -lcufft -lcudart
echo "built $OUT"
