#!/bin/bash
# Build libcspb200.so in-tree for sm_100a (the .so travels to the GPU box with the snapshot).
# Every translation unit is compiled in its own nvcc process (in parallel), then linked.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=${OUT:-../libcspb200.so}
SRCS="capi plan fft refine search recon csp pipeline select"
OBJ=$(mktemp -d)
trap 'rm -rf "$OBJ"' EXIT
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function $EXTRA_NVCC_FLAGS"
pids=()
for s in $SRCS; do
  $NVCC $FLAGS -c $s.cu -o "$OBJ/$s.o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT $(for s in $SRCS; do echo "$OBJ/$s.o"; done) \
  -L/usr/local/cuda/lib64 -Xlinker -rpath=/usr/local/cuda/lib64 -lcufft -lcudart
echo "built $OUT"
