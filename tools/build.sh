#!/bin/bash
# Build libma3b200.so for sm_100a (in-tree; the .so travels to the GPU box with the snapshot).
set -e
cd "$(dirname "$0")/../make-an-audio-3_b200/csrc"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I../../include"
mkdir -p build
pids=()
for f in host_common gemm elementwise act1d attention; do
  if [ ! -f build/$f.o ] || [ $f.cu -nt build/$f.o ] || [ ptx.cuh -nt build/$f.o ] || [ host_common.h -nt build/$f.o ] || [ ../../include/ma3_b200.h -nt build/$f.o ]; then
    $NVCC $FLAGS -c -o build/$f.o $f.cu &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o libma3b200.so build/*.o
echo "built $(pwd)/libma3b200.so"
