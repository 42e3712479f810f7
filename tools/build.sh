#!/bin/bash
# Build libma3b200.so for sm_100a (in-tree; the .so travels to the GPU box with the snapshot).
cd "$(dirname "$0")/../make-an-audio-3_b200/csrc" || exit 1
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I../../include"
mkdir -p build
rm -f build/*.fail
# MA3_BUILD_FORCE=1 (set by __graft_entry__.build()): recompile every source, so "does it build" is really checked
if [ "${MA3_BUILD_FORCE:-0}" = "1" ]; then rm -f build/*.o; fi
for f in host_common gemm rowgemm elementwise melnet act1d attention; do
  if [ ! -f build/$f.o ] || [ $f.cu -nt build/$f.o ] || [ ptx.cuh -nt build/$f.o ] || [ host_common.h -nt build/$f.o ] || [ ../../include/ma3_b200.h -nt build/$f.o ]; then
    ( $NVCC $FLAGS -c -o build/$f.o.tmp $f.cu && mv build/$f.o.tmp build/$f.o || { rm -f build/$f.o build/$f.o.tmp; touch build/$f.fail; } ) &
  fi
done
wait
if ls build/*.fail >/dev/null 2>&1; then echo "BUILD FAILED: $(ls build/*.fail)"; rm -f libma3b200.so; exit 1; fi
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o libma3b200.so build/host_common.o build/gemm.o build/rowgemm.o build/elementwise.o build/melnet.o build/act1d.o build/attention.o || exit 1
echo "built $(pwd)/libma3b200.so"
