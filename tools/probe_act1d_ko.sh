#!/bin/bash
# Knock-out builds of the Activation1d kernel (diagnostics): libma3b200_ko{1,2,3}.so next to the product library, timed
# with tools/probe_act1d.py through MA3_LIB.  Results of a knock-out build are wrong by construction.
cd "$(dirname "$0")/../make-an-audio-3_b200/csrc" || exit 1
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I../../include"
for k in 1 2 3; do
  $NVCC $FLAGS -DMA3_ACT_KO=$k -c -o build/act1d_ko$k.o act1d.cu &
done
wait
for k in 1 2 3; do
  $NVCC -gencode arch=compute_100a,code=sm_100a -shared -o build/libma3b200_ko$k.so build/host_common.o build/gemm.o build/rowgemm.o build/elementwise.o build/melnet.o build/act1d_ko$k.o build/attention.o
done
ls -la build/*.so
