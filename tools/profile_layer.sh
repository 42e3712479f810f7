#!/bin/bash
# source-level ncu captures of the DiT block kernels at bench shapes (tools/probe_layer.py XL x)
CMD="python tools/probe_layer.py XL x"
$CMD > gpurun_out/plain_layer.log 2>&1 || { tail gpurun_out/plain_layer.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:tap_gemm -s 2 -c 1 -o gpurun_out/prof_qkv $CMD > gpurun_out/ncu_qkv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tap_gemm -s 14 -c 1 -o gpurun_out/prof_wo $CMD > gpurun_out/ncu_wo.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_kernel -s 2 -c 1 -o gpurun_out/prof_attn2 $CMD > gpurun_out/ncu_attn2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rmsnorm -s 2 -c 1 -o gpurun_out/prof_rms2 $CMD > gpurun_out/ncu_rms2.log 2>&1
ls -la gpurun_out/*.ncu-rep
