#!/bin/bash
# ncu evidence for round 1 (run under gpurun, 1 GPU).  Every ncu command is preceded by a plain run of the same command.
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain.log; exit 1; }
tail -c 600 gpurun_out/plain.log
# (1) launch list of one full step: skip the launches of the 3 warm-up steps (5089 launches per step, eager)
ncu --metrics gpu__time_duration.sum --clock-control none -s 15267 -c 5089 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
# (2) full-set captures: the four DiT GEMMs of one block, attention, the tensor-core Activation1d, rmsnorm, narrow conv
ncu --set full --clock-control none --import-source on -k regex:tap_gemm -s 400 -c 4 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn2_kernel -s 30 -c 1 -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_attn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:act1d_mma -s 40 -c 2 -o gpurun_out/prof_act1d $CMD > gpurun_out/ncu_act1d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rmsnorm -s 40 -c 1 -o gpurun_out/prof_rms $CMD > gpurun_out/ncu_rms.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_narrow -s 20 -c 2 -o gpurun_out/prof_convn $CMD > gpurun_out/ncu_convn.log 2>&1
ls -la gpurun_out/ | tail -15
