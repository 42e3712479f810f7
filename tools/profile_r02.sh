#!/bin/bash
# ncu evidence for round 2 (run under gpurun, 1 GPU).  Every ncu command is preceded by a plain run of the same command
# that exited 0.  Outputs go to gpurun_out/; tools/ncu_summarise.py turns the .ncu-rep files into the text summaries
# and profiles/traffic.json that are committed under profiles/.
set -x
T=${TAG:-r02w}   # file tag of this pass (r02 = mid-round capture, r02w = final build of round 2)
mkdir -p gpurun_out
# (0) the step itself, un-profiled
BENCH="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-lib-baseline"
$BENCH > gpurun_out/${T}_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/${T}_plain.log; exit 1; }
tail -c 400 gpurun_out/${T}_plain.log
# (1) launch list of ONE eager step of the benchmarked workload (skip the launches of the 3 warm-up steps)
N=$(TAG_=$T python - <<'PY'
import json, os
d = json.loads(open("gpurun_out/%s_plain.log" % os.environ["TAG_"]).read().strip().splitlines()[-1])
print(d["gpu_launches"])
PY
)
echo "launches per step: $N"
ncu --metrics gpu__time_duration.sum --clock-control none -s $((3 * N)) -c $N --csv \
    --log-file gpurun_out/${T}_ncu_launches_XL8.csv $BENCH > gpurun_out/${T}_ncu_launch.log 2>&1
# (2) full-set captures on a reduced-depth driver with the same shapes (XL width, 8 prompts, depth 2, eager)
STEP="python tools/repro_step.py 2 8 3"
$STEP > gpurun_out/${T}_plain_step.log 2>&1 || { echo "plain step failed"; tail gpurun_out/${T}_plain_step.log; exit 1; }
cap() {  # name, kernel regex, skip, count
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 -o gpurun_out/${T}_prof_$1 $STEP > gpurun_out/${T}_ncu_$1.log 2>&1
}
cap attn attn2_kernel 4 1
cap rownorm rowgemm_norm 8 2          # wo (K=1152) and w2 (K=3072) of one block
cap gemm_qkv 'tap_gemm_kernel<\(int\)3, \(int\)2' 2 1   # QKV+RoPE of one block (EPI 3), third launch of that kind
cap gemm_swiglu 'tap_gemm_kernel<\(int\)2' 2 1   # w1|w3+SwiGLU of one block (EPI 2)
cap rms rmsnorm_modulate 2 1
# vocoder kernels from the bench driver (graph-free)
ncu --set full --clock-control none --import-source on -k regex:act1d_tma -s 40 -c 2 -o gpurun_out/${T}_prof_act1d $BENCH > gpurun_out/${T}_ncu_act1d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_narrow -s 20 -c 2 -o gpurun_out/${T}_prof_convn $BENCH > gpurun_out/${T}_ncu_convn.log 2>&1
ls -la gpurun_out/*.ncu-rep
