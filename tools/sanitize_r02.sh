#!/bin/bash
# compute-sanitizer passes over the kernel parity tests (run under gpurun, 1 GPU).  Logs go to gpurun_out/ and are copied
# to profiles/ (memcheck / racecheck / synccheck as the stand-in for the race detection the reference
# does not have).  Each pass is bounded by its own timeout; a pass that times out is reported as such.
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
run() {  # tool, timeout, pytest -k expression
  local tool=$1 to=$2 expr=$3
  echo "== $tool: pytest tests/test_kernels_gpu.py -k \"$expr\"" > gpurun_out/r02_sanitizer_$tool.log
  timeout $to $SAN --tool $tool --error-exitcode 9 --print-limit 20 \
      python -m pytest tests/test_kernels_gpu.py -q -x -m gpu -k "$expr" >> gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "exit code $? (124 = timed out after ${to}s, 9 = sanitizer errors)" >> gpurun_out/r02_sanitizer_$tool.log
  grep -E "ERROR SUMMARY|passed|failed|exit code|Error|error" gpurun_out/r02_sanitizer_$tool.log | sort | uniq -c | head -12
}
run memcheck 600 "attention or rownorm or gate_residual or swiglu or act1d or rmsnorm or final_layer or conv"
run synccheck 420 "attention or rownorm or gate_residual or act1d"
run racecheck 420 "act1d or rmsnorm or final_layer or groupnorm"
