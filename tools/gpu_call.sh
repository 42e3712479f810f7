# scratch: the command list of the next gpurun call (rewritten per call)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "act1d" -x 2>&1 | tail -3 > gpurun_out/r02t_tests_act.log; cat gpurun_out/r02t_tests_act.log
timeout 300 python tools/probe_act1d.py > gpurun_out/r02t_probe_act1d.log 2>&1; grep v0 gpurun_out/r02t_probe_act1d.log
