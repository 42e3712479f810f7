# scratch: the command list of the next gpurun call (rewritten per call)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "narrow or conv" -x 2>&1 | tail -8 > gpurun_out/r02z_tests_conv.log; cat gpurun_out/r02z_tests_conv.log
timeout 600 python -m pytest tests/test_modules_gpu.py tests/test_parity_full_size_gpu.py -q -m gpu -k "bigvgan or vocode" -x 2>&1 | tail -3
timeout 300 python tools/probe_vocoder.py > gpurun_out/r02z_probe_vocoder.log 2>&1; grep "conv" gpurun_out/r02z_probe_vocoder.log | grep "C96\|C192"
MA3_CONV_NARROW_WIDE=0 timeout 300 python tools/probe_vocoder.py > gpurun_out/r02z_probe_vocoder_generic.log 2>&1; grep "conv" gpurun_out/r02z_probe_vocoder_generic.log | grep "C96"
