# scratch: the command list of the next gpurun call (rewritten per call)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "gemm or conv" -x 2>&1 | tail -3
timeout 600 python -m pytest tests/test_modules_gpu.py tests/test_parity_full_size_gpu.py -q -m gpu -k "bigvgan or vocode or vae" -x 2>&1 | tail -3
timeout 300 python tools/probe_vocoder.py > gpurun_out/r02z2_probe_vocoder.log 2>&1; grep "acc" gpurun_out/r02z2_probe_vocoder.log
