#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -x -q -m gpu > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?" >> gpurun_out/t_kernels.log
tail -3 gpurun_out/t_kernels.log
timeout 300 python tools/probe_layer.py > gpurun_out/probe_layer.log 2>&1; cat gpurun_out/probe_layer.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err; tail -3 gpurun_out/bench_r1b.err; cat gpurun_out/bench_r1b.json
