#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "conv or gemm" 2>&1 | tail -3
timeout 300 python tools/probe_vocoder.py 2>&1 | grep -E "conv.*(C48|C32)"
timeout 600 python -m pytest tests/test_modules_gpu.py -x -q -m gpu -k "bigvgan or pipeline or vae" 2>&1 | tail -3
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1k.json 2> gpurun_out/bench_r1k.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_r1k.json').read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['ms_per_step'],2), d['stage_ms'], round(d['roofline']['frac'],3))"
