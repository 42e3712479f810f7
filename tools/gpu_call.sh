#!/bin/bash
mkdir -p gpurun_out
MA3_RMS_BULK=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "rmsnorm" 2>&1 | tail -3
timeout 300 python tools/probe_layer.py 2>&1 | sed -n 1,1p
MA3_RMS_BULK=1 timeout 300 python tools/probe_layer.py 2>&1 | sed -n 1,1p
MA3_RMS_BULK=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1m.json 2> gpurun_out/bench_r1m.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_r1m.json').read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['ms_per_step'],2), d['stage_ms'], {k:(v['ms'],v['launches']) for k,v in d['kernel_breakdown'].items() if v['ms']>5})"
