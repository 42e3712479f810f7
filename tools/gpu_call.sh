#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/tests_gpu.log 2>&1; tail -3 gpurun_out/tests_gpu.log
timeout 300 python tools/probe_vocoder.py 2>&1 | grep -E "conv.*(C192|C96).*res"
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1l.json 2> gpurun_out/bench_r1l.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_r1l.json').read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['ms_per_step'],2), d['stage_ms'], round(d['roofline']['frac'],3))"
