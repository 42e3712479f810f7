#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err; tail -3 gpurun_out/bench_r1d.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_r1d.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['stage_ms'], d['roofline']['frac'], {k:(v['ms'],v.get('achieved_GBps')) for k,v in d['kernel_breakdown'].items() if v['ms']>1})"
