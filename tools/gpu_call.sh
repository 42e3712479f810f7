#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1g.json 2> gpurun_out/bench_r1g.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_r1g.json').read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['ms_per_step'],2), d['stage_ms'], {k:(v['ms'],v['launches']) for k,v in d['kernel_breakdown'].items() if v['ms']>0.3})"
