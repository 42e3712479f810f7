#!/bin/bash
mkdir -p gpurun_out
for cfg in "M 1" "M 16" "XXL 1" "MOE 1"; do set -- $cfg; timeout 900 python bench.py --model $1 --prompts $2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$1_$2.json 2> gpurun_out/bench_$1_$2.err || tail -5 gpurun_out/bench_$1_$2.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_$1_$2.json').read().strip().splitlines()[-1]); print('$1 $2', round(d['value'],1), round(d['ms_per_step'],2), d['stage_ms'], round(d['roofline']['frac'],3))"; done
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1]); print('XL 8', round(d['value'],1), round(d['ms_per_step'],2), d['stage_ms'], round(d['roofline']['frac'],3), d['cpu_baseline'], d['clocks'])"
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 2>/dev/null | tail -1 | cut -c1-600
