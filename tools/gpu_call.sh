mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/tests_gpu_r01g.log; cat gpurun_out/tests_gpu_r01g.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_default_r01g.json 2> gpurun_out/bench_default_r01g.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_default_r01g.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],2), d['clocks'], round(d['roofline']['frac'],4), round(d['e2e']['value'],1), d['cpu_baseline']['value'], d['stage_ms'])
"
for m in "M 16" "XXL 1"; do set -- $m; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --model $1 --prompts $2 > gpurun_out/bench_r01g_$1_$2.json 2>/dev/null; python -c "
import json
d=json.loads(open('gpurun_out/bench_r01g_$1_$2.json').read().strip().splitlines()[-1])
print('$1 $2', round(d['value'],1), round(d['ms_per_step'],2), d['stage_ms'])
"; done
