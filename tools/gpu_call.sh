mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/tests_gpu_r01f.log; cat gpurun_out/tests_gpu_r01f.log
timeout 900 python bench.py > gpurun_out/bench_default_r01f.json 2> gpurun_out/bench_default_r01f.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_default_r01f.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],2), d['clocks'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'])
for k,v in list(d['gemm_shapes'].items())[:4]: print(k, v)
"
for m in "M 1" "M 16" "XXL 1" "MOE 1"; do set -- $m; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --model $1 --prompts $2 > gpurun_out/bench_r01f_$1_$2.json 2>/dev/null; python -c "
import json
d=json.loads(open('gpurun_out/bench_r01f_$1_$2.json').read().strip().splitlines()[-1])
print('$1 $2', round(d['value'],1), round(d['ms_per_step'],2), d['stage_ms'])
"; done
