# scratch: the command list of the next gpurun call (rewritten per call)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3 > gpurun_out/r02zz_tests_gpu.log; cat gpurun_out/r02zz_tests_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-lib-baseline > gpurun_out/r02zz_bench_default.json 2> gpurun_out/r02zz_bench_default.err; python -c "
import json; d=json.loads(open('gpurun_out/r02zz_bench_default.json').read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['ms_per_step'],2), d['e2e']['value'], d['stage_ms'], d['clocks']['sm_mhz'], d['gpu_launches'])
for k,v in d['gemm_shapes'].items():
    if 'taps3' in k: print(k, v)"
