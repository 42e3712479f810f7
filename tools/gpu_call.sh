mkdir -p gpurun_out
run() { timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ab_$1.json 2> gpurun_out/bench_ab_$1.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_ab_$1.json').read().strip().splitlines()[-1])
g=[(k.split('/')[1][:22],v['us_per_launch']) for k,v in d['gemm_shapes'].items() if k.startswith('gate_res')]
print('$1', round(d['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], 'tap_gemm', d['kernel_breakdown']['tap_gemm']['ms'], g)
"; }
run base1
MA3_TILE_1=176,1 run wo176
MA3_TILE_4=176,1 run w2_176
MA3_TILE_1=176,1 MA3_TILE_4=176,1 run both176
run base2
