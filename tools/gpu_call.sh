mkdir -p gpurun_out
run() { timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ab_$1.json 2> gpurun_out/bench_ab_$1.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_ab_$1.json').read().strip().splitlines()[-1])
print('$1', round(d['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], 'tap_gemm', d['kernel_breakdown']['tap_gemm']['ms'])
"; }
run base1
MA3_TILE_3=256,2 run qkv256c2
MA3_TILE_3=192,2 run qkv192c2
MA3_TILE_3=128,2 run qkv128c2
MA3_TILE_2=256,2 run w13_256c2
MA3_TILE_2=192,2 run w13_192c2
MA3_TILE_1=192,1 run wo192c1
MA3_TILE_1=192,2 run wo192c2
MA3_TILE_1=128,2 run wo128c2
MA3_TILE_4=192,1 run w2_192c1
MA3_TILE_4=192,2 run w2_192c2
MA3_TILE_4=256,2 run w2_256c2
run base2
