mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_modules_gpu.py -x -q -m gpu -k "swiglu or dit or sampler or pipeline" 2>&1 | tail -2
run() { timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ab_$1.json 2> gpurun_out/bench_ab_$1.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_ab_$1.json').read().strip().splitlines()[-1])
q=[v for k,v in d['gemm_shapes'].items() if k.startswith('swiglu')][0]
print('$1', round(d['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], 'tap_gemm', d['kernel_breakdown']['tap_gemm']['ms'], 'w13', q)
"; }
MA3_LIB=$PWD/make-an-audio-3_b200/csrc/libma3b200_lean.so run old1
run new1
MA3_LIB=$PWD/make-an-audio-3_b200/csrc/libma3b200_lean.so run old2
run new2
