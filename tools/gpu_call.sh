# scratch: the command list of the next gpurun call (rewritten per call)
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-lib-baseline > gpurun_out/r02z_bench_default.json 2> gpurun_out/r02z_bench_default.err; tail -c 200 gpurun_out/r02z_bench_default.json
timeout 600 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline --no-lib-baseline > gpurun_out/r02z_bench_config4.json 2> gpurun_out/r02z_bench_config4.err; tail -c 200 gpurun_out/r02z_bench_config4.json
