# scratch: the command list of the next gpurun call (rewritten per call)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3 > gpurun_out/r02z_tests_gpu.log; cat gpurun_out/r02z_tests_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r02z_bench_default.json 2> gpurun_out/r02z_bench_default.err; tail -c 300 gpurun_out/r02z_bench_default.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02z_bench_reference.json 2> gpurun_out/r02z_bench_reference.err; tail -c 200 gpurun_out/r02z_bench_reference.json
for c in 1 3 5; do timeout 600 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-lib-baseline > gpurun_out/r02z_bench_config$c.json 2> gpurun_out/r02z_bench_config$c.err; tail -c 200 gpurun_out/r02z_bench_config$c.json; done
