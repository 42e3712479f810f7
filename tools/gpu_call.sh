mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/tests_gpu_r01e.log; cat gpurun_out/tests_gpu_r01e.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_default_r01e.json 2> gpurun_out/bench_default_r01e.err; tail -c 600 gpurun_out/bench_default_r01e.json | head -c 600; echo
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference_r01e.json 2> gpurun_out/bench_reference_r01e.err; cat gpurun_out/bench_reference_r01e.json | head -c 900
