# scratch: the command list of the next gpurun call (rewritten per call)
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline | tail -c 400
