# scratch: the command list of the next gpurun call (rewritten per call)
mkdir -p gpurun_out
for v in 0 1 0 1; do MA3_PDL_TAIL=$v timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-lib-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('PDL_TAIL=$v', round(d['value'],1), round(d['ms_per_step'],2), d['stage_ms'], d['clocks']['sm_mhz'])"; done | tee gpurun_out/r02z_pdl_tail.log
