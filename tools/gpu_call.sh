# scratch: the command list of the next gpurun call (rewritten per call)
mkdir -p gpurun_out
python tools/probe_act1d_one.py 384 9984 8 || exit 1
ncu --set full --clock-control none --import-source on -k regex:act1d_tma -s 2 -c 1 -f -o gpurun_out/r02x_prof_act1d python tools/probe_act1d_one.py 384 9984 8 > gpurun_out/r02x_ncu_act1d.log 2>&1; tail -2 gpurun_out/r02x_ncu_act1d.log
