# scratch: the command list of the next gpurun call (rewritten per call)
mkdir -p gpurun_out
NB3=$PWD/make-an-audio-3_b200/csrc/build/libma3b200_nb3.so
MA3_LIB=$NB3 MA3_ACT_INPLACE=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "act1d" -x 2>&1 | tail -2
( echo "== product"; python tools/probe_act1d.py | grep "v0"
echo "== in place, 2 buffers"; MA3_ACT_INPLACE=1 python tools/probe_act1d.py | grep "v0"
echo "== in place, 3 buffers"; MA3_LIB=$NB3 MA3_ACT_INPLACE=1 python tools/probe_act1d.py | grep "v0" ) 2>&1 | tee gpurun_out/r02z_act1d_nbuf.log
