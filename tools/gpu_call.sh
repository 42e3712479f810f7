#!/bin/bash
python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "act1d or attention" 2>&1 | tail -3
timeout 300 python tools/probe_vocoder.py act 2>&1 | tail -6
timeout 300 python tools/probe_layer.py 2>&1 | sed -n 3,3p
