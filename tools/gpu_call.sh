#!/bin/bash
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "attention" 2>&1 | tail -3
timeout 300 python tools/probe_layer.py 2>&1 | sed -n 3,3p
MA3_ATTN_GRID=0 timeout 300 python tools/probe_layer.py 2>&1 | sed -n 3,3p
