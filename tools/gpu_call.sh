#!/bin/bash
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "conv or gemm" 2>&1 | tail -3
timeout 300 python tools/probe_vocoder.py 2>&1 | grep -E "conv.*(C48|C32).*res"
