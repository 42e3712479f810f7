#!/bin/bash
python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "gemm or qkv or conv" 2>&1 | tail -3
timeout 600 python tools/probe_tiles.py 2>&1 | grep -E "auto"
