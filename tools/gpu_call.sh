#!/bin/bash
python -m pytest tests/test_kernels_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python tools/probe_tiles.py 2>&1 | grep -E "auto|tile_n=192 cta_group=2|tile_n=256 cta_group=2"
