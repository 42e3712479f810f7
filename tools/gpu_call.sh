#!/bin/bash
python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "act1d" 2>&1 | tail -3
timeout 300 python tools/probe_vocoder.py act 2>&1 | tail -6
