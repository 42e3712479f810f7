#!/bin/bash
timeout 900 python -m pytest tests/test_modules_gpu.py -x -q -m gpu -k "full_size_properties" 2>&1 | tail -12
