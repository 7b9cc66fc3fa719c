#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_cpp_shim.py tests/test_front.py tests/test_g2o_front.py -m gpu -x -q 2>&1 | tail -15
