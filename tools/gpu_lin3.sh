#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/time_lin.py 10 2>&1 | tee gpurun_out/time_lin3.log
STBA_LIN2=1 timeout 300 python tools/time_lin.py 10 2>&1 | tee gpurun_out/time_lin2.log
