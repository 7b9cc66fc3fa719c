#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/time_lin.py 10 2>&1 | tee gpurun_out/time_lin3.log
STBA_LIB=/root/repo/slam-tricks_b200/libstba_timing.so timeout 200 python tools/l3_clocks.py 1
STBA_LIB=/root/repo/slam-tricks_b200/libstba_timing.so timeout 200 python tools/l3_clocks.py 10
