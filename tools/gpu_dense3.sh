#!/bin/bash
# Stress of the own dense solve: repeated parity tests, then timings.
set -u
mkdir -p gpurun_out
for r in 1 2 3; do
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dense_cholesky" 2>&1 | tail -2
done
timeout 300 python tools/bench_dense.py --backends own,own,own,hybrid --reps 10 2>&1 | tail -4
STBA_CHOL_PROF=1 timeout 300 python tools/bench_dense.py --backends own --reps 3 > gpurun_out/dense_dag.log 2>&1
tail -19 gpurun_out/dense_dag.log | cut -c1-400
