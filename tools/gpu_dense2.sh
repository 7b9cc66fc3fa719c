#!/bin/bash
# Quick dense check: parity tests of the own factorisation, DAG profile (trace of the chain CTAs), own vs hybrid.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dense_cholesky" > gpurun_out/pytest_dense.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_dense.log
STBA_CHOL_PROF=1 timeout 300 python tools/bench_dense.py --backends own --reps 3 > gpurun_out/dense_dag.log 2>&1
echo "dag rc=$?"; tail -19 gpurun_out/dense_dag.log | cut -c1-400
timeout 300 python tools/bench_dense.py --backends own,hybrid --reps 8 > gpurun_out/dense_cmp.log 2>&1
echo "cmp rc=$?"; cat gpurun_out/dense_cmp.log
