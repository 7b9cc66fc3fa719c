#!/bin/bash
# Dense back end on the GPU box: parity tests of the own factorisation, then timings (DAG kernel with its
# per-CTA cycle profile, the round-1 graph schedule, the library yard-stick).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dense_cholesky" > gpurun_out/pytest_dense.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_dense.log
STBA_CHOL_PROF=1 timeout 300 python tools/bench_dense.py --backends own --reps 4 > gpurun_out/dense_dag.log 2>&1
echo "dag rc=$?"; tail -30 gpurun_out/dense_dag.log
timeout 300 python tools/bench_dense.py --backends own,hybrid --reps 8 > gpurun_out/dense_cmp.log 2>&1
echo "cmp rc=$?"; cat gpurun_out/dense_cmp.log
STBA_CHOL_GRAPH=1 timeout 300 python tools/bench_dense.py --backends own --reps 6 > gpurun_out/dense_graph.log 2>&1
echo "graph rc=$?"; cat gpurun_out/dense_graph.log
