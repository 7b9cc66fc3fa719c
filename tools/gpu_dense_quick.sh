#!/bin/bash
set -u
mkdir -p gpurun_out
for cfg in "2 1" "4 1" "8 1" "12 1" "6 2" "8 2"; do set -- $cfg
echo "W=$1 G=$2: $(STBA_CHOL_WINDOW=$1 STBA_CHOL_AGG=$2 timeout 300 python tools/bench_dense.py --backends own --reps 6 2>&1 | tail -1)"
done
STBA_CHOL_WINDOW=8 STBA_CHOL_AGG=1 STBA_CHOL_PROF=1 timeout 300 python tools/bench_dense.py --backends own --reps 3 > gpurun_out/dense_dag.log 2>&1
tail -22 gpurun_out/dense_dag.log | cut -c1-330 | grep -v "trsm clocks\|potrf128 clocks"
