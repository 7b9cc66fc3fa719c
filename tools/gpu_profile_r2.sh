#!/bin/bash
# Round-2 ncu evidence: launch list of the bench command, --set full of the DAG Cholesky, the linearisation and the
# Schur kernels.  Numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/profile_r2.log; }
stamp "ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --scaled 0 > gpurun_out/bench_under_ncu.log 2>&1; stamp "launch list rc=$?"
stamp "ncu full dense (k_chol_dag2)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_chol_dag2' -s 2 -c 1 -f -o gpurun_out/r2_dense_full \
    python tools/profile_solve.py --solves 1 > gpurun_out/ncu_dense.log 2>&1; stamp "ncu dense rc=$?"
stamp "ncu full lin"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_lin' -s 4 -c 4 -f -o gpurun_out/r2_lin_full \
    python tools/profile_lin.py > gpurun_out/ncu_lin.log 2>&1; stamp "ncu lin rc=$?"
stamp "ncu full schur/backsub"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_schur|k_backsub' -s 6 -c 6 -f -o gpurun_out/r2_schur_full \
    python tools/profile_solve.py --solves 1 > gpurun_out/ncu_schur.log 2>&1; stamp "ncu schur rc=$?"
stamp done
