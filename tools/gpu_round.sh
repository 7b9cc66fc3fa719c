#!/bin/bash
# One GPU-box pass: parity tests, bench line, ncu launch list of the bench command, ncu --set full of the
# linearisation / Schur kernels.  Everything lands in gpurun_out/ (copied into profiles/ by hand).
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/round.log; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
stamp "pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; stamp "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
stamp "bench"
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; stamp "bench rc=$?"
tail -c 3000 gpurun_out/bench.json
if [ "${SKIP_NCU:-0}" != "1" ]; then
stamp "ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --scaled 0 > gpurun_out/bench_under_ncu.log 2>&1; stamp "launch list rc=$?"
stamp "ncu full lin"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_lin' -s 4 -c 4 -f -o gpurun_out/lin_full \
    python tools/profile_lin.py > gpurun_out/ncu_lin.log 2>&1; stamp "ncu lin rc=$?"
if [ "${NCU_SCHUR:-0}" = "1" ]; then
stamp "ncu full schur/backsub"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_schur|k_backsub' -s 5 -c 5 -f -o gpurun_out/schur_full \
    python tools/profile_solve.py --solves 1 > gpurun_out/ncu_schur.log 2>&1; stamp "ncu schur rc=$?"
fi
fi
stamp done
