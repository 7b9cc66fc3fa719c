#!/bin/bash
# Round-2 (second half) evidence: GPU suite, bench C, ncu --set full of k_lin3, launch list of the bench command.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2b_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2b_bench_C.json 2> gpurun_out/r2b_bench_C.err; echo "bench rc=$?"; tail -2 gpurun_out/r2b_bench_C.err
python tools/show_bench.py gpurun_out/r2b_bench_C.json 2>/dev/null | head -40
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_lin3' -s 3 -c 2 -f -o gpurun_out/r2b_lin3_full \
    python tools/profile_lin.py > gpurun_out/ncu_lin3.log 2>&1; echo "ncu lin3 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2b_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --scaled 0 > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
