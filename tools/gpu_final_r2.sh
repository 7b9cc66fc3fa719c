#!/bin/bash
# Final evidence of round 2: smoke(), GPU suite, bench lines of all four workloads, launch list of the bench command.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r2_pytest_gpu.log
for w in C B PG CALIB; do
  timeout 900 python bench.py --workload $w > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err; echo "bench $w rc=$?"; tail -1 gpurun_out/r2_bench_$w.err
done
python tools/show_bench.py gpurun_out/r2_bench_C.json | head -12
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --scaled 0 > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
