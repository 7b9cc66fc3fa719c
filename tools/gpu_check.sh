#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for w in C PG CALIB; do
timeout 900 python bench.py --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
echo "bench $w rc=$?"; tail -3 gpurun_out/bench_$w.err
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_$w.json').read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ('value', 'ms_per_step', 'gpu_launches', 'iterations_per_solve')}, 'e2e', {k: d['e2e'][k] for k in ('value','steps','ms_per_step')}, 'roof', {k: d['roofline'][k] for k in ('achieved','frac','launch_ms')}, 'cpu', d.get('cpu_baseline') and {k: d['cpu_baseline'].get(k) for k in ('value','cores','one_thread')})
    print(d.get('phase_ms_isolated'))
except Exception as e:
    print('parse failed', e)
PY
done
