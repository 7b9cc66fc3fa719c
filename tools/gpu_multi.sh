#!/bin/bash
# N-GPU box: the multi-rank parity tests, then the bench line (both arms) at N ranks
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench N=$N rc=$?"; tail -3 gpurun_out/bench_n$N.err
python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/bench_n$N.json').read().strip().splitlines() if l.startswith('{')][-1])
    print({k: d.get(k) for k in ('value', 'n_gpus', 'ms_per_step', 'gpu_launches', 'final_cost')}, 'e2e', {k: d['e2e'][k] for k in ('value','steps','ms_per_step')}, d.get('phase_ms_per_solve'))
except Exception as e:
    print('parse failed', e)
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2> gpurun_out/ref_n$N.err | cut -c1-300
