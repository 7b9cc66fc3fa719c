#!/bin/bash
# N-GPU box: bench line with the split dense solve (default for N > 1) and with the replicated one
set -u
N=${1:-4}
mkdir -p gpurun_out
for mode in split nosplit; do
  if [ $mode = nosplit ]; then export STBA_CHOL_NOSPLIT=1; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n${N}_$mode.json 2> gpurun_out/bench_n${N}_$mode.err
  echo "bench N=$N $mode rc=$?"
  python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/bench_n${N}_$mode.json').read().strip().splitlines() if l.startswith('{')][-1])
    print({k: d.get(k) for k in ('value', 'n_gpus', 'ms_per_step', 'final_cost')}, 'e2e', d['e2e']['value'], d.get('phase_ms_per_solve'))
except Exception as e:
    print('parse failed', e)
PY
done
