#!/bin/bash
set -u
for cfg in "1 3" "2 2" "2 4" "2 6" "3 3" "4 3" "4 4" "6 4" "3 8"; do set -- $cfg
echo "W=$1 G=$2: $(STBA_CHOL_WINDOW=$1 STBA_CHOL_AGG=$2 timeout 300 python tools/bench_dense.py --backends own --reps 6 2>&1 | tail -1 | cut -c1-200)"
done
