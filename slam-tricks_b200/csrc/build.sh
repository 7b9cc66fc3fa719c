#!/bin/bash
# Builds libstba.so in-tree for sm_100a (the only target).  nvcc cross-compiles without a GPU.
# NCCL: link the copy PyTorch ships (nvidia/nccl, 2.28.x) so that one process never holds two
# different libnccl.so.2 — torch's libtorch_cuda needs symbols the system 2.27 lacks.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
PY=${PYTHON:-python}
NCCL_DIR=$($PY - <<'PYEOF'
import importlib.util, os
spec = importlib.util.find_spec("nvidia.nccl")
print(os.path.dirname(spec.origin) if spec and spec.origin else list(spec.submodule_search_locations)[0] if spec else "")
PYEOF
)
if [ -n "$NCCL_DIR" ] && [ -f "$NCCL_DIR/lib/libnccl.so.2" ]; then
  NCCL_FLAGS="-I$NCCL_DIR/include -Xlinker $NCCL_DIR/lib/libnccl.so.2 -Xlinker -rpath,$NCCL_DIR/lib"
else
  NCCL_FLAGS="-lnccl"
fi
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function --expt-relaxed-constexpr"
OUT=${OUT:-../libstba.so}
$NVCC $FLAGS -shared -o $OUT stba_engine.cu stba_chol.cu stba_problem.cu stba_front.cu stba_pnp.cu stba_calib.cu stba_posegraph.cu -I../../include \
  $NCCL_FLAGS -L/usr/local/cuda/lib64 -lcusolver -lcublas -Xlinker -rpath,/usr/local/cuda/lib64 "$@"
echo "built $(readlink -f $OUT)"
