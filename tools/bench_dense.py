"""Dense reduced-camera solve in isolation: own DMMA Cholesky vs the cuSOLVER yard-stick."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stba  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=5988)
    ap.add_argument("--reps", type=int, default=6)
    ap.add_argument("--backends", default="own,cusolver")
    a = ap.parse_args()
    rng = np.random.default_rng(0)
    n = a.n
    A = rng.normal(size=(n, 64))
    S = A @ A.T
    S[np.diag_indices(n)] += 64.0 + rng.uniform(0, 1, n)
    rhs = rng.normal(size=n)
    ref = None
    for be in a.backends.split(","):
        x, info, ms = stba.engine.dense_cholesky_solve(np.tril(S), rhs, {"own": stba.capi.DENSE_OWN, "cusolver": stba.capi.DENSE_CUSOLVER, "hybrid": stba.capi.DENSE_HYBRID}[be], reps=a.reps)
        res = np.linalg.norm(S @ x - rhs) / np.linalg.norm(rhs)
        gf = (n ** 3 / 3 + 2 * n * n) / 1e9
        best = ms[1:].min() if len(ms) > 1 else ms[0]
        print("%-9s n=%d info=%d  ms=%s  best %.3f ms = %.2f TFLOP/s  residual %.2e" % (be, n, info, np.round(ms, 3).tolist(), best, gf / best, res))
