"""Landmark sharding for the multi-GPU path (host logic).

Observations are landmark-major, so a contiguous landmark range owns a contiguous observation
slice.  Cameras are replicated on every rank; `obs_cam` stays global, `obs_lm` becomes local."""
import numpy as np


def landmark_ranges(obs_lm, n_lm, nranks):
    """Contiguous landmark ranges balanced by observation count.  Returns int64[nranks+1] bounds.
    Deterministic and identical on every rank (depends only on the degree histogram)."""
    deg = np.bincount(np.asarray(obs_lm, dtype=np.int64), minlength=n_lm)
    csum = np.concatenate([[0], np.cumsum(deg)])
    total = csum[-1]
    bounds = [0]
    for r in range(1, nranks):
        target = total * r / nranks
        b = int(np.searchsorted(csum, target, side="left"))
        bounds.append(min(max(b, bounds[-1]), n_lm))
    bounds.append(n_lm)
    return np.asarray(bounds, dtype=np.int64)


def shard_scene(lm, obs_cam, obs_lm, obs_uv, rank, nranks, lm_const=None):
    """This rank's landmarks and observations: (lm, obs_cam, obs_lm_local, obs_uv, lm_const, (lo, hi))."""
    obs_lm = np.asarray(obs_lm)
    n_lm = len(lm)
    b = landmark_ranges(obs_lm, n_lm, nranks)
    lo, hi = int(b[rank]), int(b[rank + 1])
    o0 = int(np.searchsorted(obs_lm, lo, side="left"))
    o1 = int(np.searchsorted(obs_lm, hi, side="left"))
    lc = None if lm_const is None else np.asarray(lm_const)[lo:hi]
    return (np.asarray(lm)[lo:hi], np.asarray(obs_cam)[o0:o1], (obs_lm[o0:o1] - lo).astype(np.int32),
            np.asarray(obs_uv)[o0:o1], lc, (lo, hi))
