#!/usr/bin/env python
"""Known-answer vectors produced by the REFERENCE'S OWN CODE, run in this container.

examples/replay_ref (tools/make_ref_replay.py) compiles st17-ceres/src/include/solver.hpp and
st20-g2o/src/include/test_ceres.h from /root/reference unmodified (with the Eigen / Sophus stand-ins of
include/compat and the ceres::Jet of include/ceres/jet.h).  Its `kat` and `gn` modes evaluate, on the host:
  ProjectFactor + Jet Jacobians, LieLocalParameterization<SO3d>, Triangulation, PnPSizedCostFunction,
  LieR3LocalParameterization, and the complete hand Gauss-Newton SelfGaussNewton (solver.hpp:387-462).
The outputs are committed as tests/golden/ref_kat.npz; tests/test_oracle_kat.py holds oracle/ against them, so the
oracle's residual, Jacobian, manifold and Gauss-Newton arithmetic is pinned to reference code and not only to itself.
(The trust-region control flow lives in Ceres, which is not in /root/reference: that part stays unpinned.)

  python tests/golden/make_ref_kat.py        # needs /root/reference; rewrites ref_kat.npz
"""
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import stba
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "make_ref_replay.py")])
    exe = os.path.join(ROOT, "examples", "replay_ref")
    rng = np.random.default_rng(20221109)
    n = 64
    q = rng.normal(size=(n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    t = rng.normal(0, 2.0, (n, 3))
    # points in front of the camera: P = t + R (x, y, z), z in [1.5, 8]
    from oracle import lie          # generator script = test infrastructure
    pc = np.stack([rng.uniform(-1.5, 1.5, n), rng.uniform(-1.0, 1.0, n), rng.uniform(1.5, 8.0, n)], axis=-1)
    P = np.array([ti + lie.quat_to_rot(qi) @ p for qi, ti, p in zip(q, t, pc)])
    uv = pc[:, :2] / pc[:, 2:3] + rng.normal(0, 0.05, (n, 2))
    delta = rng.normal(0, 0.1, (n, 3))
    delta[:4] *= 1e-12          # the small-angle branch of exp
    cases = np.concatenate([q, t, P, uv, delta], axis=1)
    with tempfile.TemporaryDirectory() as d:
        fin, fout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        with open(fin, "wb") as f:
            f.write(struct.pack("ii", n, 0))
            f.write(np.ascontiguousarray(cases).tobytes())
        subprocess.check_call([exe, "kat", fin, fout])
        out = np.fromfile(fout).reshape(n, 85)
        s = stba.synth.pnp_scene()
        with open(fin, "wb") as f:
            f.write(struct.pack("i", len(s["points"])))
            for a in (s["points"], s["uv"], s["q_init"], s["t_init"]):
                f.write(np.ascontiguousarray(a, np.float64).tobytes())
        r = subprocess.run([exe, "gn", fin, fout], capture_output=True, text=True, check=True)
        gn_pose = np.fromfile(fout)
        gn_iters = int(r.stdout.split("iter num:")[1].split()[0])
    np.savez(os.path.join(ROOT, "tests", "golden", "ref_kat.npz"), cases=cases, out=out, gn_pose=gn_pose, gn_iter_num=gn_iters)
    print("wrote tests/golden/ref_kat.npz:", cases.shape, out.shape, gn_pose, gn_iters)


if __name__ == "__main__":
    main()
