"""Generates the pose-graph fixtures (run in the build container, where /root/reference exists):
    python tests/golden/make_posegraph_golden.py
* posegraph_st4.npz   — the reference's recorded tracks st4-kalman/output/truth.csv and obs.csv (1000 SE(3) poses each,
                        `x,y,z,qx,qy,qz,qw`, written by st4-kalman/src/main.cpp:7-29 from `simulation`,
                        pose_simulation.cpp:17-88) + the pose graph built on them: edges (i, i+1..i+4), measurements
                        = relative poses of the truth track with seeded noise, initial guess = the obs track;
                        and the oracle's solution of that graph (costs per iteration, final poses).
* posegraph_10k.json  — oracle run of BASELINE.json configs[4] (10 000 poses / 39 990 edges, synthetic spiral,
                        oracle/pg_oracle.py make_graph): per-iteration costs, termination, ATE, a strided sample of the
                        final poses.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pg_oracle as pg  # noqa: E402

REF = "/root/reference/st4-kalman/output"


def read_csv(path):
    a = np.array([list(map(float, l.split(","))) for l in open(path) if l.strip() and not l.startswith("#")])
    return np.ascontiguousarray(a[:, 3:7]), np.ascontiguousarray(a[:, :3])


if __name__ == "__main__":
    qT, tT = read_csv(os.path.join(REF, "truth.csv"))
    qO, tO = read_csv(os.path.join(REF, "obs.csv"))
    qT, qO = qT / np.linalg.norm(qT, axis=1, keepdims=True), qO / np.linalg.norm(qO, axis=1, keepdims=True)
    n = len(qT)
    ei, ej = pg.band_edges(n, (1, 2, 3, 4))
    zq, zt = pg.measurements_from(qT, tT, ei, ej, 2e-3, 2e-3, np.random.default_rng(20221109))
    q0, t0 = qO.copy(), tO.copy()
    q0[0], t0[0] = qT[0], tT[0]
    t1 = time.time()
    q, t, s = pg.solve(q0, t0, ei, ej, zq, zt)
    print("st4:", s.brief_report(), s.message, "%.1f s" % (time.time() - t1), "ATE %.4f -> %.4f" % (pg.ate(qT, tT, q0, t0), pg.ate(qT, tT, q, t)))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "posegraph_st4.npz"), q_truth=qT, t_truth=tT, q_obs=qO, t_obs=tO,
                        ei=ei, ej=ej, zq=zq, zt=zt, q0=q0, t0=t0, q_final=q, t_final=t, costs=np.array([i["cost"] for i in s.iterations]),
                        termination=np.array(s.termination_type), message=np.array(s.message))
    G = pg.make_graph(10000)
    t1 = time.time()
    q, t, s = pg.solve(G["q0"], G["t0"], G["ei"], G["ej"], G["zq"], G["zt"])
    print("10k:", s.brief_report(), s.message, "%.1f s" % (time.time() - t1))
    idx = np.arange(0, 10000, 157)
    out = dict(n_poses=10000, n_edges=int(len(G["ei"])), costs=[i["cost"] for i in s.iterations], termination_type=s.termination_type,
               message=s.message, ate_initial=pg.ate(G["q_truth"], G["t_truth"], G["q0"], G["t0"]), ate_final=pg.ate(G["q_truth"], G["t_truth"], q, t),
               sample_index=idx.tolist(), q_sample=q[idx].tolist(), t_sample=t[idx].tolist(), oracle_seconds=time.time() - t1)
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "posegraph_10k.json"), "w"))
