"""Dense phase at C: own DAG solve vs the split path (STBA_CHOL_SPLIT=1, one rank: no exchange), solution compared."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, stba
d = bench.load_scene("C")
with stba.engine.BAEngine(d["cam_q"], d["cam_t"], d["lm"], d["obs_cam"], d["obs_lm"], d["obs_uv"], d["cam_const"]) as e:
    ms = e.time_phase("dense_own", reps=8)[2:]
    e.linearize(); e.reduced_system(1e4, fetch=False)
    yc, yl, mcc = e.solve_step()
    print(json.dumps({"split": bool(os.environ.get("STBA_CHOL_SPLIT")), "m": os.environ.get("STBA_CHOL_SPLIT_M"), "dense_ms": float(ms.mean()), "min": float(ms.min()),
                      "yc_norm": float(np.linalg.norm(yc)), "yc_sum": float(yc.sum()), "mcc": float(mcc)}))
