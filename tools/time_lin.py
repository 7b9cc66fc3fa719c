"""Linearisation timing at C and on k side-by-side copies (10 M observations): the one-launch k_lin3 against the
three-launch path of rounds 1-2 (STBA_LIN2=1, separate process because the switch is read at engine creation)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, stba

def run(k):
    d = bench.load_scene("C")
    if k > 1:
        nc, nl = len(d["cam_q"]), len(d["lm"])
        args = (np.tile(d["cam_q"], (k, 1)), np.tile(d["cam_t"], (k, 1)), np.tile(d["lm"], (k, 1)),
                np.concatenate([d["obs_cam"] + i * nc for i in range(k)]), np.concatenate([d["obs_lm"] + i * nl for i in range(k)]),
                np.tile(d["obs_uv"], (k, 1)), np.tile(d["cam_const"], k))
    else:
        args = (d["cam_q"], d["cam_t"], d["lm"], d["obs_cam"], d["obs_lm"], d["obs_uv"], d["cam_const"])
    with stba.engine.BAEngine(*args, linearize_only=True) as e:
        ms = e.time_phase("linearize", reps=24, flush_l2=True)[4:]
        e.linearize()
        out = {"copies": k, "lin_ms_mean": float(ms.mean()), "lin_ms_min": float(ms.min())}
        if os.environ.get("STBA_LIN2"):
            out["lin_lm_ms"] = float(e.time_phase("lin_lm", reps=8, flush_l2=True)[2:].mean())
            out["lin_cam_ms"] = float(e.time_phase("lin_cam", reps=8, flush_l2=True)[2:].mean())
        H = e.blocks()
        out["cost"] = H[4]
        out["chk"] = [float(np.abs(x).sum()) for x in H[:4]]
    return out

if __name__ == "__main__":
    for k in [1] + [int(a) for a in sys.argv[1:]]:
        print(json.dumps({"path": "lin2 (3 launches)" if os.environ.get("STBA_LIN2") else "lin3 (1 launch)", **run(k)}), flush=True)
