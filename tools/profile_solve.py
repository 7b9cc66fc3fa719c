"""Small driver for ncu: N full LM solves of a workload on cuda:0 (no timing claims — profiler run)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import stba  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C")
    ap.add_argument("--solves", type=int, default=2)
    ap.add_argument("--dense", default="default")
    a = ap.parse_args()
    d = bench.load_scene(a.workload)
    opt = stba.capi.Options()
    if a.dense != "default":
        opt.dense_backend = stba.capi.DENSE_OWN if a.dense == "own" else stba.capi.DENSE_CUSOLVER
    with stba.engine.BAEngine(d["cam_q"], d["cam_t"], d["lm"], d["obs_cam"], d["obs_lm"], d["obs_uv"], d["cam_const"]) as e:
        e.save_state()
        for _ in range(a.solves):
            e.restore_state()
            s = e.solve(opt)
        print(s.BriefReport(), s.phase_ms, s.gpu_launches)
