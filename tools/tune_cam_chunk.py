"""Tuning experiment: camera-major chunk size vs lin_cam / schur time on workload C."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, stba
d = bench.load_scene("C")
for cs in sys.argv[1:]:
    os.environ["STBA_CAM_CHUNK"] = cs
    with stba.engine.BAEngine(d["cam_q"], d["cam_t"], d["lm"], d["obs_cam"], d["obs_lm"], d["obs_uv"], d["cam_const"]) as e:
        e.linearize()
        lc = e.time_phase("lin_cam", reps=12, flush_l2=True)[2:]
        ll = e.time_phase("lin_lm", reps=12, flush_l2=True)[2:]
        li = e.time_phase("linearize", reps=12, flush_l2=True)[2:]
        sc = e.time_phase("schur", reps=6)[1:]
        print("chunk %5s: lin_cam %.2f us  lin_lm %.2f us  linearize %.2f us  schur %.1f us" % (cs, 1e3 * lc.mean(), 1e3 * ll.mean(), 1e3 * li.mean(), 1e3 * sc.mean()))
