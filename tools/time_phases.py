"""Isolated phase times at C (CUDA events on the engine stream)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, stba
d = bench.load_scene("C")
with stba.engine.BAEngine(d["cam_q"], d["cam_t"], d["lm"], d["obs_cam"], d["obs_lm"], d["obs_uv"], d["cam_const"]) as e:
    print(json.dumps({ph: round(float(e.time_phase(ph, reps=8)[2:].mean()), 4) for ph in (sys.argv[1:] or ["linearize", "schur", "dense_own", "backsub", "cost"])}))
