"""ncu driver: a few linearisations of workload C on cuda:0."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, stba
d = bench.load_scene("C")
with stba.engine.BAEngine(d["cam_q"], d["cam_t"], d["lm"], d["obs_cam"], d["obs_lm"], d["obs_uv"], d["cam_const"], linearize_only=True) as e:
    for _ in range(4):
        e.linearize()
