"""ncu driver: visibility (1000 cameras x 100 000 points) and batched triangulation at config C."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, stba
d = bench.load_scene("C")
v = stba.front.visibility(d["cam_q"], d["cam_t"], d["lm"])
print("visible pairs", len(v["obs_cam"]))
out = stba.front.triangulate(d["cam_q"], d["cam_t"], d["lm"], d["obs_cam"], d["obs_lm"], d["obs_uv"])
print("triangulate kernel ms", out[4])
