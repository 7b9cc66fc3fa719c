"""Generates tests/golden/calib_fixture.json (run in the build container, where /root/reference exists):
    python tests/golden/make_calib_golden.py
* `views`   — the reference's own calibration input, st3-calibration/calib/1..9.txt (9 views x 5x8 corners),
              as `CBCorners::read` parses it (float32-rounded pixel coordinates, cbcorner.cpp:67-68),
              board size 2.8e-2 (st3-calibration/src/main.cpp:4).  The reference ships NO expected outputs.
* `oracle`  — outputs of oracle/calib_oracle.py (the literal NumPy restatement of calib.cpp) on those views:
              initial intrinsics / poses, final intrinsics / distortion / poses, per-iteration update norms, costs.
"""
import glob
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import calib_oracle as co  # noqa: E402

REF = "/root/reference/st3-calibration/calib"

if __name__ == "__main__":
    files = sorted(glob.glob(os.path.join(REF, "*.txt")))          # helper.cpp:4-11: lexicographic
    views, objs, imgs = [], [], []
    for f in files:
        rows, cols, pts = co.read_corners(f)
        views.append(dict(file=os.path.basename(f), rows=rows, cols=cols, corners=pts.reshape(-1, 2).tolist()))
        objs.append(co.object_points(rows, cols, 2.8e-2)); imgs.append(pts.reshape(-1, 2))
    K0, poses0, Hs, K, D, poses, info = co.solve(objs, imgs)
    out = dict(cb_size=2.8e-2, views=views,
               oracle=dict(init_intrinsics=K0.tolist(), init_poses=poses0.tolist(), intrinsics=K.tolist(), distortion=D.tolist(),
                           poses=poses.tolist(), update_norms=info["update_norms"], costs=info["costs"], iterations=info["iterations"]))
    with open(os.path.join(ROOT, "tests", "golden", "calib_fixture.json"), "w") as f:
        json.dump(out, f)
    print("views", len(views), "K", K, "D", D, "iterations", info["iterations"])
