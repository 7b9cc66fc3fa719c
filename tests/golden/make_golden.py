"""Generates tests/golden/*.json with the CPU oracle (run from the repo root, CPU only):
    python tests/golden/make_golden.py
ba_C_oracle.json — Ceres-faithful LM on BASELINE.json configs[2] (1k cameras / 100k landmarks /
1M observations, seeds 20221105/20221106) by oracle.ba_oracle.solve(backend="c"): per-iteration
costs, termination, and checksums + a strided sample of the final state."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import stba  # noqa: E402
from oracle import ba_fast, ba_oracle  # noqa: E402

if __name__ == "__main__":
    ba_fast.set_num_threads(os.cpu_count() or 1)
    sc = stba.synth.make_scene(*stba.synth.CONFIGS["C"])
    q, t, l, s = ba_oracle.solve(sc.cam_q, sc.cam_t, sc.lm, sc.obs_cam, sc.obs_lm, sc.obs_uv, sc.cam_const, backend="c")
    state = {}
    for name, a in (("cam_q", q), ("cam_t", t), ("lm", l)):
        idx = np.arange(0, a.size, max(1, a.size // 64))[:64]
        state[name] = dict(sum=float(a.sum()), sample_index=idx.tolist(), sample=a.reshape(-1)[idx].tolist())
    out = dict(config="C", sizes=[sc.n_cam, sc.n_lm, sc.n_obs], termination_type=s.termination_type, message=s.message,
               initial_cost=s.initial_cost, final_cost=s.iterations[-1]["cost"], costs=[i["cost"] for i in s.iterations],
               state=state)
    with open(os.path.join(ROOT, "tests", "golden", "ba_C_oracle.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(s.brief_report(), s.message)
