import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["STBA_TRACE"] = "1"
import numpy as np, torch
import bench, stba
d = bench.load_scene("C")
keys = ("cam_q", "cam_t", "lm", "obs_cam", "obs_lm", "obs_uv", "cam_const")
keep = {k: torch.from_numpy(np.ascontiguousarray(d[k])).pin_memory() for k in keys}
hp = {k: v.numpy() for k, v in keep.items()}
for i in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e = stba.engine.BAEngine(hp["cam_q"], hp["cam_t"], hp["lm"], hp["obs_cam"], hp["obs_lm"], hp["obs_uv"], hp["cam_const"])
    t1 = time.perf_counter()
    s = e.solve()
    t2 = time.perf_counter()
    e.get_state()
    t3 = time.perf_counter()
    e.close()
    t4 = time.perf_counter()
    print("create %.2f ms  solve %.2f ms (device phases %.2f)  get_state %.2f ms  close %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, sum(s.phase_ms.values()), (t3 - t2) * 1e3, (t4 - t3) * 1e3), flush=True)
