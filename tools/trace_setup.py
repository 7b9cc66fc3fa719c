import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["STBA_TRACE"] = "1"
import bench, stba
d = bench.load_scene("C")
for i in range(3):
    t0 = time.perf_counter()
    e = stba.engine.BAEngine(d["cam_q"], d["cam_t"], d["lm"], d["obs_cam"], d["obs_lm"], d["obs_uv"], d["cam_const"])
    t1 = time.perf_counter()
    s = e.solve()
    t2 = time.perf_counter()
    e.get_state()
    t3 = time.perf_counter()
    e.close()
    t4 = time.perf_counter()
    print("create %.1f ms  solve %.1f ms  get_state %.1f ms  close %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3), flush=True)
