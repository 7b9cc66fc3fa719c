"""Phase timeline of k_lin3 (instrumented build: STBA_LIB=.../libstba_timing.so, built with -DSTBA_L3_TIMING)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, stba

k = int(sys.argv[1]) if len(sys.argv) > 1 else 1
d = bench.load_scene("C")
if k > 1:
    nc, nl = len(d["cam_q"]), len(d["lm"])
    args = (np.tile(d["cam_q"], (k, 1)), np.tile(d["cam_t"], (k, 1)), np.tile(d["lm"], (k, 1)),
            np.concatenate([d["obs_cam"] + i * nc for i in range(k)]), np.concatenate([d["obs_lm"] + i * nl for i in range(k)]),
            np.tile(d["obs_uv"], (k, 1)), np.tile(d["cam_const"], k))
else:
    args = (d["cam_q"], d["cam_t"], d["lm"], d["obs_cam"], d["obs_lm"], d["obs_uv"], d["cam_const"])
with stba.engine.BAEngine(*args, linearize_only=True) as e:
    ms = e.time_phase("linearize", reps=6, flush_l2=True)
    L = stba.capi.lib()
    buf = np.zeros(160 * 64, dtype=np.int64)
    L.stba_debug_l3_clocks.restype = C.c_int
    assert L.stba_debug_l3_clocks(buf.ctypes.data_as(C.c_void_p)) == 0
    c = buf.reshape(160, 64)[:148].astype(np.float64)
    t0 = c[:, 0].min()
    us = lambda x: (x - t0) * 1e-3
    print("event ms (last rep):", ms[-1] * 1e3, "us")
    print("kernel span (max end - min start): %.1f us" % max(us(c[:, 5]).max(), us(c[:, 6]).max()))
    for name, col in (("start", 0), ("after Rt + syncthreads", 1), ("table staged", 2), ("lm chunks done", 3), ("lm group enters cam queue", 4),
                      ("thread 0 end", 5), ("warp 8 end", 6)):
        v = us(c[:, col])
        print("%-28s min %7.1f  mean %7.1f  max %7.1f" % (name, v.min(), v.mean(), v.max()))
    print("cam chunks done by warp 0: mean %.1f max %d | by warp 8: mean %.1f max %d" % (c[:, 7].mean(), c[:, 7].max(), c[:, 63].mean(), c[:, 63].max()))
    r = c[:, 59].sum()
    print("warp 8 rounds: %d per CTA; cycles per round: stream wait %.0f, gather wait %.0f, compute %.0f" % (c[:, 59].mean(), c[:, 56].sum() / r, c[:, 57].sum() / r, c[:, 58].sum() / r))
    nch = c[:, 63].sum()
    print("warp 8 per chunk: cycles in the round loop %.0f, epilogue (fold, store, fence, ticket, finish) %.0f" % (c[:, 60].sum() / nch, c[:, 61].sum() / nch))
    for b in (0, 73, 147):
        row = c[b]
        print("CTA %d lm chunks (wait end, compute end):" % b, " ".join("%.1f/%.1f" % (us(row[8 + 2 * i]), us(row[9 + 2 * i])) for i in range(8) if row[8 + 2 * i] > 0))
        print("CTA %d warp 8 chunk ends:" % b, " ".join("%.1f" % us(row[32 + i]) for i in range(0, 24) if row[32 + i] > 0))
