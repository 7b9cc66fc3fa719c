"""SE(3) pose graph on the B200 (`stba_pg_*`, SURVEY.md §8 f1 / BASELINE.json configs[4]).  Inputs follow the
reference's pose-chain simulator and its recorded tracks (st4-kalman/src/src/pose_simulation.cpp:17-88,
st4-kalman/output/{truth,obs}.csv: `x,y,z,qx,qy,qz,qw` per line, st4-kalman/src/main.cpp:7-29)."""
import ctypes as C

import numpy as np

from . import capi, engine


def read_trajectory_csv(path):
    """The trajectory dump of st4-kalman/src/main.cpp:7-29: '#' comment lines, then `x,y,z,qx,qy,qz,qw`.
    Returns (q [n,4] xyzw, t [n,3])."""
    rows = [list(map(float, line.split(","))) for line in open(path) if line.strip() and not line.startswith("#")]
    a = np.array(rows, dtype=np.float64)
    return np.ascontiguousarray(a[:, 3:7]), np.ascontiguousarray(a[:, :3])


def write_trajectory_csv(path, q, t, fps=10.0):
    with open(path, "w") as f:
        f.write("# fps: %f\n# x,y,z,qx,qy,qz,qw\n" % fps)
        for qi, ti in zip(q, t):
            f.write(",".join("%g" % v for v in (*ti, *qi)) + "\n")


class PoseGraph:
    def __init__(self, q, t, ei, ej, zq, zt, device=0):
        self.n, self.m = len(q), len(ei)
        q = capi.as_f64(q, (self.n, 4)); t = capi.as_f64(t, (self.n, 3))
        ei = np.ascontiguousarray(ei, dtype=np.int32); ej = np.ascontiguousarray(ej, dtype=np.int32)
        zq = capi.as_f64(zq, (self.m, 4)); zt = capi.as_f64(zt, (self.m, 3))
        self._h = C.c_void_p()
        self._L = capi.lib()
        capi.check(self._L.stba_pg_create(C.byref(self._h), device, self.n, self.m, capi.dptr(q), capi.dptr(t), capi.iptr(ei), capi.iptr(ej),
                                          capi.dptr(zq), capi.dptr(zt)), "stba_pg_create")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.stba_pg_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def linearize(self):
        cost = C.c_double(0); bw = C.c_int32(0)
        g = np.zeros((self.n, 6)); H = np.zeros((self.n, 6, 6))
        capi.check(self._L.stba_pg_linearize(self._h, C.byref(cost), capi.dptr(g), capi.dptr(H), C.byref(bw)), "stba_pg_linearize")
        return cost.value, g, H, bw.value

    def solve(self, options=None, callback=None):
        return engine.run_solve(self._L.stba_pg_solve, self._h, options, callback)

    def set_state(self, q, t):
        capi.check(self._L.stba_pg_set_state(self._h, capi.dptr(capi.as_f64(q, (self.n, 4))), capi.dptr(capi.as_f64(t, (self.n, 3)))), "stba_pg_set_state")

    def save_state(self):
        capi.check(self._L.stba_pg_save_state(self._h), "stba_pg_save_state")

    def restore_state(self):
        capi.check(self._L.stba_pg_restore_state(self._h), "stba_pg_restore_state")

    def time_linearize(self, reps=10):
        ms = (C.c_float * reps)()
        capi.check(self._L.stba_pg_time_linearize(self._h, reps, ms), "stba_pg_time_linearize")
        return np.array(list(ms), dtype=np.float64)

    def get_state(self):
        q = np.zeros((self.n, 4)); t = np.zeros((self.n, 3))
        capi.check(self._L.stba_pg_get_state(self._h, capi.dptr(q), capi.dptr(t)), "stba_pg_get_state")
        return q, t
