"""Mirror of `ns_st3::CalibSolver` (st3-calibration/src/include/calib.h, src/src/calib.cpp) over the C ABI:
same method names, same pipeline (`solve()` = computeHomoMats -> reconstructIntriMat ->
reconstructExtriMat -> totalOptimization, calib.cpp:38-47); the joint Gauss-Newton runs on the B200."""
import ctypes as C
import os

import numpy as np

from . import capi


def read_corners(path):
    """`CBCorners::read`, st3-calibration/src/src/cbcorner.cpp:50-72: header `rows,cols`, then `r,c,x,y`;
    x, y are parsed with std::stof (:67-68) — float32, then widened."""
    with open(path) as f:
        rows, cols = (int(x) for x in f.readline().strip().split(","))
        pts = np.zeros((rows, cols, 2))
        for line in f:
            if line.strip():
                r, c, x, y = line.strip().split(",")
                pts[int(r), int(c)] = (np.float32(x), np.float32(y))
    return rows, cols, pts


def write_corners(path, pts):
    """`CBCorners::write`, cbcorner.cpp:34-48 (3 decimals)."""
    rows, cols = pts.shape[:2]
    with open(path, "w") as f:
        f.write("%d,%d\n" % (rows, cols))
        for i in range(rows):
            for j in range(cols):
                f.write("%d,%d,%.3f,%.3f\n" % (i, j, pts[i, j, 0], pts[i, j, 1]))


class CalibSolver:
    def __init__(self, corner_dir=None, chess_board_size=2.8e-2, views=None, device=0):
        """`CalibSolver(cornerDir, chessBoardSize)`, calib.cpp:11-36: every file of the directory in
        lexicographic order (helper.cpp:4-11); corner (i, j) has board coordinates (j, i) * cbSize.
        Alternatively pass `views` = list of (obj_xy [n,2], img_uv [n,2])."""
        self.cbSize, self.device = chess_board_size, device
        objs, imgs = [], []
        if corner_dir is not None:
            for name in sorted(os.listdir(corner_dir)):
                p = os.path.join(corner_dir, name)
                if os.path.isdir(p):
                    continue
                rows, cols, pts = read_corners(p)
                j, i = np.meshgrid(np.arange(cols), np.arange(rows))
                objs.append(np.stack([j.ravel() * chess_board_size, i.ravel() * chess_board_size], axis=-1).astype(np.float64))
                imgs.append(pts.reshape(-1, 2))
        for o, m in (views or []):
            objs.append(np.asarray(o, dtype=np.float64).reshape(-1, 2)); imgs.append(np.asarray(m, dtype=np.float64).reshape(-1, 2))
        self.cbsCount = len(objs)
        self.view_ptr = np.concatenate([[0], np.cumsum([len(o) for o in objs])]).astype(np.int32)
        self.obj = np.ascontiguousarray(np.concatenate(objs)) if objs else np.zeros((0, 2))
        self.img = np.ascontiguousarray(np.concatenate(imgs)) if imgs else np.zeros((0, 2))
        self.intrinsics = np.zeros(4)          # alpha, beta, u0, v0
        self.distortion = np.zeros(5)          # k1, k2, k3, p1, p2
        self.imgPos = np.zeros((self.cbsCount, 6))   # se3.log(), Sophus order [rho, theta]
        self.HomoMats = np.zeros((self.cbsCount, 3, 3))
        self.update_norms, self.costs, self.gpu_launches = [], [], 0

    def solve(self):
        self.initialize()
        self.totalOptimization()
        return self

    def initialize(self):
        """computeHomoMats + reconstructIntriMat + reconstructExtriMat (host)."""
        capi.check(capi.lib().stba_calib_initialize(self.cbsCount, capi.iptr(self.view_ptr), capi.dptr(self.obj), capi.dptr(self.img),
                                                   capi.dptr(self.intrinsics), capi.dptr(self.imgPos), capi.dptr(self.HomoMats)),
                   "stba_calib_initialize")
        self.distortion[:] = 0.0               # calib.cpp:290
        return self

    def totalOptimization(self, max_iterations=10, tolerance=1e-8):
        it = C.c_int32(0); nl = C.c_int64(0)
        loop_ms = C.c_float(0); acc_ms = C.c_float(0)
        norms = np.zeros(max(max_iterations, 1)); costs = np.zeros(max(max_iterations, 1))
        capi.check(capi.lib().stba_calib_optimize_timed(self.device, self.cbsCount, capi.iptr(self.view_ptr), capi.dptr(self.obj), capi.dptr(self.img),
                                                       capi.dptr(self.intrinsics), capi.dptr(self.distortion), capi.dptr(self.imgPos),
                                                       max_iterations, tolerance, C.byref(it), capi.dptr(norms), capi.dptr(costs), C.byref(nl),
                                                       C.byref(loop_ms), C.byref(acc_ms)),
                   "stba_calib_optimize_timed")
        self.loop_ms, self.accumulate_ms = float(loop_ms.value), float(acc_ms.value)
        self.update_norms, self.costs, self.gpu_launches = norms[:it.value].tolist(), costs[:it.value].tolist(), int(nl.value)
        return self

    def __str__(self):
        # operator<<, calib.cpp:424-431
        a, b, u0, v0 = self.intrinsics
        k1, k2, k3, p1, p2 = self.distortion
        return ("{'alpha(fx)': %g, 'beta(fy)': %g, 'u0(cx)': %g, 'v0(cy)': %g, 'k1': %g, 'k2': %g, 'k3': %g, 'p1': %g, 'p2': %g}"
                % (a, b, u0, v0, k1, k2, k3, p1, p2))
