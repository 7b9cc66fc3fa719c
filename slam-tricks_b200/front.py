"""The two stages in front of the bundle-adjustment path, on the B200 (`stba_visibility`,
`stba_triangulate`): `ProblemScene::CreateMeasurements` (st20-g2o/src/src/sim_data.cpp:119-142) and the
per-landmark triangulation solves of `ProblemScene::Simulation` (sim_data.cpp:298-311)."""
import ctypes as C

import numpy as np

from . import capi

HALF_W, HALF_H = 0.8, 0.6          # CAM_PLANE_HALF_WIDTH / HEIGHT, sim_data.h:211-212


def visibility(cam_q, cam_t, pts, half_w=HALF_W, half_h=HALF_H, round_uv_f32=True, device=0):
    """Returns dict(lm_deg, cam_deg, obs_cam, obs_lm, obs_uv, cam_lm): `_landmarkCameraVec` flattened
    landmark-major / camera-ascending and `_cameraLandmarkVec` flattened camera-major / landmark-ascending."""
    L = capi.lib()
    n_cam, n_lm = len(cam_q), len(pts)
    q = capi.as_f64(cam_q, (n_cam, 4)); t = capi.as_f64(cam_t, (n_cam, 3)); p = capi.as_f64(pts, (n_lm, 3))
    n = C.c_int64(0)
    lm_deg = np.zeros(n_lm, np.int32); cam_deg = np.zeros(n_cam, np.int32)
    capi.check(L.stba_visibility(device, n_cam, n_lm, capi.dptr(q), capi.dptr(t), capi.dptr(p), half_w, half_h, int(round_uv_f32),
                                 0, C.byref(n), capi.iptr(lm_deg), capi.iptr(cam_deg), None, None, None, None), "stba_visibility")
    m = int(n.value)
    oc = np.zeros(m, np.int32); ol = np.zeros(m, np.int32); uv = np.zeros((m, 2)); cl = np.zeros(m, np.int32)
    if m:
        capi.check(L.stba_visibility(device, n_cam, n_lm, capi.dptr(q), capi.dptr(t), capi.dptr(p), half_w, half_h, int(round_uv_f32),
                                     m, C.byref(n), capi.iptr(lm_deg), capi.iptr(cam_deg), capi.iptr(oc), capi.iptr(ol), capi.dptr(uv),
                                     capi.iptr(cl)), "stba_visibility")
    return dict(lm_deg=lm_deg, cam_deg=cam_deg, obs_cam=oc, obs_lm=ol, obs_uv=uv, cam_lm=cl)


def triangulate(cam_q, cam_t, lm0, obs_cam, obs_lm, obs_uv, options=None, device=0):
    """Returns (lm [n,3], iterations i32[n], final_cost f64[n], termination list[str], kernel_ms)."""
    L = capi.lib()
    n_cam, n_lm, n_obs = len(cam_q), len(lm0), len(obs_cam)
    q = capi.as_f64(cam_q, (n_cam, 4)); t = capi.as_f64(cam_t, (n_cam, 3))
    lm = np.array(lm0, dtype=np.float64, order="C").reshape(n_lm, 3)
    oc = np.ascontiguousarray(obs_cam, dtype=np.int32); ol = np.ascontiguousarray(obs_lm, dtype=np.int32)
    uv = capi.as_f64(obs_uv, (n_obs, 2))
    its = np.zeros(n_lm, np.int32); cost = np.zeros(n_lm); term = np.zeros(n_lm, np.int32); ms = C.c_float(0)
    capi.check(L.stba_triangulate(device, n_cam, n_lm, n_obs, capi.dptr(q), capi.dptr(t), capi.dptr(lm), capi.iptr(oc), capi.iptr(ol),
                                  capi.dptr(uv), C.byref(options) if options is not None else None, capi.iptr(its), capi.dptr(cost),
                                  capi.iptr(term), C.byref(ms)), "stba_triangulate")
    return lm, its, cost, [capi.TERMINATION[int(x)] for x in term], float(ms.value)


def _tptr(t, ctype):
    return C.cast(C.c_void_p(t.data_ptr()), C.POINTER(ctype))


def visibility_device(cam_q, cam_t, pts, half_w=HALF_W, half_h=HALF_H, round_uv_f32=True):
    """Device hand-off form of `visibility`: torch CUDA tensors in, torch CUDA tensors out (the C entry point takes
    host or device pointers alike).  Nothing but the 8-byte observation count crosses PCIe; the lists feed
    `triangulate_device` and `engine.BAEngine` directly."""
    import torch
    L = capi.lib()
    dev = cam_q.device
    n_cam, n_lm = cam_q.shape[0], pts.shape[0]
    q = cam_q.contiguous(); t = cam_t.contiguous(); p = pts.contiguous()
    n = C.c_int64(0)
    lm_deg = torch.empty(n_lm, dtype=torch.int32, device=dev); cam_deg = torch.empty(n_cam, dtype=torch.int32, device=dev)
    args = (dev.index or 0, n_cam, n_lm, _tptr(q, C.c_double), _tptr(t, C.c_double), _tptr(p, C.c_double), half_w, half_h, int(round_uv_f32))
    capi.check(L.stba_visibility(*args, 0, C.byref(n), None, None, None, None, None, None), "stba_visibility")
    m = int(n.value)
    oc = torch.empty(m, dtype=torch.int32, device=dev); ol = torch.empty(m, dtype=torch.int32, device=dev)
    uv = torch.empty((m, 2), dtype=torch.float64, device=dev); cl = torch.empty(m, dtype=torch.int32, device=dev)
    if m:
        capi.check(L.stba_visibility(*args, m, C.byref(n), _tptr(lm_deg, C.c_int32), _tptr(cam_deg, C.c_int32), _tptr(oc, C.c_int32),
                                     _tptr(ol, C.c_int32), _tptr(uv, C.c_double), _tptr(cl, C.c_int32)), "stba_visibility")
    return dict(lm_deg=lm_deg, cam_deg=cam_deg, obs_cam=oc, obs_lm=ol, obs_uv=uv, cam_lm=cl)


def triangulate_device(cam_q, cam_t, lm0, obs_cam, obs_lm, obs_uv, options=None):
    """Device hand-off form of `triangulate` (torch CUDA tensors).  Returns (lm, iterations, final_cost, termination, kernel_ms),
    the per-landmark outputs as CUDA tensors."""
    import torch
    L = capi.lib()
    dev = cam_q.device
    n_cam, n_lm, n_obs = cam_q.shape[0], lm0.shape[0], obs_cam.shape[0]
    lm = lm0.clone().contiguous()
    its = torch.empty(n_lm, dtype=torch.int32, device=dev); cost = torch.empty(n_lm, dtype=torch.float64, device=dev)
    term = torch.empty(n_lm, dtype=torch.int32, device=dev); ms = C.c_float(0)
    capi.check(L.stba_triangulate(dev.index or 0, n_cam, n_lm, n_obs, _tptr(cam_q.contiguous(), C.c_double), _tptr(cam_t.contiguous(), C.c_double),
                                  _tptr(lm, C.c_double), _tptr(obs_cam.contiguous(), C.c_int32), _tptr(obs_lm.contiguous(), C.c_int32),
                                  _tptr(obs_uv.contiguous(), C.c_double), C.byref(options) if options is not None else None,
                                  _tptr(its, C.c_int32), _tptr(cost, C.c_double), _tptr(term, C.c_int32), C.byref(ms)), "stba_triangulate")
    return lm, its, cost, term, float(ms.value)


PNP_JACOBIAN_REFERENCE, PNP_JACOBIAN_EXACT = 0, 1


def pnp_gauss_newton(points, uv, q0, t0, ptr=None, max_iterations=10, tolerance=1e-8, jacobian="reference", device=0):
    """`SelfGaussNewton` (st17-ceres/src/include/solver.hpp:387-462) on the GPU, one kernel for the whole loop.
    Single problem: points [n,3], uv [n,2], q0 [4], t0 [3].  Batch: pass ptr (i32[n_problems+1]) and q0 [p,4], t0 [p,3].
    Returns (q, t, iterations, last_change, kernel_ms); `iterations` is the reference's "iter num"."""
    L = capi.lib()
    single = ptr is None
    pts = capi.as_f64(points, (len(points), 3)); uvs = capi.as_f64(uv, (len(uv), 2))
    ptr = np.ascontiguousarray([0, len(pts)] if single else ptr, dtype=np.int32)
    n = len(ptr) - 1
    q = np.array(q0, dtype=np.float64, order="C").reshape(n, 4); t = np.array(t0, dtype=np.float64, order="C").reshape(n, 3)
    its = np.zeros(n, np.int32); change = np.zeros(n); ms = C.c_float(0)
    capi.check(L.stba_pnp_gauss_newton(device, n, capi.iptr(ptr), capi.dptr(pts), capi.dptr(uvs), capi.dptr(q), capi.dptr(t), max_iterations, tolerance,
                                       PNP_JACOBIAN_REFERENCE if jacobian == "reference" else PNP_JACOBIAN_EXACT, capi.iptr(its), capi.dptr(change),
                                       C.byref(ms)), "stba_pnp_gauss_newton")
    if single:
        return q[0], t[0], int(its[0]), float(change[0]), float(ms.value)
    return q, t, its, change, float(ms.value)
