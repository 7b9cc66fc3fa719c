"""Oracle for the two stages in FRONT of the bundle-adjustment path.  TEST INFRASTRUCTURE ONLY.

* `visibility`   restates `ProblemScene::CreateMeasurements`, st20-g2o/src/src/sim_data.cpp:119-142
                 (SURVEY.md §8 a11): per (camera, landmark) `p_c = T_wc^-1 * P`, keep iff
                 `p_c.z >= 0` (:129), `|x/z| < CAM_PLANE_HALF_WIDTH` and `|y/z| < CAM_PLANE_HALF_HEIGHT`
                 (:133-134, constants sim_data.h:211-212); the feature is stored as a float32
                 `pcl::PointXY` (:135-136) and widened to double again at sim_data.cpp:266-269.
                 Emits `landmark -> [(camera, uv)]` (camera-ascending, because the outer loop runs
                 over cameras, :123) and `camera -> [landmark]` (landmark-ascending).
* `triangulate`  restates the per-landmark solve of `ProblemScene::Simulation`,
                 sim_data.cpp:298-311, with the `Triangulation` functor of sim_data.h:165-194:
                 residual = uv - (R_cw P + t_cw).xy / z  (note the sign, :191), one 3-parameter
                 `ceres::Problem` per landmark, default `ceres::Solver::Options`.

The index outputs are a BIT-EXACT contract, so the floating-point predicate is written without
fused multiply-adds and in one fixed order (NumPy never fuses); the CUDA kernel evaluates the same
expression tree with `__dmul_rn/__dadd_rn`:

    R       = quat_to_rot(q)                                   (camera -> world)
    t_cw_i  = -((R[0][i] t0 + R[1][i] t1) + R[2][i] t2)        OptPose::inverse(), sim_data.h:30-32
    p_c_i   =  ((R[0][i] P0 + R[1][i] P1) + R[2][i] P2) + t_cw_i
"""
import numpy as np

from . import dense_lm, lie

HALF_W, HALF_H = 0.8, 0.6          # sim_data.h:211-212


def world_to_camera(cam_q, cam_t):
    """(R^T rows as used below, t_cw) of `OptPose::inverse()` — fixed evaluation order, no FMA."""
    R = lie.quat_to_rot(np.asarray(cam_q, dtype=np.float64))      # [c, row, col]
    t = np.asarray(cam_t, dtype=np.float64)
    tcw = -((R[:, 0, :] * t[:, 0:1] + R[:, 1, :] * t[:, 1:2]) + R[:, 2, :] * t[:, 2:3])
    return R, tcw


def camera_points(R, tcw, pts):
    """p_c for every (landmark, camera): [n_pts, n_cam, 3]."""
    P = np.asarray(pts, dtype=np.float64)
    return ((R[None, :, 0, :] * P[:, None, 0:1] + R[None, :, 1, :] * P[:, None, 1:2]) + R[None, :, 2, :] * P[:, None, 2:3]) + tcw[None]


def visibility(cam_q, cam_t, pts, half_w=HALF_W, half_h=HALF_H, round_uv_f32=True, chunk=4096):
    """Returns dict(lm_deg i32[n_lm], cam_deg i32[n_cam], obs_cam i32[n], obs_lm i32[n], obs_uv f64[n,2],
    cam_lm i32[n]) — observations landmark-major / camera-ascending, `cam_lm` camera-major / landmark-ascending."""
    R, tcw = world_to_camera(cam_q, cam_t)
    pts = np.asarray(pts, dtype=np.float64)
    n_lm, n_cam = len(pts), len(R)
    oc, ol, ouv = [], [], []
    for s in range(0, n_lm, chunk):
        pc = camera_points(R, tcw, pts[s:s + chunk])
        z = pc[..., 2]
        with np.errstate(divide="ignore", invalid="ignore"):
            x, y = pc[..., 0] / z, pc[..., 1] / z
        vis = ~(z < 0.0) & (np.abs(x) < half_w) & (np.abs(y) < half_h)
        l, c = np.nonzero(vis)                       # row-major: landmark-major, camera-ascending
        oc.append(c.astype(np.int32)); ol.append((l + s).astype(np.int32))
        ouv.append(np.stack([x[l, c], y[l, c]], axis=-1))
    obs_cam = np.concatenate(oc) if oc else np.zeros(0, np.int32)
    obs_lm = np.concatenate(ol) if ol else np.zeros(0, np.int32)
    obs_uv = np.concatenate(ouv) if ouv else np.zeros((0, 2))
    if round_uv_f32:
        obs_uv = obs_uv.astype(np.float32).astype(np.float64)
    order = np.argsort(obs_cam, kind="stable")
    return dict(lm_deg=np.bincount(obs_lm, minlength=n_lm).astype(np.int32), cam_deg=np.bincount(obs_cam, minlength=n_cam).astype(np.int32),
                obs_cam=obs_cam, obs_lm=obs_lm, obs_uv=obs_uv, cam_lm=obs_lm[order].astype(np.int32))


def triangulation_residual_jacobian(R, tcw, uv, P):
    """Triangulation functor (sim_data.h:181-193) for one landmark: r [2d], J [2d, 3]."""
    pc = np.einsum("cji,j->ci", R, P) + tcw
    iz = 1.0 / pc[:, 2]
    u, v = pc[:, 0] * iz, pc[:, 1] * iz
    r = np.stack([uv[:, 0] - u, uv[:, 1] - v], axis=-1).ravel()
    # d(u,v)/dP = Pi' R^T ; residual = uv - proj  ->  J = -Pi' R^T
    J0 = -(iz[:, None] * (R[:, :, 0] - u[:, None] * R[:, :, 2]))
    J1 = -(iz[:, None] * (R[:, :, 1] - v[:, None] * R[:, :, 2]))
    return r, np.stack([J0, J1], axis=1).reshape(-1, 3)


def triangulate(cam_q, cam_t, lm0, obs_cam, obs_lm, obs_uv, options=None):
    """Per-landmark Ceres-default LM from `lm0`.  Returns (lm [n,3], iterations i32[n], final_cost f64[n],
    termination list)."""
    R, tcw = world_to_camera(cam_q, cam_t)
    lm = np.array(lm0, dtype=np.float64)
    obs_cam = np.asarray(obs_cam); obs_lm = np.asarray(obs_lm); obs_uv = np.asarray(obs_uv, dtype=np.float64)
    n = len(lm)
    ptr = np.concatenate([[0], np.cumsum(np.bincount(obs_lm, minlength=n))])
    assert np.all(np.diff(obs_lm) >= 0), "observations must be landmark-major"
    its = np.zeros(n, dtype=np.int32); costs = np.zeros(n); term = []
    for l in range(n):
        sl = slice(ptr[l], ptr[l + 1])
        if ptr[l + 1] == ptr[l]:
            term.append("CONVERGENCE"); continue
        Rl, tl, uvl = R[obs_cam[sl]], tcw[obs_cam[sl]], obs_uv[sl]
        x, s = dense_lm.solve(lm[l], lambda P: triangulation_residual_jacobian(Rl, tl, uvl, P), options)
        lm[l] = x
        its[l] = len(s.iterations)
        costs[l] = s.final_cost
        term.append(s.termination_type)
    return lm, its, costs, term
