"""NumPy oracle of the Zhang-calibration path.  TEST INFRASTRUCTURE ONLY.

Restates `ns_st3::CalibSolver` (st3-calibration/src/src/calib.cpp) literally:

* `read_corners`          CBCorners::read, st3-calibration/src/src/cbcorner.cpp:50-72 — pixel coordinates are
                          parsed with `std::stof` (:67-68), i.e. rounded to float32 and widened
* `object_points`         CalibSolver::init, calib.cpp:16-36: corner (i, j) -> (j, i) * cbSize
* `homography`            computeHomoMat, calib.cpp:55-93 (DLT, last right singular vector as-is)
* `intrinsics_from_homographies`  reconstructIntriMat, calib.cpp:95-140
* `extrinsics`            reconstructExtriMat, calib.cpp:142-173
* `total_optimization`    totalOptimization, calib.cpp:282-422 — plain Gauss-Newton, dense H of size
                          9 + 6 V, `H.ldlt().solve(g)`, <= 10 iterations, stop when |update| < 1e-8;
                          pose update `log(exp(d) * exp(pose))` (LEFT perturbation, Sophus tangent
                          order [rho, theta]), residual in pixels.

Sophus (`SE3d::exp/log`) is a third-party dependency absent from /root/reference; its published
closed forms are restated in `se3_exp` / `se3_log`.  The reference ships the input fixture
(`st3-calibration/calib/1..9.txt`) but NO expected outputs: parity for this path is pinned to the
inputs only (PARITY UNPINNED for the outputs).
"""
import numpy as np

from . import lie

EPS = lie.SOPHUS_EPS


def read_corners(path):
    with open(path) as f:
        rows, cols = (int(x) for x in f.readline().strip().split(","))
        pts = np.zeros((rows, cols, 2))
        for line in f:
            line = line.strip()
            if not line:
                continue
            r, c, x, y = line.split(",")
            pts[int(r), int(c)] = (np.float32(x), np.float32(y))       # std::stof
    return rows, cols, pts


def object_points(rows, cols, cb_size):
    j, i = np.meshgrid(np.arange(cols), np.arange(rows))
    return np.stack([j.ravel() * cb_size, i.ravel() * cb_size], axis=-1).astype(np.float64)


# ------------------------------------------------------------------ Sophus SE3
def so3_exp_matrix(omega):
    return lie.quat_to_rot(lie.so3_exp_quat(omega))


def se3_exp(xi):
    """Sophus::SE3d::exp: xi = [upsilon, omega] -> (R, t = V upsilon)."""
    ups, om = np.asarray(xi[:3], dtype=np.float64), np.asarray(xi[3:], dtype=np.float64)
    theta = np.linalg.norm(om)
    R = so3_exp_matrix(om)
    Om = lie.hat(om)
    if abs(theta) < EPS:
        V = R
    else:
        V = np.eye(3) + (1 - np.cos(theta)) / theta ** 2 * Om + (theta - np.sin(theta)) / theta ** 3 * (Om @ Om)
    return R, V @ ups


def se3_log(R, t):
    """Sophus::SE3d::log."""
    om = lie.so3_log_quat(lie.rot_to_quat(R))
    theta = np.linalg.norm(om)
    Om = lie.hat(om)
    if abs(theta) < EPS:
        Vinv = np.eye(3) - 0.5 * Om + (1.0 / 12.0) * (Om @ Om)
    else:
        h = 0.5 * theta
        Vinv = np.eye(3) - 0.5 * Om + (1 - theta * np.cos(h) / (2 * np.sin(h))) / theta ** 2 * (Om @ Om)
    return np.concatenate([Vinv @ t, om])


# ------------------------------------------------------------------ initialisation (a15)
def homography(img, obj, fix_sign=True):
    """DLT.  The reference takes `svd.matrixV().col(8)` as-is (calib.cpp:80-81), so the sign of H is
    whatever Eigen's JacobiSVD returns; H and -H yield the same intrinsics, distortion and cost but
    mirrored poses (board behind the camera).  Eigen is not available to pin that sign, so the oracle
    and the product both choose H[2,2] > 0 (board in front of the camera)."""
    n = len(img)
    A = np.zeros((2 * n, 9))
    x, y, u, v = obj[:, 0], obj[:, 1], img[:, 0], img[:, 1]
    A[0::2, 0], A[0::2, 1], A[0::2, 2] = x, y, 1.0
    A[0::2, 6], A[0::2, 7], A[0::2, 8] = -u * x, -u * y, -u
    A[1::2, 3], A[1::2, 4], A[1::2, 5] = x, y, 1.0
    A[1::2, 6], A[1::2, 7], A[1::2, 8] = -v * x, -v * y, -v
    h = np.linalg.svd(A)[2][8]
    if fix_sign and h[8] < 0:
        h = -h
    return h.reshape(3, 3)


def intrinsics_from_homographies(Hs):
    def cof(H, i, j):
        hi, hj = H[:, i], H[:, j]
        return np.array([hi[0] * hj[0], hi[2] * hj[0] + hi[0] * hj[2], hi[1] * hj[1], hi[2] * hj[1] + hi[1] * hj[2], hi[2] * hj[2]])
    C = np.zeros((2 * len(Hs), 5))
    for i, H in enumerate(Hs):
        C[2 * i] = cof(H, 0, 1)
        C[2 * i + 1] = cof(H, 0, 0) - cof(H, 1, 1)
    b11, b13, b22, b23, b33 = np.linalg.svd(C)[2][4]
    v0 = -b23 / b22
    lam = b33 - (b13 * b13 - v0 * b11 * b23) / b11
    alpha, beta = np.sqrt(lam / b11), np.sqrt(lam / b22)
    u0 = -b13 * alpha * alpha / lam
    return np.array([alpha, beta, u0, v0])


def extrinsics(K4, Hs):
    alpha, beta, u0, v0 = K4
    Kinv = np.linalg.inv(np.array([[alpha, 0, u0], [0, beta, v0], [0, 0, 1.0]]))
    poses = []
    for H in Hs:
        r1, r2 = Kinv @ H[:, 0], Kinv @ H[:, 1]
        lam = 1.0 / (2.0 * np.linalg.norm(r1)) + 1.0 / (2.0 * np.linalg.norm(r2))
        r1, r2 = r1 / np.linalg.norm(r1), r2 / np.linalg.norm(r2)
        r3 = np.cross(r1, r2)
        r1 = np.cross(r2, r3)
        t = lam * (Kinv @ H[:, 2])
        U, _, Vt = np.linalg.svd(np.stack([r1, r2, r3], axis=1))
        poses.append(se3_log(U @ Vt, t))
    return np.array(poses)


# ------------------------------------------------------------------ total optimisation (a14)
def residual_jacobian(K4, D5, pose, obj, img):
    """One view: e [n,2] and the three Jacobian groups J_intri [n,4,2], J_dist [n,5,2], J_pos [n,6,2]
    exactly as calib.cpp:318-380."""
    alpha, beta, u0, v0 = K4
    k1, k2, k3, p1, p2 = D5
    R, t = se3_exp(pose)
    P = (R @ np.stack([obj[:, 0], obj[:, 1], np.zeros(len(obj))], axis=0)).T + t
    Xp, Yp, Zp = P[:, 0], P[:, 1], P[:, 2]
    xn, yn = Xp / Zp, Yp / Zp
    r2 = xn * xn + yn * yn
    r4, r6 = r2 * r2, r2 * r2 * r2
    rad = 1.0 + k1 * r2 + k2 * r4 + k3 * r6
    xd = xn * rad + 2.0 * p1 * xn * yn + p2 * (r2 + 2.0 * xn * xn)
    yd = yn * rad + 2.0 * p2 * xn * yn + p1 * (r2 + 2.0 * yn * yn)
    e = np.stack([alpha * xd + u0 - img[:, 0], beta * yd + v0 - img[:, 1]], axis=-1)
    n = len(obj)
    Ji = np.zeros((n, 4, 2))
    Ji[:, 0, 0], Ji[:, 2, 0], Ji[:, 1, 1], Ji[:, 3, 1] = xd, 1.0, yd, 1.0
    Jd = np.zeros((n, 5, 2))
    Jd[:, 0, 0], Jd[:, 0, 1] = alpha * xn * r2, beta * yn * r2
    Jd[:, 1, 0], Jd[:, 1, 1] = alpha * xn * r4, beta * yn * r4
    Jd[:, 2, 0], Jd[:, 2, 1] = alpha * xn * r6, beta * yn * r6
    Jd[:, 3, 0], Jd[:, 3, 1] = 2.0 * alpha * xn * yn, beta * (r2 + 2.0 * yn * yn)
    Jd[:, 4, 0], Jd[:, 4, 1] = alpha * (r2 + 2.0 * xn * xn), 2.0 * beta * xn * yn
    dr = 2.0 * k1 + 4.0 * k2 * r2 + 6.0 * k3 * r4
    pd = np.zeros((n, 2, 2))
    pd[:, 0, 0] = rad + xn * (dr * xn) + 2.0 * p1 * yn + 6.0 * p2 * xn
    pd[:, 0, 1] = xn * (dr * yn) + 2.0 * p1 * xn + 2.0 * p2 * yn
    pd[:, 1, 0] = yn * (dr * xn) + 2.0 * p1 * xn + 2.0 * p2 * yn
    pd[:, 1, 1] = rad + yn * (dr * yn) + 2.0 * p2 * xn + 6.0 * p1 * yn
    iz = 1.0 / Zp
    pn = np.zeros((n, 2, 3))
    pn[:, 0, 0], pn[:, 0, 2], pn[:, 1, 1], pn[:, 1, 2] = iz, -Xp * iz * iz, iz, -Yp * iz * iz
    PP = np.zeros((n, 3, 6))
    PP[:, :, :3] = np.eye(3)
    PP[:, :, 3:] = -lie.hat(P)
    Jp = np.einsum("ij,njk,nkl,nlm->nim", np.diag([alpha, beta]), pd, pn, PP).transpose(0, 2, 1)
    return e, Ji, Jd, Jp


def total_optimization(K4, poses, objs, imgs, max_iterations=10, tol=1e-8):
    """Returns (K4, D5, poses, info) with info = dict(update_norms, costs, iterations)."""
    V = len(poses)
    param = np.concatenate([np.asarray(K4, dtype=np.float64), np.zeros(5), np.asarray(poses, dtype=np.float64).ravel()])
    norms, costs = [], []
    for _ in range(max_iterations):
        H = np.zeros((9 + 6 * V, 9 + 6 * V))
        g = np.zeros(9 + 6 * V)
        cost = 0.0
        for i in range(V):
            e, Ji, Jd, Jp = residual_jacobian(param[:4], param[4:9], param[9 + 6 * i:15 + 6 * i], objs[i], imgs[i])
            J = np.concatenate([Ji, Jd, Jp], axis=1)            # [n, 15, 2]
            idx = np.concatenate([np.arange(9), 9 + 6 * i + np.arange(6)])
            H[np.ix_(idx, idx)] += np.einsum("nak,nbk->ab", J, J)
            g[idx] -= np.einsum("nak,nk->a", J, e)
            cost += 0.5 * float(np.sum(e * e))
        costs.append(cost)
        update = np.linalg.solve(H, g)
        param[:9] += update[:9]
        for i in range(V):
            Rd, td = se3_exp(update[9 + 6 * i:15 + 6 * i])
            Rp, tp = se3_exp(param[9 + 6 * i:15 + 6 * i])
            param[9 + 6 * i:15 + 6 * i] = se3_log(Rd @ Rp, Rd @ tp + td)
        norms.append(float(np.linalg.norm(update)))
        if norms[-1] < tol:
            break
    return param[:4].copy(), param[4:9].copy(), param[9:].reshape(V, 6).copy(), dict(update_norms=norms, costs=costs, iterations=len(norms))


def solve(objs, imgs):
    """CalibSolver::solve, calib.cpp:38-47."""
    Hs = [homography(im, ob) for im, ob in zip(imgs, objs)]
    K4 = intrinsics_from_homographies(Hs)
    poses = extrinsics(K4, Hs)
    return (K4, poses, Hs) + total_optimization(K4, poses, objs, imgs)


def synthetic_views(n_views=20, rows=8, cols=11, cb_size=2.8e-2, seed=20221107, noise_px=0.1):
    """BASELINE.json configs[3] asks for 20 views x 88 corners; the reference ships 9 x 40 only, so this
    synthesises the larger case from a known camera (pinhole + the reference's distortion model)."""
    rng = np.random.default_rng(seed)
    K4 = np.array([3200.0, 3180.0, 2010.0, 1490.0])
    D5 = np.array([0.08, -0.15, 0.05, 1e-3, -8e-4])
    obj = object_points(rows, cols, cb_size)
    objs, imgs, poses = [], [], []
    centre = np.array([0.5 * (cols - 1) * cb_size, 0.5 * (rows - 1) * cb_size, 0.0])
    for _ in range(n_views):
        om = rng.normal(0, 0.25, 3)
        R = so3_exp_matrix(om)
        t = np.array([rng.normal(0, 0.03), rng.normal(0, 0.03), 0.55 + rng.uniform(-0.1, 0.15)]) - R @ centre
        P = (R @ np.stack([obj[:, 0], obj[:, 1], np.zeros(len(obj))])).T + t
        xn, yn = P[:, 0] / P[:, 2], P[:, 1] / P[:, 2]
        r2 = xn * xn + yn * yn
        rad = 1 + D5[0] * r2 + D5[1] * r2 ** 2 + D5[2] * r2 ** 3
        xd = xn * rad + 2 * D5[3] * xn * yn + D5[4] * (r2 + 2 * xn * xn)
        yd = yn * rad + 2 * D5[4] * xn * yn + D5[3] * (r2 + 2 * yn * yn)
        uv = np.stack([K4[0] * xd + K4[2], K4[1] * yd + K4[3]], axis=-1) + rng.normal(0, noise_px, (len(obj), 2))
        objs.append(obj.copy())
        imgs.append(np.round(uv, 3).astype(np.float32).astype(np.float64))    # written with 3 decimals, read with stof
        poses.append(se3_log(R, t))
    return objs, imgs, K4, D5, np.array(poses)
