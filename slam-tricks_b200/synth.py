"""Deterministic synthetic bundle-adjustment scenes (host side, NumPy).

Restates the reference's `ProblemScene` generator
(`st20-g2o/src/src/sim_data.cpp:22-142, 244-296`) at arbitrary size and with fixed
seeds (the original is clock-seeded, `sim_data.cpp:273`, and needs the missing
`slam-scene-viewer` submodule for `CubePlane::GenerateFeatures`, `sim_data.cpp:36`):

* cameras    — the spherical spiral of `CreateTrajectory` (`sim_data.cpp:47-96`):
               s in [0,290), z = -2.9 + 0.02 s, phi = 10 deg * s, radius 3, optical axis
               through the origin; `n_cam = 29` reproduces the reference's every-10th sampling.
* landmarks  — six faces of the cube |coord| = 5 in the order of `CreateScene`
               (`sim_data.cpp:23-30`), uniform on each face.
* visibility — `CreateMeasurements` (`sim_data.cpp:119-142`): z >= 0, |x/z| < 0.8, |y/z| < 0.6.
* thinning   — to hit an exact observation budget each landmark keeps a random subset of
               the cameras that see it (not in the reference: it keeps them all).
* noise      — observation noise N(0, sigma_uv) (the reference adds none); initial guess as
               `Simulation` (`sim_data.cpp:273-296`): R*Rz(a)Ry(b)Rx(c), a,b,c ~ N(0, 3 deg),
               t + N(0, 0.3)^3, first and last camera exact and constant.

Output layout (the SoA the solver takes): cam_q f64[n_cam,4] (xyzw), cam_t f64[n_cam,3],
lm f64[n_lm,3], obs_cam i32[n_obs], obs_lm i32[n_obs], obs_uv f64[n_obs,2], cam_const u8[n_cam];
observations landmark-major, camera-ascending inside a landmark (`test_ceres.h:109-110`).
"""
from dataclasses import dataclass

import numpy as np

SEED_DATA = 20221105
SEED_INIT = 20221106
HALF_W, HALF_H = 0.8, 0.6          # sim_data.h:211-212
_CHUNK = 8192                       # landmarks per RNG/visibility chunk (part of the seed contract)


@dataclass
class BAScene:
    cam_q: np.ndarray
    cam_t: np.ndarray
    lm: np.ndarray
    obs_cam: np.ndarray
    obs_lm: np.ndarray
    obs_uv: np.ndarray
    cam_const: np.ndarray
    true_cam_q: np.ndarray
    true_cam_t: np.ndarray
    true_lm: np.ndarray

    @property
    def n_cam(self):
        return len(self.cam_q)

    @property
    def n_lm(self):
        return len(self.lm)

    @property
    def n_obs(self):
        return len(self.obs_cam)


def _quat_from_rot(R):
    """Batched rotation matrix -> unit quaternion xyzw, w >= 0."""
    R = np.asarray(R, dtype=np.float64)
    q = np.empty(R.shape[:-2] + (4,))
    for i in range(len(R)):
        m = R[i]
        tr = m[0, 0] + m[1, 1] + m[2, 2]
        if tr > 0:
            s = np.sqrt(tr + 1.0) * 2
            qi = [(m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s, 0.25 * s]
        else:
            a = int(np.argmax([m[0, 0], m[1, 1], m[2, 2]]))
            b, c = (a + 1) % 3, (a + 2) % 3
            s = np.sqrt(1.0 + m[a, a] - m[b, b] - m[c, c]) * 2
            qi = [0.0, 0.0, 0.0, (m[c, b] - m[b, c]) / s]
            qi[a] = 0.25 * s
            qi[b] = (m[b, a] + m[a, b]) / s
            qi[c] = (m[c, a] + m[a, c]) / s
        qi = np.array(qi)
        if qi[3] < 0:
            qi = -qi
        q[i] = qi / np.linalg.norm(qi)
    return q


def _rot_from_quat(q):
    x, y, z, w = np.moveaxis(q, -1, 0)
    R = np.empty(x.shape + (3, 3))
    R[..., 0, 0] = 1 - 2 * (y * y + z * z)
    R[..., 0, 1] = 2 * (x * y - z * w)
    R[..., 0, 2] = 2 * (x * z + y * w)
    R[..., 1, 0] = 2 * (x * y + z * w)
    R[..., 1, 1] = 1 - 2 * (x * x + z * z)
    R[..., 1, 2] = 2 * (y * z - x * w)
    R[..., 2, 0] = 2 * (x * z - y * w)
    R[..., 2, 1] = 2 * (y * z + x * w)
    R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def spiral_cameras(n_cam):
    """`CreateTrajectory` (sim_data.cpp:47-96) sampled at n_cam spiral parameters s = i*290/n_cam."""
    s = np.arange(n_cam) * (290.0 / n_cam)
    z = -2.9 + 0.02 * s
    phi = np.deg2rad(10.0 * s)
    c = np.sqrt(9.0 - z * z) / 3.0
    pos = np.stack([3.0 * c * np.cos(phi), 3.0 * c * np.sin(phi), z], axis=-1)
    xa = np.stack([-pos[:, 1], pos[:, 0], np.zeros(n_cam)], axis=-1)
    xa /= np.linalg.norm(xa, axis=-1, keepdims=True)
    za = -pos / np.linalg.norm(pos, axis=-1, keepdims=True)
    ya = np.cross(za, xa)
    R = np.stack([xa, ya, za], axis=-1)            # columns = axes (sim_data.cpp:75-77)
    return R, pos


def _face_points(idx, ab):
    """Candidate landmark `idx` lies on cube face idx % 6 (order of sim_data.cpp:23-30)."""
    face = idx % 6
    p = np.empty((len(idx), 3))
    a, b = ab[:, 0], ab[:, 1]
    for f, (axis, sign) in enumerate([(1, 5.0), (1, -5.0), (0, 5.0), (0, -5.0), (2, 5.0), (2, -5.0)]):
        m = face == f
        others = [k for k in range(3) if k != axis]
        p[m, axis] = sign
        p[m, others[0]] = a[m]
        p[m, others[1]] = b[m]
    return p


def visibility(R, pos, pts):
    """`CreateMeasurements` predicate (sim_data.cpp:124-134) for every (landmark, camera):
    returns bool[n_pts, n_cam] and the exact projections (x/z, y/z)."""
    d = pts[:, None, :] - pos[None, :, :]
    pc = np.einsum("cji,pcj->pci", R, d)
    z = pc[..., 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        u, v = pc[..., 0] / z, pc[..., 1] / z
    vis = (z >= 0.0) & (np.abs(u) < HALF_W) & (np.abs(v) < HALF_H)
    return vis, u, v


def make_scene(n_cam, n_lm, n_obs, sigma_uv=1e-3, pos_noise=0.3, angle_noise_deg=3.0,
               lm_noise=0.05, seed=SEED_DATA, seed_init=SEED_INIT):
    """Build a scene with EXACTLY (n_cam, n_lm, n_obs)."""
    rng = np.random.default_rng(seed)
    R, pos = spiral_cameras(n_cam)
    k_max = -(-n_obs // n_lm)

    # -- landmarks: rejection-sample candidates seen by >= 2 cameras, chunked
    pts, viss = [], []
    cand = 0
    have = 0
    while have < n_lm:
        idx = np.arange(cand, cand + _CHUNK)
        cand += _CHUNK
        p = _face_points(idx, rng.uniform(-5.0, 5.0, size=(_CHUNK, 2)))
        vis, _, _ = visibility(R, pos, p)
        ok = vis.sum(axis=1) >= 2
        pts.append(p[ok])
        viss.append(np.packbits(vis[ok], axis=1))
        have += int(ok.sum())
    pts = np.concatenate(pts)[:n_lm]
    vis_bits = np.concatenate(viss)[:n_lm]
    nvis = np.zeros(n_lm, dtype=np.int64)
    for s in range(0, n_lm, _CHUNK):
        nvis[s:s + _CHUNK] = np.unpackbits(vis_bits[s:s + _CHUNK], axis=1, count=n_cam).sum(axis=1)

    # -- degrees: min(visible, k_max), then fix the total to n_obs exactly
    deg = np.minimum(nvis, k_max)
    total = int(deg.sum())
    if total > n_obs:                      # shave from the tail, never below 2
        for l in range(n_lm - 1, -1, -1):
            take = min(total - n_obs, int(deg[l]) - 2)
            deg[l] -= take
            total -= take
            if total == n_obs:
                break
    while total < n_obs:                   # top up landmarks that have spare visibility
        spare = np.nonzero(deg < nvis)[0]
        if len(spare) == 0:
            raise ValueError("scene cannot reach n_obs=%d (max %d)" % (n_obs, total))
        spare = spare[: n_obs - total]
        deg[spare] += 1
        total += len(spare)
    assert total == n_obs and deg.min() >= 2

    # -- per landmark: choose deg[l] of its visible cameras (random keys, smallest first)
    obs_cam = np.empty(n_obs, dtype=np.int32)
    obs_lm = np.repeat(np.arange(n_lm, dtype=np.int32), deg)
    ptr = np.concatenate([[0], np.cumsum(deg)])
    for s in range(0, n_lm, _CHUNK):
        e = min(s + _CHUNK, n_lm)
        vis = np.unpackbits(vis_bits[s:e], axis=1, count=n_cam).astype(bool)
        key = rng.random(size=vis.shape)
        key[~vis] = 2.0
        order = np.argsort(key, axis=1, kind="stable")
        for l in range(s, e):
            obs_cam[ptr[l]:ptr[l + 1]] = np.sort(order[l - s, :deg[l]])

    # -- observations: exact projection + noise
    d = pts[obs_lm] - pos[obs_cam]
    pc = np.einsum("nji,nj->ni", R[obs_cam], d)
    uv = np.stack([pc[:, 0] / pc[:, 2], pc[:, 1] / pc[:, 2]], axis=-1)
    if sigma_uv > 0:
        uv = uv + rng.normal(0.0, sigma_uv, size=uv.shape)

    # -- initial guess (sim_data.cpp:273-296)
    rng2 = np.random.default_rng(seed_init)
    true_q = _quat_from_rot(R)
    ang = rng2.normal(0.0, np.deg2rad(angle_noise_deg), size=(n_cam, 3))
    dpos = rng2.normal(0.0, pos_noise, size=(n_cam, 3))
    ca, sa = np.cos(ang), np.sin(ang)
    Rz = np.zeros((n_cam, 3, 3)); Ry = np.zeros((n_cam, 3, 3)); Rx = np.zeros((n_cam, 3, 3))
    Rz[:, 0, 0] = ca[:, 0]; Rz[:, 0, 1] = -sa[:, 0]; Rz[:, 1, 0] = sa[:, 0]; Rz[:, 1, 1] = ca[:, 0]; Rz[:, 2, 2] = 1
    Ry[:, 0, 0] = ca[:, 1]; Ry[:, 0, 2] = sa[:, 1]; Ry[:, 2, 0] = -sa[:, 1]; Ry[:, 2, 2] = ca[:, 1]; Ry[:, 1, 1] = 1
    Rx[:, 1, 1] = ca[:, 2]; Rx[:, 1, 2] = -sa[:, 2]; Rx[:, 2, 1] = sa[:, 2]; Rx[:, 2, 2] = ca[:, 2]; Rx[:, 0, 0] = 1
    Rn = R @ Rz @ Ry @ Rx
    cam_q = _quat_from_rot(Rn)
    cam_t = pos + dpos
    cam_q[0], cam_q[-1] = true_q[0], true_q[-1]
    cam_t[0], cam_t[-1] = pos[0], pos[-1]
    cam_const = np.zeros(n_cam, dtype=np.uint8)
    cam_const[0] = cam_const[-1] = 1
    lm0 = pts + rng2.normal(0.0, lm_noise, size=pts.shape)
    return BAScene(cam_q=cam_q, cam_t=cam_t.copy(), lm=lm0, obs_cam=obs_cam, obs_lm=obs_lm,
                   obs_uv=np.ascontiguousarray(uv), cam_const=cam_const,
                   true_cam_q=true_q, true_cam_t=pos.copy(), true_lm=pts)


def replicate(scene, times):
    """`times` independent copies of a scene side by side (cameras, landmarks and observations
    all replicated, indices shifted) — used to scale the observation stream past L2."""
    nc, nl = scene.n_cam, scene.n_lm
    k = np.arange(times)
    return BAScene(
        cam_q=np.tile(scene.cam_q, (times, 1)), cam_t=np.tile(scene.cam_t, (times, 1)),
        lm=np.tile(scene.lm, (times, 1)),
        obs_cam=(scene.obs_cam[None, :] + (k * nc)[:, None]).astype(np.int32).ravel(),
        obs_lm=(scene.obs_lm[None, :] + (k * nl)[:, None]).astype(np.int32).ravel(),
        obs_uv=np.tile(scene.obs_uv, (times, 1)), cam_const=np.tile(scene.cam_const, times),
        true_cam_q=np.tile(scene.true_cam_q, (times, 1)), true_cam_t=np.tile(scene.true_cam_t, (times, 1)),
        true_lm=np.tile(scene.true_lm, (times, 1)))


def _angle_axis(angle_deg, axis):
    a = np.deg2rad(angle_deg)
    x, y, z = axis
    K = np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]], dtype=np.float64)
    return np.eye(3) + np.sin(a) * K + (1 - np.cos(a)) * (K @ K)


def _ypr_rotation(yaw, pitch, roll):
    """`r * p * y` with y about z, p about x, r about y — st17-ceres/src/main.cpp:17-21, scene.cpp:17-20."""
    return _angle_axis(roll, (0, 1, 0)) @ _angle_axis(pitch, (1, 0, 0)) @ _angle_axis(yaw, (0, 0, 1))


def pnp_poses():
    """Ground-truth and initial camera->world poses of the PnP demo (st17-ceres/src/main.cpp:14-35):
    (q_real xyzw, t_real, q_init xyzw, t_init).  These are the poses printed in
    st17-ceres/img/release.png: q_real = (0.40958, 0.70941, -0.49673, -0.28679) up to sign."""
    R_real = _ypr_rotation(-120.0, 110.0, 0.0).T
    R_init = _ypr_rotation(-90.0, 90.0, 10.0).T
    return (_quat_from_rot(R_real[None])[0], np.array([3.0, 2.0, 1.0]),
            _quat_from_rot(R_init[None])[0], np.array([2.5, 0.0, 0.0]))


def pnp_scene(seed=SEED_DATA, features_per_plane=10):
    """The PnP problem of st17-ceres/src/main.cpp:37-87 with a fixed seed (the reference's feature
    positions are clock-seeded, scene.cpp:23): five planes (main.cpp:37-46), `features_per_plane`
    uniform features each (scene.cpp:25-41, float32 like pcl::PointXYZ), kept when they project into
    |x/z| < 1, |y/z| < 0.75, z > 0 of the TRUE camera (main.cpp:74-77).
    Returns dict(points [n,3], uv [n,2], q_real, t_real, q_init, t_init)."""
    rng = np.random.default_rng(seed)
    planes = [(0.0, 0.0, 0.0, -5.0, 0.0, 0.0, 10.0, 4.5), (0.0, 0.0, 90.0, 0.0, 5.0, 0.0, 10.0, 4.5),
              (0.0, 0.0, 0.0, 5.0, 0.0, 0.0, 10.0, 4.5), (0.0, 0.0, 90.0, 0.0, -5.0, 0.0, 10.0, 4.5),
              (90.0, 0.0, 0.0, 0.0, 0.0, -2.25, 10.0, 10.0)]
    pts = []
    for roll, pitch, yaw, dx, dy, dz, width, height in planes:
        Rp = _ypr_rotation(yaw, pitch, roll)
        local = np.stack([np.zeros(features_per_plane), rng.uniform(-0.5 * width, 0.5 * width, features_per_plane),
                          rng.uniform(-0.5 * height, 0.5 * height, features_per_plane)], axis=-1)
        pts.append((local @ Rp.T + np.array([dx, dy, dz])).astype(np.float32).astype(np.float64))
    pts = np.concatenate(pts)
    q_real, t_real, q_init, t_init = pnp_poses()
    R = _rot_from_quat(q_real)
    pc = (pts - t_real) @ R
    nx, ny = pc[:, 0] / pc[:, 2], pc[:, 1] / pc[:, 2]
    keep = (pc[:, 2] > 0) & (nx > -1.0) & (nx < 1.0) & (ny > -0.75) & (ny < 0.75)
    return dict(points=pts[keep], uv=np.stack([nx[keep], ny[keep]], axis=-1), q_real=q_real, t_real=t_real,
                q_init=q_init, t_init=t_init)


CONFIGS = {
    "ref29": (29, 600, None),          # the reference's own scene size (test_ceres.cpp:8)
    "B": (50, 5000, 50000),            # BASELINE.json configs[1]
    "C": (1000, 100000, 1000000),      # BASELINE.json configs[2]
}


# ------------------------------------------------------------------------------------------------
# SE(3) pose-graph scenes (BASELINE.json configs[4]).  Restates the shape of the reference's
# pose-chain simulator (st4-kalman/src/src/pose_simulation.cpp:17-88: a spiral on the sphere of radius 1 around
# (0,0,1), odometry that drifts) deterministically; same numbers as oracle/pg_oracle.py make_graph for the same
# seed (tests/test_posegraph.py checks that), but self-contained: product code never imports the oracle.
# ------------------------------------------------------------------------------------------------
def _qmul(a, b):
    ax, ay, az, aw = np.moveaxis(a, -1, 0)
    bx, by, bz, bw = np.moveaxis(b, -1, 0)
    return np.stack([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz], axis=-1)


def _qnorm(q):
    return q / np.linalg.norm(q, axis=-1, keepdims=True)


def _qrotate(q, v):
    return np.einsum("...ij,...j->...i", _rot_from_quat(q), v)


def _se3_compose(qa, ta, qb, tb):
    return _qnorm(_qmul(qa, qb)), ta + _qrotate(qa, tb)


def _se3_inverse(q, t):
    qi = q * np.array([-1.0, -1.0, -1.0, 1.0])
    return qi, -_qrotate(qi, t)


def _se3_exp(xi):
    """Sophus::SE3d::exp, tangent order [rho, theta] -> (q xyzw, t)."""
    rho, om = xi[..., :3], xi[..., 3:]
    th2 = np.sum(om * om, axis=-1)
    small = th2 < 1e-20
    th = np.sqrt(np.where(small, 1.0, th2))
    imag = np.where(small, 0.5 - th2 / 48.0 + th2 * th2 / 3840.0, np.sin(0.5 * th) / th)
    real = np.where(small, 1.0 - th2 / 8.0 + th2 * th2 / 384.0, np.cos(0.5 * th))
    q = np.concatenate([imag[..., None] * om, real[..., None]], axis=-1)
    theta = np.linalg.norm(om, axis=-1)
    sm = theta < 1e-10
    tt = np.where(sm, 1.0, theta)
    a = np.where(sm, 0.5, (1 - np.cos(tt)) / tt ** 2)
    b = np.where(sm, 1.0 / 6.0, (tt - np.sin(tt)) / tt ** 3)
    c1 = np.cross(om, rho)
    c2 = np.cross(om, c1)
    return q, rho + a[..., None] * c1 + b[..., None] * c2


def _relative(q, t, ei, ej):
    qi, ti = _se3_inverse(q[ei], t[ei])
    return _se3_compose(qi, ti, q[ej], t[ej])


def pose_graph(n=200, offsets=(1, 2, 3, 4), sigma_t=0.01, sigma_r=0.005, drift_t=0.02, drift_r=0.01, seed=20221108, turns=10.0, closures=0,
               closure_min=17):
    """Returns dict(q0, t0, ei, ej, zq, zt, q_truth, t_truth): truth = spiral, measurements = noisy relative poses on
    a band of index offsets (edges sorted by (i, j)), initial guess = integration of noisier (i, i+1) steps; `closures`
    extra edges (i, j) with j - i >= closure_min appended behind the band edges (loop closures)."""
    rng = np.random.default_rng(seed)
    s = np.arange(n) / max(n - 1, 1)
    z = 2.0 * s
    r = np.sqrt(np.maximum(1.0 - (z - 1.0) ** 2, 1e-6))
    th = 2 * np.pi * turns * s
    pos = np.stack([r * np.cos(th), r * np.sin(th), z], axis=-1)
    zc = np.array([0.0, 0.0, 1.0]) - pos
    zc /= np.linalg.norm(zc, axis=-1, keepdims=True)
    xa = np.stack([-np.sin(th), np.cos(th), np.zeros(n)], axis=-1)
    xa -= np.sum(xa * zc, axis=-1, keepdims=True) * zc
    xa /= np.linalg.norm(xa, axis=-1, keepdims=True)
    ya = np.cross(zc, xa)
    qT, tT = _quat_from_rot(np.stack([xa, ya, zc], axis=-1)), pos
    ei = np.concatenate([np.arange(0, n - o) for o in offsets]).astype(np.int32)
    ej = np.concatenate([np.arange(o, n) for o in offsets]).astype(np.int32)
    order = np.lexsort((ej, ei))
    ei, ej = ei[order], ej[order]

    def measure(a, b, st, sr):
        zq, zt = _relative(qT, tT, a, b)
        noise = np.concatenate([rng.normal(0, st, (len(a), 3)), rng.normal(0, sr, (len(a), 3))], axis=1)
        return _se3_compose(zq, zt, *_se3_exp(noise))

    zq, zt = measure(ei, ej, sigma_t, sigma_r)
    sq, st = measure(np.arange(n - 1), np.arange(1, n), drift_t, drift_r)
    q0, t0 = np.zeros((n, 4)), np.zeros((n, 3))
    q0[0], t0[0] = qT[0], tT[0]
    for i in range(1, n):
        q0[i], t0[i] = _se3_compose(q0[i - 1], t0[i - 1], sq[i - 1], st[i - 1])
    if closures:
        ci = rng.integers(0, n - closure_min, closures)
        cj = np.array([rng.integers(a + closure_min, n) for a in ci])
        cq, ct = measure(ci, cj, sigma_t, sigma_r)
        ei = np.concatenate([ei, ci.astype(np.int32)]); ej = np.concatenate([ej, cj.astype(np.int32)])
        zq = np.concatenate([zq, cq]); zt = np.concatenate([zt, ct])
    return dict(q0=q0, t0=t0, ei=ei, ej=ej, zq=zq, zt=zt, q_truth=qT, t_truth=tT)


def calib_views(n_views=20, rows=8, cols=11, cb_size=2.8e-2, seed=20221107, noise_px=0.1):
    """Synthetic Zhang-calibration input of BASELINE.json configs[3] (20 views x 88 corners; the reference ships 9 x 40
    only): a known pinhole camera with the reference's distortion model (st3-calibration/src/src/calib.cpp:254-262),
    corner (i, j) at board coordinates (j, i) * cb_size (calib.cpp:11-36), pixels rounded to 3 decimals and through
    float32 as `CBCorners::write` / `read` do (cbcorner.cpp:34-72).  Returns (objs, imgs, K4, D5)."""
    rng = np.random.default_rng(seed)
    K4 = np.array([3200.0, 3180.0, 2010.0, 1490.0])
    D5 = np.array([0.08, -0.15, 0.05, 1e-3, -8e-4])
    j, i = np.meshgrid(np.arange(cols), np.arange(rows))
    obj = np.stack([j.ravel() * cb_size, i.ravel() * cb_size], axis=-1).astype(np.float64)
    centre = np.array([0.5 * (cols - 1) * cb_size, 0.5 * (rows - 1) * cb_size, 0.0])
    objs, imgs = [], []
    for _ in range(n_views):
        om = rng.normal(0, 0.25, 3)
        th = np.linalg.norm(om)
        Kx = np.array([[0, -om[2], om[1]], [om[2], 0, -om[0]], [-om[1], om[0], 0]])
        R = np.eye(3) + (np.sin(th) / th) * Kx + ((1 - np.cos(th)) / th ** 2) * (Kx @ Kx) if th > 1e-12 else np.eye(3)
        t = np.array([rng.normal(0, 0.03), rng.normal(0, 0.03), 0.55 + rng.uniform(-0.1, 0.15)]) - R @ centre
        P = (R @ np.stack([obj[:, 0], obj[:, 1], np.zeros(len(obj))])).T + t
        xn, yn = P[:, 0] / P[:, 2], P[:, 1] / P[:, 2]
        r2 = xn * xn + yn * yn
        rad = 1 + D5[0] * r2 + D5[1] * r2 ** 2 + D5[2] * r2 ** 3
        xd = xn * rad + 2 * D5[3] * xn * yn + D5[4] * (r2 + 2 * xn * xn)
        yd = yn * rad + 2 * D5[4] * xn * yn + D5[3] * (r2 + 2 * yn * yn)
        uv = np.stack([K4[0] * xd + K4[2], K4[1] * yd + K4[3]], axis=-1) + rng.normal(0, noise_px, (len(obj), 2))
        objs.append(obj.copy())
        imgs.append(np.round(uv, 3).astype(np.float32).astype(np.float64))
    return objs, imgs, K4, D5
