"""NumPy oracle of the SE(3) pose-graph path (BASELINE.json configs[4]).  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED BY CONSTRUCTION: the reference contains no pose-graph residual and no pose-graph solver.
What it does contain, and what this file follows:

* the pose-chain simulator `simulation`, st4-kalman/src/src/pose_simulation.cpp:17-88 (spiral on a sphere,
  odometry noise accumulated by left-multiplication) and its recorded output
  st4-kalman/output/{truth,obs}.csv (1000 poses, `x,y,z,qx,qy,qz,qw`) — the input fixture;
* the metric `absTrajectoryError`, pose_simulation.cpp:198-209: sqrt(mean |Log(T_truth^-1 T_est)|^2);
* the SE(3) right-perturbation / adjoint / Jacobian identities of st23-lie-group-v2/doc.tex:862-997 and
  st21-lie/lie-group.tex:218-278, used for the analytic Jacobians;
* Sophus conventions (tangent order [rho, theta], `SE3::exp/log`) as everywhere else in the reference.

Problem: poses T_i = (q_i xyzw, t_i); edge (i, j, Z_ij): r_ij = Log(Z_ij^-1 T_i^-1 T_j) in R^6; manifold
T <- T Exp(delta) (right perturbation); pose 0 constant (gauge).  Jacobians (exact):
    dr/d delta_j = J_r^-1(r),      dr/d delta_i = -J_r^-1(r) Ad(T_j^-1 T_i).
Solver: the same Ceres-faithful trust-region LM as `ba_oracle.solve` with an exact
sparse solve of the damped normal equations (Ceres: SPARSE_NORMAL_CHOLESKY).
"""
import numpy as np
import scipy.sparse
import scipy.sparse.linalg

from . import lie
from .ba_oracle import LMOptions, LMSummary

EPS = 1e-10          # Sophus epsilon for exp/log branches
SMALL = 1e-5         # |theta| below which the Jacobian series are used


# ------------------------------------------------------------------ SE(3), batched, (q xyzw, t)
def quat_rotate(q, v):
    return np.einsum("...ij,...j->...i", lie.quat_to_rot(q), v)


def quat_conj(q):
    return q * np.array([-1.0, -1.0, -1.0, 1.0])


def compose(qa, ta, qb, tb):
    return lie.quat_normalize(lie.quat_mul(qa, qb)), ta + quat_rotate(qa, tb)


def inverse(q, t):
    qi = quat_conj(q)
    return qi, -quat_rotate(qi, t)


def _V_coeffs(theta):
    small = theta < EPS
    th = np.where(small, 1.0, theta)
    a = np.where(small, 0.5, (1 - np.cos(th)) / th ** 2)
    b = np.where(small, 1.0 / 6.0, (th - np.sin(th)) / th ** 3)
    return a, b


def se3_exp(xi):
    """Sophus::SE3d::exp: [rho, theta] -> (q, t = V rho)."""
    xi = np.asarray(xi, dtype=np.float64)
    rho, om = xi[..., :3], xi[..., 3:]
    theta = np.linalg.norm(om, axis=-1)
    a, b = _V_coeffs(theta)
    c1 = np.cross(om, rho)
    c2 = np.cross(om, c1)
    return lie.so3_exp_quat(om), rho + a[..., None] * c1 + b[..., None] * c2


def se3_log(q, t):
    """Sophus::SE3d::log."""
    om = lie.so3_log_quat(q)
    theta = np.linalg.norm(om, axis=-1)
    small = theta < EPS
    th = np.where(small, 1.0, theta)
    half = 0.5 * th
    c = np.where(small, 1.0 / 12.0, (1 - th * np.cos(half) / (2 * np.sin(half))) / th ** 2)
    c1 = np.cross(om, t)
    c2 = np.cross(om, c1)
    return np.concatenate([t - 0.5 * c1 + c[..., None] * c2, om], axis=-1)


def adjoint(q, t):
    """Ad(T) for the tangent order [rho, theta]: [[R, t^ R], [0, R]]."""
    R = lie.quat_to_rot(q)
    A = np.zeros(R.shape[:-2] + (6, 6))
    A[..., :3, :3] = R
    A[..., 3:, 3:] = R
    A[..., :3, 3:] = lie.hat(t) @ R
    return A


def jl_inv_se3(xi):
    """Inverse LEFT Jacobian of SE(3), order [rho, theta]: [[A, -A Q A], [0, A]], A = J_l(theta)^-1 (SO(3)),
    Q(rho, theta) as in Barfoot, State Estimation for Robotics, eq. (7.86)."""
    xi = np.asarray(xi, dtype=np.float64)
    rho, om = xi[..., :3], xi[..., 3:]
    th = np.linalg.norm(om, axis=-1)
    small = th < SMALL
    t = np.where(small, 1.0, th)
    s, c = np.sin(t), np.cos(t)
    W, P = lie.hat(om), lie.hat(rho)
    WW = W @ W
    I = np.eye(3)
    ca = np.where(small, 1.0 / 12.0, 1.0 / t ** 2 - (1 + c) / (2 * t * s))
    A = I - 0.5 * W + ca[..., None, None] * WW
    q1 = np.where(small, 1.0 / 6.0, (t - s) / t ** 3)
    q2 = np.where(small, 1.0 / 24.0, (t * t + 2 * c - 2) / (2 * t ** 4))
    q3 = np.where(small, 1.0 / 120.0, (2 * t - 3 * s + t * c) / (2 * t ** 5))
    WP, PW = W @ P, P @ W
    WPW = WP @ W
    Q = (0.5 * P + q1[..., None, None] * (WP + PW + WPW) + q2[..., None, None] * (W @ WP + PW @ W - 3 * WPW)
         + q3[..., None, None] * (WPW @ W + W @ WPW))
    J = np.zeros(xi.shape[:-1] + (6, 6))
    J[..., :3, :3] = A
    J[..., 3:, 3:] = A
    J[..., :3, 3:] = -A @ Q @ A
    return J


def jr_inv_se3(xi):
    return jl_inv_se3(-np.asarray(xi, dtype=np.float64))


# ------------------------------------------------------------------ residual / Jacobians
def relative(q, t, ei, ej):
    qi_inv, ti_inv = inverse(q[ei], t[ei])
    return compose(qi_inv, ti_inv, q[ej], t[ej])


def residuals(q, t, ei, ej, zq, zt):
    qz, tz = inverse(zq, zt)
    qr, tr = relative(q, t, ei, ej)
    qe, te = compose(qz, tz, qr, tr)
    return se3_log(qe, te)


def residual_jacobians(q, t, ei, ej, zq, zt):
    r = residuals(q, t, ei, ej, zq, zt)
    Jj = jr_inv_se3(r)
    qji, tji = relative(q, t, ej, ei)                    # T_j^-1 T_i
    Ji = -Jj @ adjoint(qji, tji)
    return r, Ji, Jj


def plus(q, t, delta):
    dq, dt = se3_exp(delta)
    return compose(q, t, dq, dt)


def ate(q_truth, t_truth, q, t):
    """absTrajectoryError, st4-kalman/src/src/pose_simulation.cpp:198-209."""
    qi, ti = inverse(q_truth, t_truth)
    qe, te = compose(qi, ti, q, t)
    xi = se3_log(qe, te)
    return float(np.sqrt(np.mean(np.sum(xi * xi, axis=-1))))


# ------------------------------------------------------------------ problems
def measurements_from(q_truth, t_truth, ei, ej, sigma_t=0.0, sigma_r=0.0, rng=None):
    """Z_ij = T_i^-1 T_j of the truth, right-multiplied by Exp(noise)."""
    zq, zt = relative(q_truth, t_truth, ei, ej)
    if rng is not None and (sigma_t > 0 or sigma_r > 0):
        noise = np.concatenate([rng.normal(0, sigma_t, (len(ei), 3)), rng.normal(0, sigma_r, (len(ei), 3))], axis=1)
        nq, nt = se3_exp(noise)
        zq, zt = compose(zq, zt, nq, nt)
    return zq, zt


def band_edges(n, offsets=(1, 2, 3, 4)):
    ei = np.concatenate([np.arange(0, n - o) for o in offsets]).astype(np.int32)
    ej = np.concatenate([np.arange(o, n) for o in offsets]).astype(np.int32)
    order = np.lexsort((ej, ei))
    return ei[order], ej[order]


def spiral_truth(n, turns=10.0):
    """Poses on the sphere spiral of pose_simulation.cpp:19-48 (radius 1, centre (0,0,1)), z axis towards the centre."""
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
    R = np.stack([xa, ya, zc], axis=-1)
    return np.array([lie.rot_to_quat(Ri) for Ri in R]), pos


def make_graph(n=200, offsets=(1, 2, 3, 4), sigma_t=0.01, sigma_r=0.005, drift_t=0.02, drift_r=0.01, seed=20221108, closures=0,
               closure_min=17):
    """Synthetic pose graph: truth = spiral, measurements = noisy relative poses on a band of offsets, initial guess =
    odometry integration of noisier (i, i+1) steps (as the `obs` track of the reference's simulator drifts).
    `closures` extra edges (i, j), j - i >= closure_min, drawn after everything else (closures = 0 reproduces the
    graphs of the earlier fixtures bit for bit): loop closures between poses far apart along the chain."""
    rng = np.random.default_rng(seed)
    qT, tT = spiral_truth(n)
    ei, ej = band_edges(n, offsets)
    zq, zt = measurements_from(qT, tT, ei, ej, sigma_t, sigma_r, rng)
    q0, t0 = np.zeros((n, 4)), np.zeros((n, 3))
    q0[0], t0[0] = qT[0], tT[0]
    sq, st = measurements_from(qT, tT, np.arange(n - 1), np.arange(1, n), drift_t, drift_r, rng)
    for i in range(1, n):
        q0[i], t0[i] = compose(q0[i - 1], t0[i - 1], sq[i - 1], st[i - 1])
    if closures:
        ci = rng.integers(0, n - closure_min, closures)
        cj = np.array([rng.integers(a + closure_min, n) for a in ci])
        cq, ct = measurements_from(qT, tT, ci, cj, sigma_t, sigma_r, rng)
        ei = np.concatenate([ei, ci.astype(np.int32)]); ej = np.concatenate([ej, cj.astype(np.int32)])
        zq = np.concatenate([zq, cq]); zt = np.concatenate([zt, ct])
    return dict(q0=q0, t0=t0, ei=ei, ej=ej, zq=zq, zt=zt, q_truth=qT, t_truth=tT)


# ------------------------------------------------------------------ Ceres-faithful LM
def solve(q0, t0, ei, ej, zq, zt, options=None):
    """Returns (q, t, LMSummary).  Pose 0 is constant."""
    opt = options or LMOptions()
    q, t = np.array(q0, dtype=np.float64), np.array(t0, dtype=np.float64)
    ei, ej = np.asarray(ei, dtype=np.int64), np.asarray(ej, dtype=np.int64)
    n, m = len(q), len(ei)
    nv = 6 * (n - 1)
    summ = LMSummary()

    def evaluate(q, t):
        r, Ji, Jj = residual_jacobians(q, t, ei, ej, zq, zt)
        rows = (6 * np.arange(m)[:, None, None] + np.arange(6)[None, :, None] + np.zeros((1, 1, 6), dtype=np.int64))
        def block(J, e):
            keep = e > 0
            cols = 6 * (e[:, None, None] - 1) + np.arange(6)[None, None, :] + np.zeros((1, 6, 1), dtype=np.int64)
            return rows[keep].ravel(), cols[keep].ravel(), J[keep].ravel()
        ri, ci, vi = block(Ji, ei)
        rj, cj, vj = block(Jj, ej)
        J = scipy.sparse.csr_matrix((np.concatenate([vi, vj]), (np.concatenate([ri, rj]), np.concatenate([ci, cj]))), shape=(6 * m, nv))
        rr = r.ravel()
        return 0.5 * float(rr @ rr), rr, J, J.T @ rr

    def apply(q, t, delta):
        d = np.zeros((n, 6))
        d[1:] = delta.reshape(n - 1, 6)
        qn, tn = plus(q, t, d)
        qn[0], tn[0] = q[0], t[0]
        return qn, tn

    def ambient_diff(a, b):
        d = np.concatenate([(a[0] - b[0])[1:].ravel(), (a[1] - b[1])[1:].ravel()])
        return float(np.linalg.norm(d)), float(np.max(np.abs(d))) if len(d) else 0.0

    def grad_norms(q, t, g):
        return ambient_diff((q, t), apply(q, t, -g))

    x_cost, r, J, g = evaluate(q, t)
    scale = 1.0 / (1.0 + np.sqrt(np.asarray(J.multiply(J).sum(axis=0)).ravel())) if opt.jacobi_scaling else np.ones(nv)
    gnorm, gmax = grad_norms(q, t, g)
    x_norm = float(np.sqrt(np.sum(q[1:] ** 2) + np.sum(t[1:] ** 2)))
    summ.initial_cost = x_cost
    radius, decrease_factor, reuse_diagonal, diag, num_invalid = opt.initial_trust_region_radius, 2.0, False, None, 0
    it = dict(iteration=0, cost=x_cost, cost_change=0.0, gradient_max_norm=gmax, gradient_norm=gnorm, step_norm=0.0,
              relative_decrease=0.0, trust_region_radius=radius, step_is_valid=True, step_is_successful=True)
    while True:
        if it["step_is_successful"]:
            summ.num_successful_steps += 1
        else:
            summ.num_unsuccessful_steps += 1
        it["trust_region_radius"] = radius
        summ.iterations.append(it)
        if it["iteration"] >= opt.max_num_iterations:
            summ.termination_type, summ.message = "NO_CONVERGENCE", "Maximum number of iterations reached."
            break
        if it["step_is_successful"] and it["gradient_max_norm"] <= opt.gradient_tolerance:
            summ.termination_type, summ.message = "CONVERGENCE", "Gradient tolerance reached."
            break
        if radius < opt.min_trust_region_radius:
            summ.termination_type, summ.message = "CONVERGENCE", "Minimum trust region radius reached."
            break
        prev = it
        it = dict(iteration=prev["iteration"] + 1, cost=x_cost, cost_change=0.0, gradient_max_norm=prev["gradient_max_norm"],
                  gradient_norm=prev["gradient_norm"], step_norm=0.0, relative_decrease=0.0, trust_region_radius=radius,
                  step_is_valid=False, step_is_successful=False)
        Js = J @ scipy.sparse.diags(scale)
        Hs = (Js.T @ Js).tocsc()
        gs = Js.T @ r
        if not reuse_diagonal:
            diag = np.clip(Hs.diagonal(), opt.min_lm_diagonal, opt.max_lm_diagonal)
        valid = True
        try:
            ys = scipy.sparse.linalg.spsolve((Hs + scipy.sparse.diags(diag / radius)).tocsc(), gs)
            valid = bool(np.all(np.isfinite(ys)))
        except Exception:
            valid = False
        reuse_diagonal = True
        if valid:
            step = -ys
            Jd = Js @ step
            model_cost_change = -float(Jd @ (r + 0.5 * Jd))
            valid = model_cost_change > 0.0
        if not valid:
            num_invalid += 1
            if num_invalid >= opt.max_num_consecutive_invalid_steps:
                summ.termination_type = "FAILURE"
                summ.message = "Number of consecutive invalid steps more than Solver::Options::max_num_consecutive_invalid_steps"
                break
            radius /= decrease_factor
            decrease_factor *= 2.0
            reuse_diagonal = False
            continue
        num_invalid = 0
        it["step_is_valid"] = True
        qc, tc = apply(q, t, step * scale)
        rc = residuals(qc, tc, ei, ej, zq, zt).ravel()
        cand_cost = 0.5 * float(rc @ rc)
        cand_ok = np.isfinite(cand_cost)
        step_norm, _ = ambient_diff((qc, tc), (q, t))
        it["step_norm"] = step_norm
        if step_norm <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance):
            summ.termination_type, summ.message = "CONVERGENCE", "Parameter tolerance reached."
            break
        if cand_ok:
            it["cost_change"] = x_cost - cand_cost
            if abs(it["cost_change"]) <= opt.function_tolerance * x_cost:
                summ.termination_type, summ.message = "CONVERGENCE", "Function tolerance reached."
                break
        rho = (x_cost - cand_cost) / model_cost_change if cand_ok else -np.finfo(np.float64).max
        it["relative_decrease"] = rho
        if rho > opt.min_relative_decrease:
            q, t = qc, tc
            x_norm = float(np.sqrt(np.sum(q[1:] ** 2) + np.sum(t[1:] ** 2)))
            x_cost, r, J, g = evaluate(q, t)
            gnorm, gmax = grad_norms(q, t, g)
            it.update(cost=x_cost, gradient_norm=gnorm, gradient_max_norm=gmax, step_is_successful=True)
            radius = min(opt.max_trust_region_radius, radius / max(1.0 / 3.0, 1.0 - (2.0 * rho - 1.0) ** 3))
            decrease_factor = 2.0
            reuse_diagonal = False
        else:
            it["cost"] = cand_cost if cand_ok else x_cost
            radius /= decrease_factor
            decrease_factor *= 2.0
            reuse_diagonal = True
    summ.final_cost = x_cost
    return q, t, summ
