"""NumPy fp64 oracle for the bundle-adjustment hot path.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED (see oracle/__init__.py): Ceres Solver is a third-party
dependency absent from /root/reference; this file restates its published
algorithm and is pinned only to the KATs listed there.

What is restated, and from where:

* residual  — `ProjectFactor::operator()`  st20-g2o/src/include/test_ceres.h:63-80
              (twin: `PnPDynamicAutoDiffFunctor` st17-ceres/src/include/solver.hpp:108-124)
* manifold  — `LieLocalParameterization<SO3d>` test_ceres.h:14-45 (right perturbation)
* Jacobian  — the exact tangent-space derivative that Ceres autodiff composed with the
              plus-Jacobian produces; NOT the inexact hand formula of
              solver.hpp:195 (which drops hat(R^-1 t)); `pnp_reference_jacobian` keeps
              that formula only for the t = 0 cross-check.
* problem   — `SolveWithCeresDynamicAutoDiff` test_ceres.h:98-152: landmark-major
              residual order, first/last camera constant, SPARSE_SCHUR, 1 thread.
* solver    — Ceres 2.0/2.1 `TrustRegionMinimizer` + `LevenbergMarquardtStrategy` +
              `SchurEliminator` with default `Solver::Options`.
* integer   — `DataManager::Jacobian()/Hessian()` sim_data.h:108-159 (occupancy pattern
              of J and J^T J) restated as CSR / degree / co-visibility structures.
"""
from dataclasses import dataclass, field

import numpy as np
import scipy.linalg
import scipy.sparse
import scipy.sparse.linalg

from . import lie


# --------------------------------------------------------------------------- options
@dataclass
class LMOptions:
    """ceres::Solver::Options defaults (Ceres 2.0/2.1 solver.h) for the fields the path uses."""
    max_num_iterations: int = 50
    initial_trust_region_radius: float = 1e4
    max_trust_region_radius: float = 1e16
    min_trust_region_radius: float = 1e-32
    min_relative_decrease: float = 1e-3
    min_lm_diagonal: float = 1e-6
    max_lm_diagonal: float = 1e32
    function_tolerance: float = 1e-6
    gradient_tolerance: float = 1e-10
    parameter_tolerance: float = 1e-8
    jacobi_scaling: bool = True
    max_num_consecutive_invalid_steps: int = 5


@dataclass
class LMSummary:
    iterations: list = field(default_factory=list)
    termination_type: str = "NO_CONVERGENCE"
    message: str = ""
    initial_cost: float = 0.0
    final_cost: float = 0.0
    num_successful_steps: int = 0
    num_unsuccessful_steps: int = 0

    def brief_report(self):
        # format of ceres::Solver::Summary::BriefReport() as seen in st17-ceres/img/release.png
        return ("Ceres Solver Report: Iterations: %d, Initial cost: %e, Final cost: %e, Termination: %s"
                % (len(self.iterations), self.initial_cost, self.final_cost, self.termination_type))


# --------------------------------------------------------------------------- residual / Jacobian
def camera_frame_points(cam_q, cam_t, lm, obs_cam, obs_lm):
    """p_c = (R,t)^-1 * P = R^T (P - t)   — test_ceres.h:71-72."""
    R = lie.quat_to_rot(cam_q)[obs_cam]            # camera -> world
    d = lm[obs_lm] - cam_t[obs_cam]
    return np.einsum("nji,nj->ni", R, d), R


def residuals(cam_q, cam_t, lm, obs_cam, obs_lm, obs_uv):
    """r = (x/z, y/z) - uv   — test_ceres.h:73-77."""
    pc, _ = camera_frame_points(cam_q, cam_t, lm, obs_cam, obs_lm)
    return np.stack([pc[:, 0] / pc[:, 2], pc[:, 1] / pc[:, 2]], axis=-1) - obs_uv


def cost_of(cam_q, cam_t, lm, obs_cam, obs_lm, obs_uv):
    r = residuals(cam_q, cam_t, lm, obs_cam, obs_lm, obs_uv)
    return 0.5 * float(np.sum(r * r))


def residual_jacobian(cam_q, cam_t, lm, obs_cam, obs_lm, obs_uv):
    """Residual and exact tangent-space Jacobian blocks per observation.

    Tangent order per camera is [theta(3), t(3)] (test_g2o.h:36-39, solver.hpp:439-443).
        J_theta = Pi' * hat(p_c),   J_t = -Pi' R^T,   J_P = +Pi' R^T
    Returns r (n,2), Jc (n,2,6), Jl (n,2,3).
    """
    pc, R = camera_frame_points(cam_q, cam_t, lm, obs_cam, obs_lm)
    x, y, z = pc[:, 0], pc[:, 1], pc[:, 2]
    iz = 1.0 / z
    u, v = x * iz, y * iz
    r = np.stack([u, v], axis=-1) - obs_uv
    n = len(obs_cam)
    Pi = np.zeros((n, 2, 3))
    Pi[:, 0, 0] = iz
    Pi[:, 0, 2] = -u * iz
    Pi[:, 1, 1] = iz
    Pi[:, 1, 2] = -v * iz
    Jl = np.einsum("nij,nkj->nik", Pi, R)          # Pi' R^T
    Jth = np.einsum("nij,njk->nik", Pi, lie.hat(pc))
    Jc = np.concatenate([Jth, -Jl], axis=2)
    return r, Jc, Jl


def pnp_reference_jacobian(cam_q, cam_t, point):
    """The reference's hand-written PnP Jacobian, solver.hpp:182-198 (single observation):
    e_R = Pi' (-R^-1 hat(P_w)) (-R),  e_t = Pi' (-R^-1).  e_R is exact only when t = 0."""
    R = lie.quat_to_rot(cam_q)
    pc = R.T @ (point - cam_t)
    iz = 1.0 / pc[2]
    Pi = np.array([[iz, 0, -pc[0] * iz * iz], [0, iz, -pc[1] * iz * iz]])
    e_R = Pi @ (-R.T @ lie.hat(point)) @ (-R)
    e_t = Pi @ (-R.T)
    return e_R, e_t


# --------------------------------------------------------------------------- integer structure
def index_structures(obs_cam, obs_lm, n_cam, n_lm):
    """Integer preprocessing (bit-exact contract).  Restates the occupancy matrices of
    `DataManager::Jacobian()/Hessian()` (sim_data.h:108-159) in sparse form:

    lm_deg / cam_deg  — diagonal of hMat = jMat^T jMat (observations per landmark / camera)
    lm_ptr            — CSR offsets of the landmark-major observation list (test_ceres.h:109-110)
    cam_ptr, cam_perm — CSR of the same observations grouped by camera; cam_perm is the
                        STABLE counting sort (ties keep landmark-major order)
    covis             — strictly-lower camera-pair blocks (i>j) with hMat(i,j) restricted to
                        camera x camera after eliminating landmarks, i.e. pairs of cameras
                        sharing >= 1 landmark (the block pattern of the Schur complement),
                        as sorted unique keys i*n_cam + j
    """
    obs_cam = np.asarray(obs_cam, dtype=np.int64)
    obs_lm = np.asarray(obs_lm, dtype=np.int64)
    assert np.all(np.diff(obs_lm) >= 0), "observations must be landmark-major"
    lm_deg = np.bincount(obs_lm, minlength=n_lm).astype(np.int32)
    cam_deg = np.bincount(obs_cam, minlength=n_cam).astype(np.int32)
    lm_ptr = np.concatenate([[0], np.cumsum(lm_deg)]).astype(np.int32)
    cam_ptr = np.concatenate([[0], np.cumsum(cam_deg)]).astype(np.int32)
    cam_perm = np.argsort(obs_cam, kind="stable").astype(np.int32)
    keys = []
    for l in range(n_lm):
        c = obs_cam[lm_ptr[l]:lm_ptr[l + 1]]
        if len(c) > 1:
            a, b = np.meshgrid(c, c, indexing="ij")
            m = a > b
            keys.append(a[m] * n_cam + b[m])
    covis = np.unique(np.concatenate(keys)) if keys else np.zeros(0, dtype=np.int64)
    return dict(lm_deg=lm_deg, cam_deg=cam_deg, lm_ptr=lm_ptr, cam_ptr=cam_ptr,
                cam_perm=cam_perm, covis=covis.astype(np.int64))


def occupancy_hessian_dense(obs_cam, obs_lm, n_cam, n_lm):
    """Literal `DataManager::Jacobian()` / `Hessian()` (sim_data.h:138-159, 108-111) with dense
    int matrices — only for tiny problems, to pin `index_structures` to the reference code."""
    n = len(obs_cam)
    j = np.zeros((n, n_cam + n_lm), dtype=np.int64)
    j[np.arange(n), n_cam + np.asarray(obs_lm)] = 1
    j[np.arange(n), np.asarray(obs_cam)] = 1
    return j, j.T @ j


# --------------------------------------------------------------------------- normal equations
def normal_blocks(r, Jc, Jl, obs_cam, obs_lm, n_cam, n_lm, cam_const):
    """Block accumulation of J^T J and J^T r (what Ceres' SchurEliminator consumes).
    Observations of a constant camera keep only their landmark block (Ceres removes
    constant parameter blocks from the reduced program)."""
    free = ~np.asarray(cam_const, dtype=bool)[obs_cam]
    Hcc = np.zeros((n_cam, 6, 6))
    gc = np.zeros((n_cam, 6))
    np.add.at(Hcc, obs_cam[free], np.einsum("nki,nkj->nij", Jc[free], Jc[free]))
    np.add.at(gc, obs_cam[free], np.einsum("nki,nk->ni", Jc[free], r[free]))
    Hll = np.zeros((n_lm, 3, 3))
    gl = np.zeros((n_lm, 3))
    np.add.at(Hll, obs_lm, np.einsum("nki,nkj->nij", Jl, Jl))
    np.add.at(gl, obs_lm, np.einsum("nki,nk->ni", Jl, r))
    W = np.einsum("nki,nkj->nij", Jc, Jl) * free[:, None, None]   # H_cl block per observation
    return Hcc, gc, Hll, gl, W


def _lm_pairs(obs_lm, n_lm, free):
    """All ordered pairs (j,k) of FREE-camera observations sharing a landmark (incl. j==k)."""
    idx = np.nonzero(free)[0]
    lm_of = obs_lm[idx]
    deg = np.bincount(lm_of, minlength=n_lm)
    start = np.concatenate([[0], np.cumsum(deg)])[:-1]
    J, K = [], []
    for d in np.unique(deg):
        if d == 0:
            continue
        ls = np.nonzero(deg == d)[0]
        base = start[ls][:, None] + np.arange(d)[None, :]           # (L, d) positions in idx
        J.append(np.repeat(base, d, axis=1).ravel())
        K.append(np.tile(base, (1, d)).ravel())
    if not J:
        return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
    return idx[np.concatenate(J)], idx[np.concatenate(K)]


class SchurSystem:
    """Scaled, damped Schur-complement solve of (J^T J + D^2) y = J^T r for one radius."""

    def __init__(self, Hcc, gc, Hll, gl, W, Jc, Jl, r, obs_cam, obs_lm, cam_const, sc, sl):
        self.obs_cam, self.obs_lm = obs_cam, obs_lm
        self.cam_const = np.asarray(cam_const, dtype=bool)
        self.free_idx = np.nonzero(~self.cam_const)[0]
        self.free_of = -np.ones(len(cam_const), dtype=np.int64)
        self.free_of[self.free_idx] = np.arange(len(self.free_idx))
        self.sc, self.sl = sc, sl
        # column-scaled blocks (jacobi_scaling): J_s = J diag(s)
        self.Hcc = Hcc * sc[:, :, None] * sc[:, None, :]
        self.gc = gc * sc
        self.Hll = Hll * sl[:, :, None] * sl[:, None, :]
        self.gl = gl * sl
        self.W = W * sc[obs_cam][:, :, None] * sl[obs_lm][:, None, :]
        self.Jc = Jc * sc[obs_cam][:, None, :]
        self.Jl = Jl * sl[obs_lm][:, None, :]
        self.r = r
        self.free_obs = ~self.cam_const[obs_cam]
        self.pj, self.pk = _lm_pairs(obs_lm, len(Hll), self.free_obs)

    def diagonal(self, opt):
        """LevenbergMarquardtStrategy: clamp(diag(J_s^T J_s), min, max) (before / radius)."""
        dc = np.clip(np.einsum("nii->ni", self.Hcc), opt.min_lm_diagonal, opt.max_lm_diagonal)
        dl = np.clip(np.einsum("nii->ni", self.Hll), opt.min_lm_diagonal, opt.max_lm_diagonal)
        return dc, dl

    def reduced_system(self, dc2, dl2):
        """S and reduced rhs over the free cameras (dense)."""
        nf = len(self.free_idx)
        n_lm = len(self.Hll)
        Hd = self.Hll + np.einsum("ni,ij->nij", dl2, np.eye(3))
        M = np.linalg.inv(Hd)                                       # (E^T E + D_e^2)^-1
        Y = np.einsum("nij,njk->nik", self.W, M[self.obs_lm])       # (F^T E)(E^T E)^-1 per obs
        rhs = self.gc.copy()
        np.subtract.at(rhs, self.obs_cam[self.free_obs],
                       np.einsum("nij,nj->ni", Y[self.free_obs], self.gl[self.obs_lm[self.free_obs]]))
        Sb = np.zeros((nf * nf, 36))
        diag = self.Hcc[self.free_idx] + np.einsum("ni,ij->nij", dc2[self.free_idx], np.eye(6))
        Sb[np.arange(nf) * nf + np.arange(nf)] = diag.reshape(nf, 36)
        fj = self.free_of[self.obs_cam[self.pj]]
        fk = self.free_of[self.obs_cam[self.pk]]
        chunk = 1 << 20
        for s in range(0, len(fj), chunk):
            e = slice(s, s + chunk)
            blk = np.einsum("nij,nkj->nik", Y[self.pj[e]], self.W[self.pk[e]])
            np.subtract.at(Sb, fj[e] * nf + fk[e], blk.reshape(-1, 36))
        S = Sb.reshape(nf, nf, 6, 6).transpose(0, 2, 1, 3).reshape(6 * nf, 6 * nf)
        return S, rhs[self.free_idx].reshape(-1), M

    def solve(self, dc2, dl2):
        """Returns y_c (n_cam,6; zero rows for constant cameras) and y_l (n_lm,3)."""
        S, rhs, M = self.reduced_system(dc2, dl2)
        S = 0.5 * (S + S.T)
        cf = scipy.linalg.cho_factor(S, lower=True)
        yc_free = scipy.linalg.cho_solve(cf, rhs).reshape(-1, 6)
        yc = np.zeros_like(self.gc)
        yc[self.free_idx] = yc_free
        t = np.zeros_like(self.gl)
        np.add.at(t, self.obs_lm, np.einsum("nij,ni->nj", self.W, yc[self.obs_cam]))
        yl = np.einsum("nij,nj->ni", M, self.gl - t)
        return yc, yl

    def model_cost_change(self, step_c, step_l):
        """Ceres: model_residuals = J_s step;  -model_residuals . (r + model_residuals / 2)."""
        m = np.einsum("nij,nj->ni", self.Jl, step_l[self.obs_lm])
        m += self.free_obs[:, None] * np.einsum("nij,nj->ni", self.Jc, step_c[self.obs_cam])
        return -float(np.sum(m * (self.r + 0.5 * m)))


def full_normal_solve(Jc, Jl, r, obs_cam, obs_lm, cam_const, sc, sl, dc2, dl2):
    """Independent check of the Schur path: assemble the explicit sparse J_s over all free
    tangent dof, solve (J_s^T J_s + D^2) y = J_s^T r with a sparse direct solver."""
    cam_const = np.asarray(cam_const, dtype=bool)
    n_cam, n_lm, n = len(cam_const), len(sl), len(obs_cam)
    free_idx = np.nonzero(~cam_const)[0]
    free_of = -np.ones(n_cam, dtype=np.int64)
    free_of[free_idx] = np.arange(len(free_idx))
    nfc = 6 * len(free_idx)
    rows, cols, vals = [], [], []
    for k in range(2):
        for j in range(6):
            m = ~cam_const[obs_cam]
            rows.append(2 * np.nonzero(m)[0] + k)
            cols.append(6 * free_of[obs_cam[m]] + j)
            vals.append((Jc[:, k, j] * sc[obs_cam, j])[m])
        for j in range(3):
            rows.append(2 * np.arange(n) + k)
            cols.append(nfc + 3 * obs_lm + j)
            vals.append(Jl[:, k, j] * sl[obs_lm, j])
    J = scipy.sparse.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                                shape=(2 * n, nfc + 3 * n_lm))
    D2 = np.concatenate([dc2[free_idx].reshape(-1), dl2.reshape(-1)])
    A = (J.T @ J + scipy.sparse.diags(D2)).tocsc()
    y = scipy.sparse.linalg.spsolve(A, J.T @ r.reshape(-1))
    yc = np.zeros((n_cam, 6))
    yc[free_idx] = y[:nfc].reshape(-1, 6)
    return yc, y[nfc:].reshape(-1, 3), J


# --------------------------------------------------------------------------- state helpers
def apply_delta(cam_q, cam_t, lm, cam_const, dc, dl):
    """evaluator->Plus: q <- q*exp(d_theta) (test_ceres.h:22-29), t <- t + d_t, P <- P + d_P."""
    free = ~np.asarray(cam_const, dtype=bool)
    q2, t2 = cam_q.copy(), cam_t.copy()
    q2[free] = lie.so3_plus(cam_q[free], dc[free, :3])
    t2[free] = cam_t[free] + dc[free, 3:]
    return q2, t2, lm + dl


def ambient_norm(cam_q, cam_t, lm, cam_const, lm_free=None):
    free = ~np.asarray(cam_const, dtype=bool)
    lmf = lm if lm_free is None else lm[lm_free]
    return float(np.sqrt(np.sum(cam_q[free] ** 2) + np.sum(cam_t[free] ** 2) + np.sum(lmf ** 2)))


def ambient_diff_norms(a, b, cam_const):
    """(2-norm, max-norm) of x_a - x_b over the non-constant blocks, in AMBIENT coordinates."""
    free = ~np.asarray(cam_const, dtype=bool)
    parts = [(a[0] - b[0])[free].ravel(), (a[1] - b[1])[free].ravel(), (a[2] - b[2]).ravel()]
    d = np.concatenate(parts)
    return float(np.linalg.norm(d)), float(np.max(np.abs(d))) if len(d) else 0.0


# --------------------------------------------------------------------------- the LM loop
def solve(cam_q, cam_t, lm, obs_cam, obs_lm, obs_uv, cam_const, options=None, callback=None, backend="numpy",
          lm_const=None):
    """Ceres-faithful trust-region Levenberg-Marquardt on the BA problem of test_ceres.h:98-152.

    Returns (cam_q, cam_t, lm, LMSummary).  Control flow follows Ceres 2.0/2.1
    trust_region_minimizer.cc: IterationZero; loop { ComputeTrustRegionStep; invalid-step
    handling; candidate evaluation; parameter-tolerance and function-tolerance tests BEFORE
    acceptance; rho test; StepAccepted / StepRejected radius rules }.
    """
    opt = options or LMOptions()
    cam_q = np.array(cam_q, dtype=np.float64)
    cam_t = np.array(cam_t, dtype=np.float64)
    lm = np.array(lm, dtype=np.float64)
    obs_cam = np.asarray(obs_cam, dtype=np.int64)
    obs_lm = np.asarray(obs_lm, dtype=np.int64)
    cam_const = np.asarray(cam_const, dtype=bool)
    n_cam, n_lm = len(cam_q), len(lm)
    summ = LMSummary()
    # constant landmarks (the PnP problem of solver.hpp:247-385 has ONLY those): their Jacobian
    # block is dropped from the program, i.e. zero, so they get a zero step and do not touch S
    lm_free = np.ones(n_lm, dtype=bool) if lm_const is None else ~np.asarray(lm_const, dtype=bool)

    state = {}
    if backend == "c":
        from . import ba_fast
        obs_uv = np.ascontiguousarray(obs_uv, dtype=np.float64)
        structure = ba_fast.Structure(obs_cam, obs_lm, n_cam, n_lm, cam_const)
        state["times"] = dict(linearize=0.0, schur=0.0, dense=0.0, backsub=0.0, cost=0.0)

    def evaluate_gradient_and_jacobian():
        if backend == "c":
            import time as _t
            t0 = _t.perf_counter()
            sysm = ba_fast.CSystem(structure, cam_q, cam_t, lm, obs_uv, state.get("sc"), state.get("sl"),
                                   opt.jacobi_scaling)
            state["times"]["linearize"] += _t.perf_counter() - t0
            if not lm_free.all():
                raise NotImplementedError("constant landmarks: use the NumPy backend")
            state.setdefault("sc", sysm.sc)
            state.setdefault("sl", sysm.sl)
            gcf = sysm.gc * (~cam_const)[:, None]
            proj = apply_delta(cam_q, cam_t, lm, cam_const, -gcf, -sysm.gl)
            gnorm, gmax = ambient_diff_norms((cam_q, cam_t, lm), proj, cam_const)
            return sysm.cost, sysm, gnorm, gmax
        r, Jc, Jl = residual_jacobian(cam_q, cam_t, lm, obs_cam, obs_lm, obs_uv)
        Jl = Jl * lm_free[obs_lm][:, None, None]
        cost = 0.5 * float(np.sum(r * r))
        Hcc, gc, Hll, gl, W = normal_blocks(r, Jc, Jl, obs_cam, obs_lm, n_cam, n_lm, cam_const)
        if "sc" not in state:
            if opt.jacobi_scaling:   # 1 / (1 + sqrt(squared column norm)), computed once at x0
                state["sc"] = 1.0 / (1.0 + np.sqrt(np.einsum("nii->ni", Hcc)))
                state["sl"] = 1.0 / (1.0 + np.sqrt(np.einsum("nii->ni", Hll)))
            else:
                state["sc"] = np.ones((n_cam, 6))
                state["sl"] = np.ones((n_lm, 3))
        sysm = SchurSystem(Hcc, gc, Hll, gl, W, Jc, Jl, r, obs_cam, obs_lm, cam_const,
                           state["sc"], state["sl"])
        # gradient norms use  x - Plus(x, -g)  in ambient space (bounds-aware form)
        gcf = gc * (~cam_const)[:, None]
        proj = apply_delta(cam_q, cam_t, lm, cam_const, -gcf, -gl)
        gnorm, gmax = ambient_diff_norms((cam_q, cam_t, lm), proj, cam_const)
        return cost, sysm, gnorm, gmax

    # ---- IterationZero
    x_cost, sysm, gnorm, gmax = evaluate_gradient_and_jacobian()
    x_norm = ambient_norm(cam_q, cam_t, lm, cam_const, lm_free)
    summ.initial_cost = x_cost
    radius = opt.initial_trust_region_radius
    decrease_factor = 2.0
    reuse_diagonal = False
    diag_c = diag_l = None
    num_invalid = 0
    it = dict(iteration=0, cost=x_cost, cost_change=0.0, gradient_max_norm=gmax, gradient_norm=gnorm,
              step_norm=0.0, relative_decrease=0.0, trust_region_radius=radius,
              step_is_valid=True, step_is_successful=True)

    while True:
        # ---- FinalizeIterationAndCheckIfMinimizerCanContinue
        if it["step_is_successful"]:
            summ.num_successful_steps += 1
        else:
            summ.num_unsuccessful_steps += 1
        it["trust_region_radius"] = radius
        summ.iterations.append(it)
        if callback is not None and callback(it, (cam_q, cam_t, lm)) is False:
            summ.termination_type, summ.message = "USER_FAILURE", "User callback returned abort."
            break
        if it["iteration"] >= opt.max_num_iterations:
            summ.termination_type = "NO_CONVERGENCE"
            summ.message = "Maximum number of iterations reached."
            break
        if it["step_is_successful"] and it["gradient_max_norm"] <= opt.gradient_tolerance:
            summ.termination_type = "CONVERGENCE"
            summ.message = "Gradient tolerance reached."
            break
        if radius < opt.min_trust_region_radius:
            summ.termination_type = "CONVERGENCE"
            summ.message = "Minimum trust region radius reached."
            break

        prev = it
        it = dict(iteration=prev["iteration"] + 1, cost=x_cost, cost_change=0.0,
                  gradient_max_norm=prev["gradient_max_norm"], gradient_norm=prev["gradient_norm"],
                  step_norm=0.0, relative_decrease=0.0, trust_region_radius=radius,
                  step_is_valid=False, step_is_successful=False)

        # ---- ComputeTrustRegionStep (LevenbergMarquardtStrategy::ComputeStep)
        if not reuse_diagonal:
            diag_c, diag_l = sysm.diagonal(opt)
        dc2, dl2 = diag_c / radius, diag_l / radius
        valid = True
        try:
            yc, yl = sysm.solve(dc2, dl2)
            if not (np.all(np.isfinite(yc)) and np.all(np.isfinite(yl))):
                valid = False
        except np.linalg.LinAlgError:
            valid = False
        reuse_diagonal = True
        if valid:
            step_c, step_l = -yc, -yl
            model_cost_change = sysm.model_cost_change(step_c, step_l)
            if not (model_cost_change > 0.0):
                valid = False
        if not valid:
            # ---- HandleInvalidStep: StepIsInvalid() == StepRejected(0), reuse_diagonal = false
            num_invalid += 1
            if num_invalid >= opt.max_num_consecutive_invalid_steps:
                summ.termination_type = "FAILURE"
                summ.message = "Number of consecutive invalid steps more than Solver::Options::max_num_consecutive_invalid_steps"
                break
            radius = radius / decrease_factor
            decrease_factor *= 2.0
            reuse_diagonal = False
            continue
        num_invalid = 0
        it["step_is_valid"] = True

        # ---- undo the column scaling, candidate point, candidate cost
        dc, dl = step_c * sysm.sc, step_l * sysm.sl
        cq, ct, cl = apply_delta(cam_q, cam_t, lm, cam_const, dc, dl)
        if backend == "c":
            import time as _t
            t0 = _t.perf_counter()
            cand_cost = ba_fast.cost(structure, cq, ct, cl, obs_uv)
            state["times"]["cost"] += _t.perf_counter() - t0
            for k in ("schur", "dense", "backsub"):
                state["times"][k] += sysm.times.get(k, 0.0)
        else:
            cand_cost = cost_of(cq, ct, cl, obs_cam, obs_lm, obs_uv)
        cand_ok = np.isfinite(cand_cost)

        # ---- ParameterToleranceReached (ambient step norm)
        step_norm, _ = ambient_diff_norms((cam_q, cam_t, lm), (cq, ct, cl), cam_const)
        it["step_norm"] = step_norm
        if step_norm <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance):
            summ.termination_type = "CONVERGENCE"
            summ.message = "Parameter tolerance reached."
            break
        # ---- FunctionToleranceReached
        if cand_ok:
            it["cost_change"] = x_cost - cand_cost
            if abs(it["cost_change"]) <= opt.function_tolerance * x_cost:
                summ.termination_type = "CONVERGENCE"
                summ.message = "Function tolerance reached."
                break
        # ---- IsStepSuccessful
        rho = (x_cost - cand_cost) / model_cost_change if cand_ok else -np.finfo(np.float64).max
        it["relative_decrease"] = rho
        if rho > opt.min_relative_decrease:
            cam_q, cam_t, lm = cq, ct, cl
            x_norm = ambient_norm(cam_q, cam_t, lm, cam_const, lm_free)
            x_cost, sysm, gnorm, gmax = evaluate_gradient_and_jacobian()
            it.update(cost=x_cost, gradient_norm=gnorm, gradient_max_norm=gmax, step_is_successful=True)
            radius = min(opt.max_trust_region_radius,
                         radius / max(1.0 / 3.0, 1.0 - (2.0 * rho - 1.0) ** 3))
            decrease_factor = 2.0
            reuse_diagonal = False
        else:
            it["cost"] = cand_cost if cand_ok else x_cost   # Ceres logs the candidate's cost; x unchanged
            radius = radius / decrease_factor
            decrease_factor *= 2.0
            reuse_diagonal = True

    summ.final_cost = min(i["cost"] for i in summ.iterations)
    summ.phase_times = state.get("times")
    return cam_q, cam_t, lm, summ
