"""Small dense Ceres-faithful trust-region LM.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED (see oracle/__init__.py): restates the control flow of Ceres 2.0/2.1
`TrustRegionMinimizer` + `LevenbergMarquardtStrategy` for problems
small enough to hold J densely — the per-landmark triangulation problems of
`st20-g2o/src/src/sim_data.cpp:298-311` (default `ceres::Solver::Options`), which the
reference hands to `ceres::Solve` one landmark at a time.  Same rules, same order of tests as
`ba_oracle.solve`; the linear solve is a dense normal-equation solve (Ceres' DENSE_QR gives the
same solution to rounding).
"""
import numpy as np

from .ba_oracle import LMOptions, LMSummary


def solve(x0, residual_jacobian, options=None, plus=None):
    """Minimise 1/2 |r(x)|^2.  `residual_jacobian(x) -> (r [m], J [m, n])` with J in the tangent
    space; `plus(x, delta)` defaults to x + delta (Euclidean block)."""
    opt = options or LMOptions()
    plus = plus or (lambda x, d: x + d)
    x = np.array(x0, dtype=np.float64)
    summ = LMSummary()

    def evaluate(x):
        r, J = residual_jacobian(x)
        return 0.5 * float(r @ r), r, J, J.T @ r

    x_cost, r, J, g = evaluate(x)
    scale = 1.0 / (1.0 + np.sqrt(np.sum(J * J, axis=0))) if opt.jacobi_scaling else np.ones(len(x))
    gmax = float(np.max(np.abs(x - plus(x, -g)))) if len(g) else 0.0
    gnorm = float(np.linalg.norm(x - plus(x, -g)))
    x_norm = float(np.linalg.norm(x))
    summ.initial_cost = x_cost
    radius = opt.initial_trust_region_radius
    decrease_factor = 2.0
    reuse_diagonal = False
    diag = None
    num_invalid = 0
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
        Js = J * scale[None, :]
        Hs = Js.T @ Js
        gs = Js.T @ r
        if not reuse_diagonal:
            diag = np.clip(np.diag(Hs), opt.min_lm_diagonal, opt.max_lm_diagonal)
        d2 = diag / radius
        valid = True
        try:
            L = np.linalg.cholesky(Hs + np.diag(d2))
            ys = np.linalg.solve(L.T, np.linalg.solve(L, gs))
            valid = bool(np.all(np.isfinite(ys)))
        except np.linalg.LinAlgError:
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
        delta = step * scale
        xc = plus(x, delta)
        cand_cost = evaluate(xc)[0]
        cand_ok = np.isfinite(cand_cost)
        step_norm = float(np.linalg.norm(xc - x))
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
            x = xc
            x_norm = float(np.linalg.norm(x))
            x_cost, r, J, g = evaluate(x)
            gmax = float(np.max(np.abs(x - plus(x, -g))))
            gnorm = float(np.linalg.norm(x - plus(x, -g)))
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
    return x, summ
