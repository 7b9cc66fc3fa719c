"""Pins the oracle to everything the reference records for this path:
closed forms from the reference's notes, the PnP poses of st17-ceres/img/release.png, the
reference's own hand Gauss-Newton, and self-consistency of the restated Ceres algebra."""
import numpy as np
import pytest

from oracle import ba_oracle as bo
from oracle import lie
from conftest import scene_args


def _rand_quat(rng, n=None):
    q = rng.normal(size=(4,) if n is None else (n, 4))
    return q / np.linalg.norm(q, axis=-1, keepdims=True)


def test_plus_jacobian_closed_form_matches_finite_differences():
    # test_ceres.h:32-38 / st17-ceres/docs/notes.tex:131-144
    rng = np.random.default_rng(0)
    q = _rand_quat(rng)
    J = lie.so3_plus_jacobian(q)
    eps = 1e-7
    for k in range(3):
        d = np.zeros(3); d[k] = eps
        fd = (lie.quat_mul(q, lie.so3_exp_quat(d)) - lie.quat_mul(q, lie.so3_exp_quat(-d))) / (2 * eps)
        assert np.allclose(fd, J[:, k], atol=1e-9)


def test_exp_log_round_trip_and_taylor_branch():
    rng = np.random.default_rng(1)
    w = rng.normal(size=(50, 3))
    w *= rng.uniform(1e-12, 3.0, size=(50, 1)) / np.linalg.norm(w, axis=-1, keepdims=True)   # |w| < pi
    assert np.allclose(lie.so3_log_quat(lie.so3_exp_quat(w)), w, atol=1e-12)
    tiny = np.array([1e-11, -2e-11, 3e-12])
    q = lie.so3_exp_quat(tiny)
    assert np.allclose(q, [0.5e-11, -1e-11, 1.5e-12, 1.0], atol=1e-22)
    assert abs(np.linalg.norm(lie.so3_exp_quat(w), axis=-1) - 1).max() < 1e-15


def test_exact_jacobian_matches_central_differences(scene_small):
    sc = scene_small
    r, Jc, Jl = bo.residual_jacobian(sc.cam_q, sc.cam_t, sc.lm, sc.obs_cam, sc.obs_lm, sc.obs_uv)
    eps = 1e-6
    sel = np.arange(0, sc.n_obs, 97)
    for k in range(6):
        d = np.zeros((sc.n_cam, 6)); d[:, k] = eps
        free = np.zeros(sc.n_cam, dtype=np.uint8)
        qp, tp, _ = bo.apply_delta(sc.cam_q, sc.cam_t, sc.lm, free, d, np.zeros_like(sc.lm))
        qm, tm, _ = bo.apply_delta(sc.cam_q, sc.cam_t, sc.lm, free, -d, np.zeros_like(sc.lm))
        fd = (bo.residuals(qp, tp, sc.lm, sc.obs_cam, sc.obs_lm, sc.obs_uv) -
              bo.residuals(qm, tm, sc.lm, sc.obs_cam, sc.obs_lm, sc.obs_uv)) / (2 * eps)
        assert np.allclose(fd[sel], Jc[sel, :, k], atol=2e-8)
    for k in range(3):
        d = np.zeros_like(sc.lm); d[:, k] = eps
        fd = (bo.residuals(sc.cam_q, sc.cam_t, sc.lm + d, sc.obs_cam, sc.obs_lm, sc.obs_uv) -
              bo.residuals(sc.cam_q, sc.cam_t, sc.lm - d, sc.obs_cam, sc.obs_lm, sc.obs_uv)) / (2 * eps)
        assert np.allclose(fd[sel], Jl[sel, :, k], atol=2e-8)


def test_reference_hand_jacobian_agrees_when_t_is_zero_only():
    # solver.hpp:182-198: e_R drops hat(R^-1 t); exact iff t = 0, e_t always exact
    rng = np.random.default_rng(2)
    q = _rand_quat(rng); P = np.array([0.3, -0.2, 4.0]) + rng.normal(size=3) * 0.1
    for t, same in ((np.zeros(3), True), (np.array([0.4, -0.3, 0.2]), False)):
        P_w = lie.quat_to_rot(q) @ P + t          # keeps the point in front of the camera
        e_R, e_t = bo.pnp_reference_jacobian(q, t, P_w)
        _, Jc, _ = bo.residual_jacobian(q[None], t[None], P_w[None], np.array([0]), np.array([0]), np.zeros((1, 2)))
        assert np.allclose(e_t, Jc[0, :, 3:], atol=1e-13)
        assert np.allclose(e_R, Jc[0, :, :3], atol=1e-12) == same


def test_pnp_poses_match_release_png(stba):
    # st17-ceres/img/release.png (readme.md:246): real pose quaternion and the init pose
    q_real, t_real, q_init, t_init = stba.synth.pnp_poses()
    want_real = np.array([0.40958, 0.70941, -0.49673, -0.28679])
    want_init = np.array([0.45452, 0.54168, -0.54168, -0.45452])
    assert min(abs(q_real - want_real).max(), abs(q_real + want_real).max()) < 6e-6
    assert min(abs(q_init - want_init).max(), abs(q_init + want_init).max()) < 6e-6
    assert np.allclose(t_real, [3, 2, 1]) and np.allclose(t_init, [2.5, 0, 0])


def _pnp_problem(stba):
    s = stba.synth.pnp_scene()
    n = len(s["points"])
    return s, dict(cam_q=s["q_init"][None].copy(), cam_t=s["t_init"][None].copy(), lm=s["points"],
                   obs_cam=np.zeros(n, np.int32), obs_lm=np.arange(n, dtype=np.int32), obs_uv=s["uv"],
                   cam_const=np.zeros(1, np.uint8), lm_const=np.ones(n, np.uint8))


def test_oracle_pnp_recovers_the_published_pose(stba):
    # release.png: all Ceres variants recover the truth to 5 decimals from the init pose, final cost
    # ~1e-21 (zero-residual problem), CONVERGENCE, 6 iterations for the exact-Jacobian variants
    s, p = _pnp_problem(stba)
    assert len(s["points"]) >= 8
    q, t, _, summ = bo.solve(p["cam_q"], p["cam_t"], p["lm"], p["obs_cam"], p["obs_lm"], p["obs_uv"], p["cam_const"],
                             lm_const=p["lm_const"])
    assert summ.termination_type == "CONVERGENCE"
    assert min(abs(q[0] - s["q_real"]).max(), abs(q[0] + s["q_real"]).max()) < 1e-5
    assert abs(t[0] - s["t_real"]).max() < 1e-5
    assert summ.final_cost < 1e-15
    assert 4 <= len(summ.iterations) <= 10


def test_hand_gauss_newton_of_the_reference_also_recovers_it(stba):
    # SelfGaussNewton, solver.hpp:387-462, restated literally (inexact e_R included): H = sum J^T J,
    # g = -sum J^T r, ldlt solve, R <- R exp(d_theta), t <- t + d_t, stop at |d| < 1e-8, <= 10 iterations
    s = stba.synth.pnp_scene()
    q, t = s["q_init"].copy(), s["t_init"].copy()
    for it in range(10):
        H = np.zeros((6, 6)); g = np.zeros(6)
        R = lie.quat_to_rot(q)
        for P, uv in zip(s["points"], s["uv"]):
            pc = R.T @ (P - t)
            r = pc[:2] / pc[2] - uv
            e_R, e_t = bo.pnp_reference_jacobian(q, t, P)
            J = np.concatenate([e_R, e_t], axis=1)
            H += J.T @ J; g -= J.T @ r
        d = np.linalg.solve(H, g)
        q = lie.so3_plus(q, d[:3]); t = t + d[3:]
        if np.linalg.norm(d[:3]) + np.linalg.norm(d[3:]) < 1e-8:
            break
    assert min(abs(q - s["q_real"]).max(), abs(q + s["q_real"]).max()) < 1e-5
    assert abs(t - s["t_real"]).max() < 1e-5


def test_schur_solution_equals_full_normal_equations(scene_small):
    sc = scene_small
    r, Jc, Jl = bo.residual_jacobian(sc.cam_q, sc.cam_t, sc.lm, sc.obs_cam, sc.obs_lm, sc.obs_uv)
    Hcc, gc, Hll, gl, W = bo.normal_blocks(r, Jc, Jl, sc.obs_cam, sc.obs_lm, sc.n_cam, sc.n_lm, sc.cam_const)
    s_c = 1 / (1 + np.sqrt(np.einsum("nii->ni", Hcc))); s_l = 1 / (1 + np.sqrt(np.einsum("nii->ni", Hll)))
    sysm = bo.SchurSystem(Hcc, gc, Hll, gl, W, Jc, Jl, r, sc.obs_cam, sc.obs_lm, sc.cam_const, s_c, s_l)
    opt = bo.LMOptions()
    dc, dl = sysm.diagonal(opt)
    yc, yl = sysm.solve(dc / 1e4, dl / 1e4)
    yc2, yl2, J = bo.full_normal_solve(Jc, Jl, r, sc.obs_cam, sc.obs_lm, sc.cam_const, s_c, s_l, dc / 1e4, dl / 1e4)
    assert np.allclose(yc, yc2, rtol=0, atol=1e-9 * abs(yc2).max())
    assert np.allclose(yl, yl2, rtol=0, atol=1e-9 * abs(yl2).max())
    # closed-form model cost change == Ceres' per-residual form
    mcc = sysm.model_cost_change(-yc, -yl)
    closed = 0.5 * (np.sum(yc * (sysm.gc + dc / 1e4 * yc)) + np.sum(yl * (sysm.gl + dl / 1e4 * yl)))
    assert abs(mcc - closed) < 1e-9 * abs(mcc)


def test_index_structures_match_the_dense_occupancy_matrices(stba):
    # DataManager::Jacobian()/Hessian(), sim_data.h:108-159, on a problem small enough to be dense
    sc = stba.synth.make_scene(6, 40, 80)
    ix = bo.index_structures(sc.obs_cam, sc.obs_lm, sc.n_cam, sc.n_lm)
    j, h = bo.occupancy_hessian_dense(sc.obs_cam, sc.obs_lm, sc.n_cam, sc.n_lm)
    assert np.array_equal(np.diag(h)[:sc.n_cam], ix["cam_deg"])
    assert np.array_equal(np.diag(h)[sc.n_cam:], ix["lm_deg"])
    cov = (j[:, :sc.n_cam].T @ (j[:, sc.n_cam:] @ j[:, sc.n_cam:].T) @ j[:, :sc.n_cam]) > 0
    want = sorted(i * sc.n_cam + k for i in range(sc.n_cam) for k in range(i) if cov[i, k])
    assert ix["covis"].tolist() == want
    assert np.array_equal(sc.obs_cam[ix["cam_perm"]], np.sort(sc.obs_cam))
    for c in range(sc.n_cam):
        seg = ix["cam_perm"][ix["cam_ptr"][c]:ix["cam_ptr"][c + 1]]
        assert np.all(np.diff(seg) > 0)


def test_lm_loop_converges_and_is_deterministic(scene_small):
    q, t, l, s = bo.solve(*scene_args(scene_small))
    q2, t2, l2, s2 = bo.solve(*scene_args(scene_small))
    assert s.termination_type == "CONVERGENCE" and s.final_cost < 1e-2 * s.initial_cost
    assert np.array_equal(q, q2) and np.array_equal(l, l2)
    assert s.brief_report().startswith("Ceres Solver Report: Iterations: ")


def test_c_twin_equals_numpy_oracle(scene_small):
    from oracle import ba_fast
    ba_fast.set_num_threads(2)
    q, t, l, s = bo.solve(*scene_args(scene_small))
    q2, t2, l2, s2 = bo.solve(*scene_args(scene_small), backend="c")
    assert len(s.iterations) == len(s2.iterations) and s.termination_type == s2.termination_type
    assert abs(q - q2).max() < 1e-12 and abs(t - t2).max() < 1e-12 and abs(l - l2).max() < 1e-12
    assert abs(s.final_cost - s2.final_cost) <= 1e-12 * s.final_cost
    for a, b in zip(s.iterations, s2.iterations):
        assert abs(a["cost"] - b["cost"]) <= 1e-10 * max(a["cost"], 1e-300)
        assert abs(a["relative_decrease"] - b["relative_decrease"]) < 1e-8


# ---- pinned against outputs of the reference's OWN code (tests/golden/ref_kat.npz, made by make_ref_kat.py) -------
import os as _os

_REF_KAT = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden", "ref_kat.npz")


@pytest.fixture(scope="module")
def ref_kat():
    z = np.load(_REF_KAT)
    c, o = z["cases"], z["out"]
    return dict(q=c[:, :4], t=c[:, 4:7], P=c[:, 7:10], uv=c[:, 10:12], delta=c[:, 12:15],
                r=o[:, 0:2], Jq=o[:, 2:10].reshape(-1, 2, 4), Jt=o[:, 10:16].reshape(-1, 2, 3), JP=o[:, 16:22].reshape(-1, 2, 3),
                plus=o[:, 22:26], lpJ=o[:, 26:38].reshape(-1, 4, 3), tri_r=o[:, 38:40], tri_J=o[:, 40:46].reshape(-1, 2, 3),
                log=o[:, 46:49], pnp_r=o[:, 49:51], pnp_JR=o[:, 51:57].reshape(-1, 2, 3), pnp_Jt=o[:, 57:63].reshape(-1, 2, 3),
                plus3=o[:, 63:66], gn_pose=z["gn_pose"], gn_iter_num=int(z["gn_iter_num"]))


def test_residual_and_exact_jacobian_equal_jet_autodiff_of_the_reference_functor(ref_kat):
    # ProjectFactor::operator()<Jet> (test_ceres.h:63-80) differentiated by ceres::Jet through the Sophus / Eigen calls
    # of the reference text, chained with LieLocalParameterization::ComputeJacobian (:32-38) = what Ceres hands its
    # linear solver; the oracle's closed form must be that Jacobian
    k = ref_kat
    n = len(k["q"])
    idx = np.arange(n, dtype=np.int32)
    r, Jc, Jl = bo.residual_jacobian(k["q"], k["t"], k["P"], idx, idx, k["uv"])
    assert np.max(np.abs(r - k["r"])) < 1e-13
    J_theta = np.einsum("nij,njk->nik", k["Jq"], k["lpJ"])
    scale = np.maximum(1.0, np.abs(Jc).max(axis=(1, 2)))[:, None, None]
    assert np.max(np.abs(Jc[:, :, :3] - J_theta) / scale) < 1e-12
    assert np.max(np.abs(Jc[:, :, 3:] - k["Jt"]) / scale) < 1e-12
    assert np.max(np.abs(Jl - k["JP"]) / scale) < 1e-12


def test_manifold_equals_the_reference_parameterizations(ref_kat):
    k = ref_kat
    for q, d, want, J in zip(k["q"], k["delta"], k["plus"], k["lpJ"]):
        assert np.max(np.abs(lie.so3_plus(q, d) - want)) < 1e-15       # LieLocalParameterization<SO3d>::Plus
        assert np.max(np.abs(lie.so3_plus_jacobian(q) - J)) < 1e-15    # ::ComputeJacobian
    # so3.log() and LieR3LocalParameterization::Plus (solver.hpp:67-78): x <- Log(Exp(x) Exp(delta)), up to the 2 pi branch
    for q, d, w, wp in zip(k["q"], k["delta"], k["log"], k["plus3"]):
        mine = lie.so3_log_quat(q)
        assert min(np.abs(mine - w).max(), np.abs(lie.quat_mul(lie.so3_exp_quat(mine), lie.so3_exp_quat(-w)) - [0, 0, 0, 1]).max(),
                   np.abs(lie.quat_mul(lie.so3_exp_quat(mine), lie.so3_exp_quat(-w)) + [0, 0, 0, 1]).max()) < 1e-12
        got = lie.so3_exp_quat(lie.so3_log_quat(lie.quat_mul(lie.so3_exp_quat(w), lie.so3_exp_quat(d))))
        ref = lie.so3_exp_quat(wp)
        assert min(np.abs(got - ref).max(), np.abs(got + ref).max()) < 1e-12


def test_triangulation_functor_equals_the_reference(ref_kat):
    from oracle import front_oracle as fo
    k = ref_kat
    R, tcw = fo.world_to_camera(k["q"], k["t"])
    for i in range(len(k["q"])):
        r, J = fo.triangulation_residual_jacobian(R[i:i + 1], tcw[i:i + 1], k["uv"][i:i + 1], k["P"][i])
        assert np.max(np.abs(r - k["tri_r"][i])) < 1e-13
        assert np.max(np.abs(J - k["tri_J"][i])) < 1e-12 * max(1.0, np.abs(J).max())


def test_reference_hand_jacobian_and_gauss_newton_reproduced(ref_kat, stba):
    k = ref_kat
    for i in range(len(k["q"])):        # PnPSizedCostFunction::Evaluate, solver.hpp:168-212
        e_R, e_t = bo.pnp_reference_jacobian(k["q"][i], k["t"][i], k["P"][i])
        s = max(1.0, np.abs(e_R).max())
        assert np.max(np.abs(e_R - k["pnp_JR"][i])) < 1e-11 * s and np.max(np.abs(e_t - k["pnp_Jt"][i])) < 1e-12 * s
        assert np.max(np.abs(k["pnp_r"][i] - k["r"][i])) < 1e-12
    # SelfGaussNewton (solver.hpp:387-462) run from the reference source: same pose, same `iter num`
    s = stba.synth.pnp_scene()
    q, t = s["q_init"].copy(), s["t_init"].copy()
    it = 0
    for it in range(10):
        H = np.zeros((6, 6)); g = np.zeros(6)
        R = lie.quat_to_rot(q)
        for P, uv in zip(s["points"], s["uv"]):
            pc = R.T @ (P - t)
            r = pc[:2] / pc[2] - uv
            e_R, e_t = bo.pnp_reference_jacobian(q, t, P)
            J = np.concatenate([e_R, e_t], axis=1)
            H += J.T @ J; g -= J.T @ r
        d = np.linalg.solve(H, g)
        q = lie.so3_plus(q, d[:3]); t = t + d[3:]
        if np.linalg.norm(d[:3]) + np.linalg.norm(d[3:]) < 1e-8:
            break
    assert it == k["gn_iter_num"]
    assert min(np.abs(q - k["gn_pose"][:4]).max(), np.abs(q + k["gn_pose"][:4]).max()) < 1e-12
    assert np.abs(t - k["gn_pose"][4:]).max() < 1e-12
