"""Front of the path: visibility predicate + ordered compaction, batched
landmark triangulation.  CPU tests pin the oracle; `-m gpu` tests compare the CUDA path with it
through the C ABI (index work bit-exact, floating point to the tolerances written below)."""
import os
import numpy as np
import pytest

from oracle import ba_oracle as bo
from oracle import dense_lm, front_oracle as fo, lie


def reference_scene(stba, n_cam=29, n_lm=600, seed=7):
    """The reference's own scene shape (test_ceres.cpp:8): 29 spiral cameras, 600 points on the cube faces."""
    R, pos = stba.synth.spiral_cameras(n_cam)
    rng = np.random.default_rng(seed)
    pts = stba.synth._face_points(np.arange(n_lm), rng.uniform(-5, 5, size=(n_lm, 2)))
    pts = pts.astype(np.float32).astype(np.float64)          # pcl::PointXYZ storage, sim_data.cpp:36
    return stba.synth._quat_from_rot(R), pos, pts


# ------------------------------------------------------------------ oracle (CPU)
def test_oracle_visibility_matches_the_scene_generator(stba):
    q, t, pts = reference_scene(stba)
    got = fo.visibility(q, t, pts)
    R = lie.quat_to_rot(q)
    vis, u, v = stba.synth.visibility(R, t, pts)             # same predicate written as R^T (P - t)
    # the two groupings of the arithmetic may differ on razor-edge points only
    l, c = np.nonzero(vis)
    a = set(zip(got["obs_lm"].tolist(), got["obs_cam"].tolist())); b = set(zip(l.tolist(), c.tolist()))
    assert len(a ^ b) <= 2 and len(a) > 1000
    assert np.all(np.diff(got["obs_lm"]) >= 0)
    same = got["obs_lm"][1:] == got["obs_lm"][:-1]
    assert np.all(np.diff(got["obs_cam"])[same] > 0)         # camera-ascending inside a landmark (sim_data.cpp:123)
    assert got["lm_deg"].sum() == got["cam_deg"].sum() == len(got["obs_cam"])
    # camera -> [landmark] lists are landmark-ascending and hold the same pairs
    ptr = np.concatenate([[0], np.cumsum(got["cam_deg"])])
    pairs = set()
    for c in range(len(q)):
        seg = got["cam_lm"][ptr[c]:ptr[c + 1]]
        assert np.all(np.diff(seg) > 0)
        pairs |= {(int(x), c) for x in seg}
    assert pairs == a
    # features are float32-rounded (pcl::PointXY, sim_data.cpp:135-136)
    assert np.array_equal(got["obs_uv"], got["obs_uv"].astype(np.float32).astype(np.float64))
    assert np.max(np.abs(got["obs_uv"][:, 0])) < 0.8 and np.max(np.abs(got["obs_uv"][:, 1])) < 0.6


def test_oracle_visibility_edge_cases():
    q = np.array([[0, 0, 0, 1.0]]); t = np.zeros((1, 3))
    pts = np.array([[0, 0, 1.0], [0, 0, -1.0], [0.8, 0, 1.0], [0.79, 0.59, 1.0], [0, 0, 0.0], [1e-3, 0, 0.0]])
    got = fo.visibility(q, t, pts, round_uv_f32=False)
    # behind: out; exactly on the border: out (strict <); z == 0: 0/0 = nan and x/0 = inf both fail the test
    assert got["obs_lm"].tolist() == [0, 3]
    assert fo.visibility(q, t, np.zeros((0, 3)))["obs_cam"].shape == (0,)


def test_dense_lm_oracle_agrees_with_the_ba_oracle_on_a_one_landmark_problem(stba):
    # one free landmark seen by constant cameras is the same Ceres problem up to the sign of the residual
    sc = stba.synth.make_scene(40, 1, 5, sigma_uv=1e-3)
    cc = np.ones(sc.n_cam, dtype=np.uint8)
    _, _, want_lm, want = bo.solve(sc.cam_q, sc.cam_t, sc.lm, sc.obs_cam, sc.obs_lm, sc.obs_uv, cc)
    R, tcw = fo.world_to_camera(sc.cam_q, sc.cam_t)
    got, s = dense_lm.solve(sc.lm[0], lambda P: fo.triangulation_residual_jacobian(R[sc.obs_cam], tcw[sc.obs_cam], sc.obs_uv, P))
    assert len(s.iterations) == len(want.iterations) and s.termination_type == want.termination_type
    assert np.allclose(got, want_lm[0], rtol=0, atol=1e-10)
    assert abs(s.final_cost - want.final_cost) <= 1e-12 * max(want.final_cost, 1e-30) + 1e-24


def test_oracle_triangulation_jacobian_by_finite_differences(stba):
    q, t, pts = reference_scene(stba, 8, 5)
    R, tcw = fo.world_to_camera(q, t)
    uv = np.zeros((8, 2))
    P = np.array([0.3, -0.2, 4.0])
    r, J = fo.triangulation_residual_jacobian(R, tcw, uv, P)
    for k in range(3):
        d = np.zeros(3); d[k] = 1e-6
        num = (fo.triangulation_residual_jacobian(R, tcw, uv, P + d)[0] - fo.triangulation_residual_jacobian(R, tcw, uv, P - d)[0]) / 2e-6
        assert np.allclose(J[:, k], num, atol=1e-7)


def test_oracle_triangulation_recovers_truth_with_exact_cameras(stba):
    q, t, pts = reference_scene(stba, 29, 60)
    vis = fo.visibility(q, t, pts, round_uv_f32=False)
    seen = np.nonzero(vis["lm_deg"] >= 2)[0]
    rng = np.random.default_rng(1)
    lm0 = pts + rng.normal(0, 0.2, pts.shape)
    lm, its, cost, term = fo.triangulate(q, t, lm0, vis["obs_cam"], vis["obs_lm"], vis["obs_uv"])
    assert np.max(np.abs(lm[seen] - pts[seen])) < 1e-6 and set(term) == {"CONVERGENCE"} and cost[seen].max() < 1e-12


# ------------------------------------------------------------------ CUDA path vs oracle (GPU)
def _vis_equal(got, want):
    for k in ("lm_deg", "cam_deg", "obs_cam", "obs_lm", "cam_lm"):
        assert np.array_equal(got[k], want[k]), k
    assert np.array_equal(got["obs_uv"], want["obs_uv"])      # same expression tree, IEEE division: bit for bit


@pytest.mark.gpu
@pytest.mark.parametrize("size", [(29, 600), (50, 5000), (1100, 700), (3, 1)])
def test_visibility_bit_exact(stba, size):
    q, t, pts = reference_scene(stba, *size)
    _vis_equal(stba.front.visibility(q, t, pts), fo.visibility(q, t, pts))
    _vis_equal(stba.front.visibility(q, t, pts, round_uv_f32=False), fo.visibility(q, t, pts, round_uv_f32=False))


@pytest.mark.gpu
def test_visibility_noisy_poses_and_edge_cases(stba):
    sc = stba.synth.make_scene(20, 300, 1200)
    _vis_equal(stba.front.visibility(sc.cam_q, sc.cam_t, sc.lm), fo.visibility(sc.cam_q, sc.cam_t, sc.lm))
    q = np.array([[0, 0, 0, 1.0]]); t = np.zeros((1, 3))
    pts = np.array([[0, 0, 1.0], [0, 0, -1.0], [0.8, 0, 1.0], [0.79, 0.59, 1.0], [0, 0, 0.0], [1e-3, 0, 0.0]])
    _vis_equal(stba.front.visibility(q, t, pts), fo.visibility(q, t, pts))
    empty = stba.front.visibility(q, t, np.zeros((0, 3)))
    assert len(empty["obs_cam"]) == 0 and empty["cam_deg"].tolist() == [0]
    none_seen = stba.front.visibility(q, t, np.array([[0, 0, -2.0]]))
    assert len(none_seen["obs_cam"]) == 0 and none_seen["lm_deg"].tolist() == [0]


@pytest.mark.gpu
def test_visibility_config_C_properties(stba):
    """1000 cameras x 100k candidate points (1e8 predicate tests): size-independent properties + a checksum
    against the oracle on a slice."""
    R, pos = stba.synth.spiral_cameras(1000)
    q = stba.synth._quat_from_rot(R)
    rng = np.random.default_rng(3)
    pts = stba.synth._face_points(np.arange(100000), rng.uniform(-5, 5, size=(100000, 2)))
    got = stba.front.visibility(q, pos, pts)
    n = len(got["obs_cam"])
    assert got["lm_deg"].sum() == n == got["cam_deg"].sum() and n > 10_000_000
    assert np.all(np.diff(got["obs_lm"]) >= 0)
    same = got["obs_lm"][1:] == got["obs_lm"][:-1]
    assert np.all(np.diff(got["obs_cam"])[same] > 0)
    assert np.array_equal(np.bincount(got["obs_cam"], minlength=1000), got["cam_deg"])
    order = np.argsort(got["obs_cam"], kind="stable")
    assert np.array_equal(got["cam_lm"], got["obs_lm"][order])
    want = fo.visibility(q, pos, pts[:3000])
    m = len(want["obs_cam"])
    assert np.array_equal(got["obs_cam"][:m], want["obs_cam"]) and np.array_equal(got["obs_uv"][:m], want["obs_uv"])


def _tri_problem(stba, n_cam, n_lm, seed=5, noise=0.2):
    sc = stba.synth.make_scene(n_cam, n_lm, 4 * n_lm, seed=seed)
    rng = np.random.default_rng(seed + 1)
    lm0 = sc.true_lm + rng.normal(0, noise, sc.true_lm.shape)
    return sc, lm0


@pytest.mark.gpu
def test_triangulation_matches_oracle_landmark_by_landmark(stba):
    sc, lm0 = _tri_problem(stba, 20, 300)
    want_lm, want_it, want_cost, want_term = fo.triangulate(sc.cam_q, sc.cam_t, lm0, sc.obs_cam, sc.obs_lm, sc.obs_uv)
    lm, its, cost, term, ms = stba.front.triangulate(sc.cam_q, sc.cam_t, lm0, sc.obs_cam, sc.obs_lm, sc.obs_uv)
    assert term == want_term
    assert np.array_equal(its, want_it)                                   # same accept/reject sequence per landmark
    assert np.max(np.abs(lm - want_lm)) < 1e-9                            # parameter bar of the north star is 1e-5
    assert np.allclose(cost, want_cost, rtol=1e-9, atol=1e-20)
    assert ms > 0


@pytest.mark.gpu
def test_triangulation_options_edge_cases(stba):
    sc, lm0 = _tri_problem(stba, 20, 40, seed=9, noise=0.5)
    opt = stba.capi.Options(max_num_iterations=2, jacobi_scaling=0)
    want = fo.triangulate(sc.cam_q, sc.cam_t, lm0, sc.obs_cam, sc.obs_lm, sc.obs_uv, bo.LMOptions(max_num_iterations=2, jacobi_scaling=False))
    lm, its, cost, term, _ = stba.front.triangulate(sc.cam_q, sc.cam_t, lm0, sc.obs_cam, sc.obs_lm, sc.obs_uv, opt)
    assert term == want[3] and np.array_equal(its, want[1]) and np.max(np.abs(lm - want[0])) < 1e-9
    assert "NO_CONVERGENCE" in term
    # a landmark without observations keeps its value; unsorted observations are rejected
    lm1, its1, _, term1, _ = stba.front.triangulate(sc.cam_q, sc.cam_t, np.vstack([lm0, [[1.0, 2.0, 3.0]]]), sc.obs_cam, sc.obs_lm, sc.obs_uv)
    assert lm1[-1].tolist() == [1.0, 2.0, 3.0] and its1[-1] == 0 and term1[-1] == "CONVERGENCE"
    with pytest.raises(stba.capi.StbaError):
        stba.front.triangulate(sc.cam_q, sc.cam_t, lm0, sc.obs_cam[::-1], sc.obs_lm[::-1], sc.obs_uv[::-1])


@pytest.mark.gpu
def test_triangulation_config_C(stba):
    """100k landmarks / 1M observations in one launch: the cost never increases, the gradient of every
    converged landmark vanishes, and a random sample agrees with the oracle."""
    import bench
    d = bench.load_scene("C")
    lm, its, cost, term, ms = stba.front.triangulate(d["cam_q"], d["cam_t"], d["lm"], d["obs_cam"], d["obs_lm"], d["obs_uv"])
    R, tcw = fo.world_to_camera(d["cam_q"], d["cam_t"])
    def cost_of(P):
        pc = np.einsum("nji,nj->ni", R[d["obs_cam"]], P[d["obs_lm"]]) + tcw[d["obs_cam"]]
        r = d["obs_uv"] - pc[:, :2] / pc[:, 2:3]
        return 0.5 * np.bincount(d["obs_lm"], weights=(r * r).sum(1), minlength=len(P))
    c0, c1 = cost_of(d["lm"]), cost_of(lm)
    assert np.all(c1 <= c0 * (1 + 1e-12)) and np.allclose(c1, cost, rtol=1e-9, atol=1e-18)
    assert term.count("CONVERGENCE") == len(term)
    ptr = np.concatenate([[0], np.cumsum(np.bincount(d["obs_lm"], minlength=len(lm)))])
    pick = np.random.default_rng(0).choice(len(lm), 150, replace=False)
    for l in pick:
        s = slice(ptr[l], ptr[l + 1])
        x, sm = dense_lm.solve(d["lm"][l], lambda P: fo.triangulation_residual_jacobian(R[d["obs_cam"][s]], tcw[d["obs_cam"][s]], d["obs_uv"][s], P))
        assert len(sm.iterations) == its[l] and np.max(np.abs(x - lm[l])) < 1e-9
    print("triangulate C: %.3f ms kernel, mean iterations %.2f" % (ms, its.mean()))


# ---- SelfGaussNewton (st17-ceres/src/include/solver.hpp:387-462) as one kernel ----------------------------------
def _numpy_gauss_newton(s, q, t, exact, max_it=10, tol=1e-8):
    from oracle import ba_oracle as bo, lie
    it = 0
    for it in range(max_it):
        H = np.zeros((6, 6)); g = np.zeros(6)
        R = lie.quat_to_rot(q)
        for P, uv in zip(s["points"], s["uv"]):
            pc = R.T @ (P - t)
            r = pc[:2] / pc[2] - uv
            e_R, e_t = bo.pnp_reference_jacobian(q, t, P)
            if exact:
                iz = 1.0 / pc[2]
                Pi = np.array([[iz, 0, -pc[0] * iz * iz], [0, iz, -pc[1] * iz * iz]])
                e_R = Pi @ lie.hat(pc)
            J = np.concatenate([e_R, e_t], axis=1)
            H += J.T @ J; g -= J.T @ r
        d = np.linalg.solve(H, g)
        q = lie.so3_plus(q, d[:3]); t = t + d[3:]
        if np.linalg.norm(d[:3]) + np.linalg.norm(d[3:]) < tol:
            break
    return q, t, it


@pytest.mark.gpu
def test_pnp_gauss_newton_kernel_equals_the_reference_self_gauss_newton(stba):
    s = stba.synth.pnp_scene()
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_kat.npz"))
    # (1) reference Jacobian form: same "iter num" and pose as SelfGaussNewton run from the reference source
    q, t, its, change, ms = stba.front.pnp_gauss_newton(s["points"], s["uv"], s["q_init"], s["t_init"])
    assert its == int(z["gn_iter_num"]) and change < 1e-8 and ms > 0
    assert min(np.abs(q - z["gn_pose"][:4]).max(), np.abs(q + z["gn_pose"][:4]).max()) < 1e-12 and np.abs(t - z["gn_pose"][4:]).max() < 1e-12
    # ... and as the oracle's restatement, iteration by iteration (capped runs)
    for cap in (1, 2, 3, 5):
        qg, tg, ig, _, _ = stba.front.pnp_gauss_newton(s["points"], s["uv"], s["q_init"], s["t_init"], max_iterations=cap)
        qo, to, _ = _numpy_gauss_newton(s, s["q_init"].copy(), s["t_init"].copy(), exact=False, max_it=cap)
        assert ig == cap and np.abs(qg - qo).max() < 1e-12 and np.abs(tg - to).max() < 1e-12
    # (2) exact Jacobian: also the published pose, in no more iterations
    q2, t2, its2, _, _ = stba.front.pnp_gauss_newton(s["points"], s["uv"], s["q_init"], s["t_init"], jacobian="exact")
    qo, to, io = _numpy_gauss_newton(s, s["q_init"].copy(), s["t_init"].copy(), exact=True)
    assert its2 == io and np.abs(q2 - qo).max() < 1e-12 and np.abs(t2 - to).max() < 1e-12
    assert min(np.abs(q2 - s["q_real"]).max(), np.abs(q2 + s["q_real"]).max()) < 1e-8 and np.abs(t2 - s["t_real"]).max() < 1e-8
    # (3) a batch: the same problem three times from three initial guesses, ragged observation counts
    n = len(s["points"])
    pts = np.concatenate([s["points"], s["points"][:n - 3], s["points"]]); uvs = np.concatenate([s["uv"], s["uv"][:n - 3], s["uv"]])
    ptr = np.array([0, n, 2 * n - 3, 3 * n - 3], np.int32)
    q0 = np.stack([s["q_init"], s["q_init"], z["gn_pose"][:4]]); t0 = np.stack([s["t_init"], s["t_init"] + 0.1, z["gn_pose"][4:]])
    qb, tb, ib, cb, _ = stba.front.pnp_gauss_newton(pts, uvs, q0, t0, ptr=ptr)
    assert np.abs(qb[0] - q).max() < 1e-15 and ib[0] == its and ib[2] <= 1
    for k in range(3):
        assert min(np.abs(qb[k] - s["q_real"]).max(), np.abs(qb[k] + s["q_real"]).max()) < 1e-7 and np.abs(tb[k] - s["t_real"]).max() < 1e-7


@pytest.mark.gpu
def test_device_hand_off_visibility_triangulate_ba(stba):
    """The three stages chained through DEVICE memory (torch CUDA tensors as the C ABI's array arguments): same
    index lists (bit-exact), same triangulated points and the same LM solve as the host-buffer route."""
    import torch
    sc = stba.synth.make_scene(20, 300, 1200)
    dev = torch.device("cuda", 0)
    q = torch.as_tensor(sc.true_cam_q, device=dev); t = torch.as_tensor(sc.true_cam_t, device=dev); p = torch.as_tensor(sc.true_lm, device=dev)
    vis_h = stba.front.visibility(sc.true_cam_q, sc.true_cam_t, sc.true_lm)
    vis_d = stba.front.visibility_device(q, t, p)
    for k in ("lm_deg", "cam_deg", "obs_cam", "obs_lm", "cam_lm"):
        assert np.array_equal(vis_d[k].cpu().numpy(), vis_h[k]), k
    assert np.array_equal(vis_d["obs_uv"].cpu().numpy(), vis_h["obs_uv"])
    keep = vis_h["lm_deg"] >= 2                       # triangulation needs two rays
    sel = keep[vis_h["obs_lm"]]
    remap = np.cumsum(keep) - 1
    oc = vis_h["obs_cam"][sel]; ol = remap[vis_h["obs_lm"][sel]].astype(np.int32); uv = vis_h["obs_uv"][sel]
    lm0 = sc.true_lm[keep] + 0.05
    lm_h, its_h, _, _, _ = stba.front.triangulate(sc.true_cam_q, sc.true_cam_t, lm0, oc, ol, uv)
    d_oc = torch.as_tensor(oc, device=dev); d_ol = torch.as_tensor(ol, device=dev); d_uv = torch.as_tensor(uv, device=dev)
    lm_d, its_d, _, _, _ = stba.front.triangulate_device(q, t, torch.as_tensor(lm0, device=dev), d_oc, d_ol, d_uv)
    assert np.array_equal(its_d.cpu().numpy(), its_h)
    assert np.array_equal(lm_d.cpu().numpy(), lm_h)
    cc = np.zeros(20, np.uint8); cc[0] = cc[-1] = 1
    q0 = torch.as_tensor(sc.cam_q, device=dev); t0 = torch.as_tensor(sc.cam_t, device=dev)
    e_d = stba.engine.BAEngine.from_device(q0, t0, lm_d, d_oc, d_ol, d_uv, cam_const=cc)
    e_h = stba.engine.BAEngine(sc.cam_q, sc.cam_t, lm_h, oc, ol, uv, cam_const=cc)
    s_d = e_d.solve(); s_h = e_h.solve()
    assert len(s_d.iterations) == len(s_h.iterations) and s_d.final_cost == s_h.final_cost
    qd, td, pd = e_d.get_state_device()
    qh, th, ph = e_h.get_state()
    assert np.array_equal(qd.cpu().numpy(), qh) and np.array_equal(pd.cpu().numpy(), ph)
    e_d.close(); e_h.close()
