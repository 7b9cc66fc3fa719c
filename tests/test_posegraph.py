"""SE(3) pose graph (BASELINE.json configs[4]).  The reference has no pose-graph solver, so the
oracle (oracle/pg_oracle.py) is the specification; it is pinned to what the reference does have: the recorded
tracks st4-kalman/output/{truth,obs}.csv (tests/golden/posegraph_st4.npz), the ATE metric of
pose_simulation.cpp:198-209 and the SE(3) identities of st23-lie-group-v2/doc.tex:862-997 (property tests)."""
import json
import os

import numpy as np
import pytest

from oracle import ba_oracle as bo
from oracle import pg_oracle as pg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def st4():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "posegraph_st4.npz")))


# ------------------------------------------------------------------ oracle (CPU)
def test_fixture_is_the_reference_track(st4):
    assert st4["q_truth"].shape == (1000, 4) and st4["t_obs"].shape == (1000, 3)
    assert st4["q_truth"][0].tolist() == [0, 0, 0, 1] and st4["t_truth"][0].tolist() == [0, 0, 0]   # truth.csv:3 "0,0,0,0,0,0,1"
    assert np.allclose(np.linalg.norm(st4["q_truth"], axis=1), 1.0, atol=1e-12)
    assert len(st4["ei"]) == 999 + 998 + 997 + 996 and int((st4["ej"] - st4["ei"]).max()) == 4
    # the obs track is the truth with accumulated odometry noise: ATE (pose_simulation.cpp:198-209) 0.445
    assert abs(pg.ate(st4["q_truth"], st4["t_truth"], st4["q0"], st4["t0"]) - 0.4450) < 1e-3


def test_se3_identities():
    rng = np.random.default_rng(0)
    xi = rng.normal(0, 0.7, (40, 6))
    q, t = pg.se3_exp(xi)
    assert np.allclose(pg.se3_log(q, t), xi, atol=1e-12)
    # Ad(T) xi = Log(T Exp(xi) T^-1)  (doc.tex adjoint identity), first order in xi
    T = pg.se3_exp(rng.normal(0, 0.8, 6))
    small = 1e-6 * rng.normal(0, 1, 6)
    lhs = pg.adjoint(*T) @ small
    qi, ti = pg.inverse(*T)
    rhs = pg.se3_log(*pg.compose(*pg.compose(*T, *pg.se3_exp(small)), qi, ti))
    assert np.allclose(lhs, rhs, atol=1e-11)
    # J_r^-1 by central differences: d Log(E Exp(d)) / d d at 0
    for _ in range(3):
        x = rng.normal(0, 0.8, 6)
        E = pg.se3_exp(x)
        num = np.zeros((6, 6))
        for k in range(6):
            d = np.zeros(6); d[k] = 1e-6
            num[:, k] = (pg.se3_log(*pg.compose(*E, *pg.se3_exp(d))) - pg.se3_log(*pg.compose(*E, *pg.se3_exp(-d)))) / 2e-6
        assert np.allclose(pg.jr_inv_se3(x), num, atol=1e-8)
    # series branch joins the closed form
    a = np.array([0.3, -0.2, 0.5, 6e-6, -5e-6, 4e-6]); b = a.copy(); b[3:] *= 1.2
    assert np.allclose(pg.jl_inv_se3(a), pg.jl_inv_se3(b), atol=1e-6)


def test_edge_jacobians_by_finite_differences():
    G = pg.make_graph(30)
    r, Ji, Jj = pg.residual_jacobians(G["q0"], G["t0"], G["ei"], G["ej"], G["zq"], G["zt"])
    for e in (0, 17, 60):
        for which, J in ((G["ei"][e], Ji), (G["ej"][e], Jj)):
            num = np.zeros((6, 6))
            for k in range(6):
                d = np.zeros((30, 6)); d[which, k] = 1e-6
                rp = pg.residuals(*pg.plus(G["q0"], G["t0"], d), G["ei"], G["ej"], G["zq"], G["zt"])[e]
                rm = pg.residuals(*pg.plus(G["q0"], G["t0"], -d), G["ei"], G["ej"], G["zq"], G["zt"])[e]
                num[:, k] = (rp - rm) / 2e-6
            assert np.allclose(J[e], num, atol=1e-7)


def test_oracle_reduces_the_drift_of_the_reference_track(st4):
    q, t, s = pg.solve(st4["q0"], st4["t0"], st4["ei"], st4["ej"], st4["zq"], st4["zt"])
    assert s.termination_type == "CONVERGENCE" and len(s.iterations) == len(st4["costs"]) == 7
    assert np.allclose([i["cost"] for i in s.iterations], st4["costs"], rtol=1e-9)
    assert np.allclose(q, st4["q_final"], atol=1e-10) and np.allclose(t, st4["t_final"], atol=1e-10)
    assert pg.ate(st4["q_truth"], st4["t_truth"], q, t) < 0.02 < 0.4 < pg.ate(st4["q_truth"], st4["t_truth"], st4["q0"], st4["t0"])


def test_product_side_generator_equals_the_oracles(stba):
    # slam-tricks_b200/synth.py pose_graph is self-contained (product code never imports oracle/) and bit-identical
    for n, off, cl in ((120, (1, 2, 3, 4), 0), (90, (1, 7, 16), 0), (150, (1, 2), 9)):
        a, b = stba.synth.pose_graph(n, offsets=off, closures=cl), pg.make_graph(n, offsets=off, closures=cl)
        assert all(np.array_equal(a[k], b[k]) for k in b)


def test_trajectory_csv_round_trip(stba, st4, tmp_path):
    p = str(tmp_path / "truth.csv")
    stba.posegraph.write_trajectory_csv(p, st4["q_truth"], st4["t_truth"])
    q, t = stba.posegraph.read_trajectory_csv(p)
    assert np.allclose(q, st4["q_truth"], atol=1e-5) and np.allclose(t, st4["t_truth"], atol=1e-5)      # %g: 6 significant digits, as the reference's dump


def test_pg_has_no_cpu_fallback(stba, st4):
    if stba.capi.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(stba.capi.StbaError) as e:
        stba.posegraph.PoseGraph(st4["q0"], st4["t0"], st4["ei"], st4["ej"], st4["zq"], st4["zt"])
    assert e.value.status == stba.capi.ERR_NO_DEVICE


# ------------------------------------------------------------------ CUDA path vs oracle
def _dense_blocks(G):
    r, Ji, Jj = pg.residual_jacobians(G["q0"], G["t0"], G["ei"], G["ej"], G["zq"], G["zt"])
    n = len(G["q0"])
    g = np.zeros((n, 6)); H = np.zeros((n, 6, 6))
    np.add.at(g, G["ei"], np.einsum("eka,ek->ea", Ji, r)); np.add.at(g, G["ej"], np.einsum("eka,ek->ea", Jj, r))
    np.add.at(H, G["ei"], np.einsum("eka,ekb->eab", Ji, Ji)); np.add.at(H, G["ej"], np.einsum("eka,ekb->eab", Jj, Jj))
    g[0] = 0.0; H[0] = np.eye(6)                       # pose 0 is constant
    return 0.5 * float(np.sum(r * r)), g, H


@pytest.mark.gpu
@pytest.mark.parametrize("offsets", [(1,), (1, 2, 3, 4), (1, 5, 16)])
def test_linearisation_matches_oracle(stba, offsets):
    G = pg.make_graph(120, offsets=offsets)
    with stba.posegraph.PoseGraph(G["q0"], G["t0"], G["ei"], G["ej"], G["zq"], G["zt"]) as p:
        cost, g, H, bw = p.linearize()
    want_cost, want_g, want_H = _dense_blocks(G)
    assert bw == max(offsets)
    assert abs(cost - want_cost) <= 1e-12 * want_cost
    assert np.allclose(g, want_g, rtol=1e-10, atol=1e-12) and np.allclose(H, want_H, rtol=1e-10, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("n,offsets", [(200, (1, 2, 3, 4)), (150, (1,)), (90, (1, 7, 16))])
def test_full_lm_solve_matches_oracle(stba, n, offsets):
    G = pg.make_graph(n, offsets=offsets)
    want_q, want_t, want = pg.solve(G["q0"], G["t0"], G["ei"], G["ej"], G["zq"], G["zt"])
    with stba.posegraph.PoseGraph(G["q0"], G["t0"], G["ei"], G["ej"], G["zq"], G["zt"]) as p:
        s = p.solve()
        q, t = p.get_state()
    assert s.termination_type == want.termination_type and s.message == want.message
    assert len(s.iterations) == len(want.iterations)
    assert [i["step_is_successful"] for i in s.iterations] == [int(i["step_is_successful"]) for i in want.iterations]
    assert np.allclose([i["cost"] for i in s.iterations], [i["cost"] for i in want.iterations], rtol=1e-8)
    # north-star bars: 1e-6 relative on the final residual norm, 1e-5 on the parameters
    assert abs(np.sqrt(2 * s.final_cost) - np.sqrt(2 * want.final_cost)) <= 1e-6 * np.sqrt(2 * want.final_cost)
    assert np.max(np.abs(q - want_q)) < 1e-7 and np.max(np.abs(t - want_t)) < 1e-7
    assert s.gpu_launches > 0


@pytest.mark.gpu
@pytest.mark.parametrize("n,offsets,parts", [(200, (1, 2, 3, 4), 5), (300, (1,), 7), (400, (1, 3, 8), 6), (131, (1, 2), 3)])
def test_partitioned_band_solve_is_the_same_solver(stba, monkeypatch, n, offsets, parts):
    """The partitioned (multi-SM) elimination is exact: forced on small graphs it must reproduce the oracle like the
    serial ring solver does (interiors of unequal length, first / last partitions without a left / right separator)."""
    monkeypatch.setenv("STBA_PG_PARTS", str(parts))
    G = pg.make_graph(n, offsets=offsets)
    want_q, want_t, want = pg.solve(G["q0"], G["t0"], G["ei"], G["ej"], G["zq"], G["zt"])
    with stba.posegraph.PoseGraph(G["q0"], G["t0"], G["ei"], G["ej"], G["zq"], G["zt"]) as p:
        s = p.solve()
        q, t = p.get_state()
    monkeypatch.setenv("STBA_PG_SERIAL", "1")
    monkeypatch.delenv("STBA_PG_PARTS")
    with stba.posegraph.PoseGraph(G["q0"], G["t0"], G["ei"], G["ej"], G["zq"], G["zt"]) as p:
        s1 = p.solve()
        q1, t1 = p.get_state()
    assert s.termination_type == want.termination_type and len(s.iterations) == len(want.iterations) == len(s1.iterations)
    assert np.allclose([i["cost"] for i in s.iterations], [i["cost"] for i in want.iterations], rtol=1e-8)
    assert np.max(np.abs(q - want_q)) < 1e-7 and np.max(np.abs(t - want_t)) < 1e-7
    assert np.max(np.abs(q - q1)) < 1e-9 and np.max(np.abs(t - t1)) < 1e-9
    assert s.gpu_launches > s1.gpu_launches          # six launches per solve instead of one


@pytest.mark.gpu
@pytest.mark.parametrize("n,offsets,closures,band", [(400, (1, 2, 3, 4), 12, None), (300, (1,), 25, None), (500, (1, 2), 1, None), (260, (1, 3), 40, None),
                                                     (200, (1, 2, 3, 4), 0, 2)])
def test_loop_closures_match_oracle(stba, monkeypatch, n, offsets, closures, band):
    """Edges longer than the band (|i - j| > 16): their endpoints become separator groups of the partitioned solve and the
    reduced system is solved densely — an exact elimination, so the LM iterates must reproduce the oracle (which solves
    the full sparse normal equations): same iterations, same accept / reject sequence, same poses.  Closures that share
    endpoints, closures at pose 0 (constant) and at the last pose, a single closure, and (last case) band edges forced
    into the closure set are covered by the seeds below."""
    if band is not None:
        monkeypatch.setenv("STBA_PG_CLOSURE_BAND", str(band))
    G = pg.make_graph(n, offsets=offsets, closures=closures)
    if closures >= 12:          # make sure the corner cases are in: pose 0, the last pose, a repeated pair
        G["ei"][-1], G["ej"][-1] = 0, n - 1
        G["ei"][-2], G["ej"][-2] = G["ei"][-3], G["ej"][-3]
        G["zq"][-1:], G["zt"][-1:] = pg.measurements_from(G["q_truth"], G["t_truth"], G["ei"][-1:], G["ej"][-1:])
        G["zq"][-2:-1], G["zt"][-2:-1] = pg.measurements_from(G["q_truth"], G["t_truth"], G["ei"][-2:-1], G["ej"][-2:-1])
    want_q, want_t, want = pg.solve(G["q0"], G["t0"], G["ei"], G["ej"], G["zq"], G["zt"])
    with stba.posegraph.PoseGraph(G["q0"], G["t0"], G["ei"], G["ej"], G["zq"], G["zt"]) as p:
        cost, g, H, bw = p.linearize()
        s = p.solve()
        q, t = p.get_state()
    want_cost, want_g, want_H = _dense_blocks(G)
    assert bw == (band if band is not None else max(offsets))
    assert abs(cost - want_cost) <= 1e-12 * want_cost and np.allclose(g, want_g, rtol=1e-10, atol=1e-12) and np.allclose(H, want_H, rtol=1e-10, atol=1e-12)
    assert s.termination_type == want.termination_type and s.message == want.message
    assert len(s.iterations) == len(want.iterations)
    assert [i["step_is_successful"] for i in s.iterations] == [int(i["step_is_successful"]) for i in want.iterations]
    assert np.allclose([i["cost"] for i in s.iterations], [i["cost"] for i in want.iterations], rtol=1e-8)
    assert abs(np.sqrt(2 * s.final_cost) - np.sqrt(2 * want.final_cost)) <= 1e-6 * np.sqrt(2 * want.final_cost)
    assert np.max(np.abs(q - want_q)) < 1e-7 and np.max(np.abs(t - want_t)) < 1e-7


@pytest.mark.gpu
def test_reference_track_fixture(stba, st4):
    with stba.posegraph.PoseGraph(st4["q0"], st4["t0"], st4["ei"], st4["ej"], st4["zq"], st4["zt"]) as p:
        s = p.solve()
        q, t = p.get_state()
    assert s.termination_type == str(st4["termination"]) and len(s.iterations) == len(st4["costs"])
    assert np.allclose([i["cost"] for i in s.iterations], st4["costs"], rtol=1e-8)
    assert np.max(np.abs(q - st4["q_final"])) < 1e-7 and np.max(np.abs(t - st4["t_final"])) < 1e-7
    assert pg.ate(st4["q_truth"], st4["t_truth"], q, t) < 0.02


@pytest.mark.gpu
def test_options_rejections_and_edge_cases(stba):
    G = pg.make_graph(60)
    opt = stba.capi.Options(max_num_iterations=2, jacobi_scaling=0)
    want_q, want_t, want = pg.solve(G["q0"], G["t0"], G["ei"], G["ej"], G["zq"], G["zt"], bo.LMOptions(max_num_iterations=2, jacobi_scaling=False))
    with stba.posegraph.PoseGraph(G["q0"], G["t0"], G["ei"], G["ej"], G["zq"], G["zt"]) as p:
        s = p.solve(opt)
        q, t = p.get_state()
    assert s.termination_type == "NO_CONVERGENCE" == want.termination_type and len(s.iterations) == 3
    assert np.max(np.abs(q - want_q)) < 1e-8 and np.max(np.abs(t - want_t)) < 1e-8
    # a tiny trust region forces rejected / shrinking steps through the same control flow
    opt = stba.capi.Options(initial_trust_region_radius=1e-3, max_num_iterations=6)
    want_q, want_t, want = pg.solve(G["q0"], G["t0"], G["ei"], G["ej"], G["zq"], G["zt"], bo.LMOptions(initial_trust_region_radius=1e-3, max_num_iterations=6))
    with stba.posegraph.PoseGraph(G["q0"], G["t0"], G["ei"], G["ej"], G["zq"], G["zt"]) as p:
        s = p.solve(opt)
    assert np.allclose([i["cost"] for i in s.iterations], [i["cost"] for i in want.iterations], rtol=1e-8)
    assert np.allclose([i["trust_region_radius"] for i in s.iterations], [i["trust_region_radius"] for i in want.iterations], rtol=1e-6)
    # offsets beyond the ring (round 1: STBA_ERR_UNSUPPORTED) are loop closures now: every pose is an endpoint here
    far = pg.make_graph(40, offsets=(1, 17))
    want_q, want_t, want = pg.solve(far["q0"], far["t0"], far["ei"], far["ej"], far["zq"], far["zt"])
    with stba.posegraph.PoseGraph(far["q0"], far["t0"], far["ei"], far["ej"], far["zq"], far["zt"]) as p:
        s = p.solve()
        q, t = p.get_state()
    assert len(s.iterations) == len(want.iterations) and np.max(np.abs(q - want_q)) < 1e-7 and np.max(np.abs(t - want_t)) < 1e-7
    # reversed edges
    with pytest.raises(stba.capi.StbaError):
        stba.posegraph.PoseGraph(G["q0"], G["t0"], G["ej"], G["ei"], G["zq"], G["zt"])
    # a single pose without edges is already solved
    with stba.posegraph.PoseGraph(G["q0"][:1], G["t0"][:1], [], [], np.zeros((0, 4)), np.zeros((0, 3))) as p:
        s = p.solve()
    assert s.termination_type == "CONVERGENCE" and s.final_cost == 0.0


@pytest.mark.gpu
def test_config_10k_poses_with_one_percent_loop_closures(stba):
    """BASELINE configs[4] size with 100 long-range closures (1 % of the poses) on top of the 39 990 band edges: the
    oracle's sparse LM (seconds) against the GPU path with closure endpoints as separators and a dense reduced solve."""
    G = pg.make_graph(10000, closures=100)
    want_q, want_t, want = pg.solve(G["q0"], G["t0"], G["ei"], G["ej"], G["zq"], G["zt"])
    with stba.posegraph.PoseGraph(G["q0"], G["t0"], G["ei"], G["ej"], G["zq"], G["zt"]) as p:
        s = p.solve()
        q, t = p.get_state()
    assert s.termination_type == want.termination_type and len(s.iterations) == len(want.iterations)
    assert [i["step_is_successful"] for i in s.iterations] == [int(i["step_is_successful"]) for i in want.iterations]
    assert np.allclose([i["cost"] for i in s.iterations], [i["cost"] for i in want.iterations], rtol=1e-8)
    assert abs(np.sqrt(2 * s.final_cost) - np.sqrt(2 * want.final_cost)) <= 1e-6 * np.sqrt(2 * want.final_cost)
    assert np.max(np.abs(q - want_q)) < 1e-6 and np.max(np.abs(t - want_t)) < 1e-6
    print("10k poses + 100 closures: %d iterations, %.1f ms, %d launches" % (len(s.iterations), s.total_time_ms, s.gpu_launches))


@pytest.mark.gpu
def test_config_10k_poses_40k_edges_against_golden(stba):
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "posegraph_10k.json")))
    G = pg.make_graph(10000)
    assert len(G["ei"]) == g["n_edges"] == 39990
    with stba.posegraph.PoseGraph(G["q0"], G["t0"], G["ei"], G["ej"], G["zq"], G["zt"]) as p:
        s = p.solve()
        q, t = p.get_state()
    assert s.termination_type == g["termination_type"] and len(s.iterations) == len(g["costs"])
    assert np.allclose([i["cost"] for i in s.iterations], g["costs"], rtol=1e-7)
    idx = np.asarray(g["sample_index"])
    assert np.max(np.abs(q[idx] - np.asarray(g["q_sample"]))) < 1e-5 and np.max(np.abs(t[idx] - np.asarray(g["t_sample"]))) < 1e-5
    ate = pg.ate(G["q_truth"], G["t_truth"], q, t)
    assert abs(ate - g["ate_final"]) < 1e-6 and ate < 0.25 * g["ate_initial"]
    print("pose graph 10k/40k: %d iterations, %.1f ms, %d launches, ATE %.3f -> %.3f" % (len(s.iterations), s.total_time_ms, s.gpu_launches, g["ate_initial"], ate))
