"""Zhang calibration (BASELINE.json configs[3]).  CPU tests pin the oracle to the
reference's fixture and check the host-side initialisation of the C ABI; `-m gpu` tests compare the
CUDA Gauss-Newton with the oracle.  Fixture: tests/golden/calib_fixture.json = the reference's own
st3-calibration/calib/1..9.txt (made by tests/golden/make_calib_golden.py)."""
import json
import os

import numpy as np
import pytest

from oracle import calib_oracle as co

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def fixture():
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "calib_fixture.json")))
    views = [(co.object_points(v["rows"], v["cols"], fx["cb_size"]), np.array(v["corners"])) for v in fx["views"]]
    return fx, views


# ------------------------------------------------------------------ oracle
def test_fixture_is_the_reference_input(fixture):
    fx, views = fixture
    assert len(views) == 9 and all(v["rows"] == 5 and v["cols"] == 8 for v in fx["views"])       # calib/1.txt:1 = "5,8"
    assert views[0][1][0].tolist() == [float(np.float32(747.984)), float(np.float32(850.560))]   # calib/1.txt:2, std::stof
    assert views[0][0][9].tolist() == [1 * 2.8e-2, 1 * 2.8e-2]                                    # (j, i) * cbSize


def test_oracle_reproduces_its_golden_outputs(fixture):
    fx, views = fixture
    objs, imgs = [v[0] for v in views], [v[1] for v in views]
    K0, p0, Hs, K, D, poses, info = co.solve(objs, imgs)
    g = fx["oracle"]
    assert np.allclose(K, g["intrinsics"], rtol=1e-9) and np.allclose(D, g["distortion"], rtol=1e-6, atol=1e-9)
    assert info["iterations"] == g["iterations"] == 8 and info["update_norms"][-1] < 1e-8 < info["update_norms"][-2]
    # Gauss-Newton from Zhang's closed form: the cost drops from 868 to 66.76 px^2 and stays there
    assert abs(info["costs"][0] - 868.4458158964181) < 1e-6 and abs(info["costs"][-1] - 66.7566002070605) < 1e-8


def test_oracle_se3_identities():
    rng = np.random.default_rng(0)
    for _ in range(20):
        xi = rng.normal(0, 0.7, 6)
        R, t = co.se3_exp(xi)
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-13) and np.allclose(co.se3_log(R, t), xi, atol=1e-12)
    R, t = co.se3_exp(np.array([1.0, 2.0, 3.0, 0, 0, 0]))
    assert np.allclose(R, np.eye(3)) and np.allclose(t, [1, 2, 3])


def test_oracle_jacobians_by_finite_differences(fixture):
    _, views = fixture
    obj, img = views[0]
    K4 = np.array([3000.0, 2990.0, 2000.0, 1500.0]); D5 = np.array([0.1, -0.2, 0.05, 1e-3, -2e-3])
    pose = np.array([-0.09, -0.05, 0.24, 0.03, -0.01, -0.013])
    e, Ji, Jd, Jp = co.residual_jacobian(K4, D5, pose, obj, img)
    h = 1e-6
    for k in range(4):
        d = np.zeros(4); d[k] = h
        num = (co.residual_jacobian(K4 + d, D5, pose, obj, img)[0] - co.residual_jacobian(K4 - d, D5, pose, obj, img)[0]) / (2 * h)
        assert np.allclose(Ji[:, k, :], num, atol=1e-5)
    for k in range(5):
        d = np.zeros(5); d[k] = h
        num = (co.residual_jacobian(K4, D5 + d, pose, obj, img)[0] - co.residual_jacobian(K4, D5 - d, pose, obj, img)[0]) / (2 * h)
        assert np.allclose(Jd[:, k, :], num, rtol=1e-5, atol=1e-4)
    for k in range(6):                                  # LEFT perturbation: T <- exp(d) T
        d = np.zeros(6); d[k] = h
        def at(dd):
            Rd, td = co.se3_exp(dd); Rp, tp = co.se3_exp(pose)
            return co.residual_jacobian(K4, D5, co.se3_log(Rd @ Rp, Rd @ tp + td), obj, img)[0]
        num = (at(d) - at(-d)) / (2 * h)
        assert np.allclose(Jp[:, k, :], num, rtol=1e-5, atol=1e-3)


def test_oracle_recovers_a_known_camera_at_the_size_BASELINE_asks():
    objs, imgs, K, D, poses = co.synthetic_views()                  # 20 views x 88 corners
    assert len(objs) == 20 and len(objs[0]) == 88
    out = co.solve(objs, imgs)
    assert np.allclose(out[3], K, rtol=5e-4) and np.allclose(out[5], poses, atol=2e-3) and out[6]["iterations"] <= 8


# ------------------------------------------------------------------ C ABI, host side (no GPU needed)
def test_initialisation_matches_the_oracle(stba, fixture):
    fx, views = fixture
    cs = stba.calib.CalibSolver(views=views).initialize()
    g = fx["oracle"]
    assert np.allclose(cs.intrinsics, g["init_intrinsics"], rtol=1e-9)
    assert np.allclose(cs.imgPos, g["init_poses"], atol=1e-9)
    for (obj, img), H in zip(views, cs.HomoMats):
        Ho = co.homography(img, obj)
        assert np.allclose(H / H[2, 2], Ho / Ho[2, 2], rtol=1e-6, atol=1e-6)
    objs, imgs, K, D, poses = co.synthetic_views()
    cs = stba.calib.CalibSolver(views=list(zip(objs, imgs))).initialize()
    K0, p0, _ = co.solve(objs, imgs)[:3]
    assert np.allclose(cs.intrinsics, K0, rtol=1e-8) and np.allclose(cs.imgPos, p0, atol=1e-8)


def test_corner_files_round_trip(stba, fixture, tmp_path):
    fx, views = fixture
    pts = views[2][1].reshape(5, 8, 2)
    stba.calib.write_corners(str(tmp_path / "3.txt"), pts)
    rows, cols, back = stba.calib.read_corners(str(tmp_path / "3.txt"))
    assert (rows, cols) == (5, 8) and np.array_equal(back, pts)     # 3 decimals survive float32
    cs = stba.calib.CalibSolver(str(tmp_path), 2.8e-2)
    assert cs.cbsCount == 1 and np.array_equal(cs.img, views[2][1]) and np.allclose(cs.obj, views[2][0])


def test_optimize_has_no_cpu_fallback(stba, fixture):
    if stba.capi.device_count() > 0:
        pytest.skip("GPU present")
    cs = stba.calib.CalibSolver(views=fixture[1]).initialize()
    with pytest.raises(stba.capi.StbaError) as e:
        cs.totalOptimization()
    assert e.value.status == stba.capi.ERR_NO_DEVICE


# ------------------------------------------------------------------ CUDA path vs oracle
@pytest.mark.gpu
def test_total_optimisation_matches_oracle_on_the_reference_fixture(stba, fixture):
    fx, views = fixture
    g = fx["oracle"]
    cs = stba.calib.CalibSolver(views=views).solve()
    assert len(cs.update_norms) == g["iterations"]                               # same stopping iteration
    assert np.allclose(cs.costs, g["costs"], rtol=1e-9)
    assert np.allclose(cs.update_norms[:4], g["update_norms"][:4], rtol=1e-6)
    assert np.allclose(cs.intrinsics, g["intrinsics"], rtol=1e-8)
    assert np.allclose(cs.distortion, g["distortion"], rtol=1e-5, atol=1e-8)
    assert np.allclose(cs.imgPos, g["poses"], atol=1e-8)
    assert cs.gpu_launches == 2 * g["iterations"]


@pytest.mark.gpu
def test_total_optimisation_20_views_88_corners(stba):
    objs, imgs, K, D, poses = co.synthetic_views()
    want = co.solve(objs, imgs)
    cs = stba.calib.CalibSolver(views=list(zip(objs, imgs))).solve()
    assert len(cs.update_norms) == want[6]["iterations"]
    assert np.allclose(cs.intrinsics, want[3], rtol=1e-8) and np.allclose(cs.distortion, want[4], rtol=1e-5, atol=1e-8)
    assert np.allclose(cs.imgPos, want[5], atol=1e-8) and np.allclose(cs.costs, want[6]["costs"], rtol=1e-9)
    assert np.allclose(cs.intrinsics, K, rtol=5e-4)                              # and it is the right camera


@pytest.mark.gpu
def test_total_optimisation_edge_cases(stba, fixture):
    _, views = fixture
    # ragged views (different corner counts, more than one staging tile) and a single-iteration run
    objs, imgs, *_ = co.synthetic_views(n_views=5, rows=13, cols=17)             # 221 corners > 128-corner tile
    ragged = [(o[: len(o) - 7 * i], m[: len(m) - 7 * i]) for i, (o, m) in enumerate(zip(objs, imgs))]
    want = co.solve([r[0] for r in ragged], [r[1] for r in ragged])
    cs = stba.calib.CalibSolver(views=ragged).solve()
    assert len(cs.update_norms) == want[6]["iterations"] and np.allclose(cs.intrinsics, want[3], rtol=1e-8)
    one = stba.calib.CalibSolver(views=views).initialize().totalOptimization(max_iterations=1)
    K0, p0, _ = co.solve([v[0] for v in views], [v[1] for v in views])[:3]
    w1 = co.total_optimization(K0, p0, [v[0] for v in views], [v[1] for v in views], max_iterations=1)
    assert np.allclose(one.intrinsics, w1[0], rtol=1e-9) and np.allclose(one.distortion, w1[1], rtol=1e-6, atol=1e-9)
    assert np.allclose(one.imgPos, w1[2], atol=1e-9)
