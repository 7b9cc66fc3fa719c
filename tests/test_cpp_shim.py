"""The C++ front door: include/ceres/ceres.h + examples/replay.cpp (headless replays of the
reference's Ceres call sequences).  The `bound` and `curve` replays are host plumbing
(BASELINE.json configs[0], "no GPU") and run on CPU; `ba` and `pnp` need the B200."""
import os
import struct
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "replay")


@pytest.fixture(scope="module")
def replay(stba):
    stba.capi.lib()
    src = os.path.join(ROOT, "examples", "replay.cpp")
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < os.path.getmtime(src):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-I" + os.path.join(ROOT, "include"), src, "-o", EXE,
                               "-L" + os.path.join(ROOT, "slam-tricks_b200"), "-lstba", "-Wl,-rpath,$ORIGIN/../slam-tricks_b200"])
    return EXE


REF = "/root/reference"
REF_EXE = os.path.join(ROOT, "examples", "replay_ref")
REF_BOUND = os.path.join(ROOT, "examples", "ref_ceres_bound")


@pytest.fixture(scope="module")
def replay_ref(stba):
    """examples/replay_ref: the reference's own source text compiled against the shim (tools/make_ref_replay.py).
    Rebuilt where /root/reference exists; the GPU box runs the prebuilt binary."""
    stba.capi.lib()
    if os.path.isdir(REF):
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "make_ref_replay.py")])
    if not os.path.exists(REF_EXE):
        pytest.skip("examples/replay_ref not built (needs /root/reference once)")
    return REF_EXE


def _trace(stdout):
    return [[float(x) for x in l.split()[1:]] for l in stdout.splitlines() if l.startswith("trace ")]


def test_bounds_demo_clamps_like_the_reference(replay):
    # st17-ceres/src/ceres_bound.cpp: unbounded minimum of (x-3)^2/2 is x = 3; with x in [-2, 2] it is 2
    out = subprocess.run([replay, "bound"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stdout + out.stderr
    x = float(out.stdout.split("unbounded: x = ")[1].split()[0])
    assert abs(x - 3.0) < 1e-6 and "bounded: x = 2.000000000" in out.stdout
    assert "gpu=0" in out.stdout
    # iterate for iterate the oracle's Ceres-faithful dense LM (Jacobian by ceres::Jet here, analytic there)
    from oracle import dense_lm
    _, s = dense_lm.solve(np.array([0.5]), lambda v: (np.array([v[0] - 3.0]), np.array([[1.0]])))
    tr = _trace(out.stdout.split("\nbounded:")[0])
    assert len(tr) == len(s.iterations)
    for a, b in zip(tr, s.iterations):
        assert abs(a[1] - b["cost"]) <= 1e-12 * max(b["cost"], 1e-30) + 1e-30 and abs(a[2] - b["trust_region_radius"]) <= 1e-9 * b["trust_region_radius"]


def test_reference_bounds_program_compiles_and_runs_unmodified(stba):
    """st17-ceres/src/ceres_bound.cpp is a complete program that only needs "ceres/ceres.h" and the author's logger: it is
    compiled from /root/reference as it is (tools/make_ref_replay.py) and must print x = 3, then x = 2."""
    stba.capi.lib()
    if os.path.isdir(REF):
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "make_ref_replay.py")])
    if not os.path.exists(REF_BOUND):
        pytest.skip("examples/ref_ceres_bound not built (needs /root/reference once)")
    out = subprocess.run([REF_BOUND], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stdout + out.stderr
    v = [float(l.split("]")[1]) for l in out.stdout.splitlines() if l.startswith("[ var ]")]
    assert len(v) == 4 and v[0] == 0 and abs(v[1] - 3.0) < 1e-5 and v[2] == 0 and v[3] == 2.0
    assert out.stdout.count("Termination: CONVERGENCE") == 2


def test_curve_fitting_plumbing_config(replay):
    # BASELINE.json configs[0]: 1 parameter block, 100 residuals, host path
    out = subprocess.run([replay, "curve"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "residual blocks=100" in out.stdout and "gpu=0" in out.stdout and "CONVERGENCE" in out.stdout
    # the same problem through the oracle's dense LM: same iterates (costs, radii, accept / reject), same answer
    from oracle import dense_lm
    state = 88172645463325252
    M = (1 << 64) - 1

    def uniform():
        nonlocal state
        state ^= (state << 13) & M; state ^= state >> 7; state ^= (state << 17) & M
        return (state >> 11) * (1.0 / 9007199254740992.0)
    xs, ys = [], []
    for i in range(100):
        x = -5.0 + 0.1 * i
        noise = 0.2 * (uniform() + uniform() + uniform() - 1.5)
        xs.append(x); ys.append(x * x + 2 * x + 3 + noise)
    xs, ys = np.array(xs), np.array(ys)
    J = np.stack([xs * xs, xs, np.ones_like(xs)], axis=1)
    sol, s = dense_lm.solve(np.zeros(3), lambda v: (J @ v - ys, J))
    tr = _trace(out.stdout)
    assert len(tr) == len(s.iterations)
    for a, b in zip(tr, s.iterations):
        assert abs(a[1] - b["cost"]) <= 1e-10 * b["cost"] and abs(a[2] - b["trust_region_radius"]) <= 1e-9 * b["trust_region_radius"]
        assert int(a[4]) == int(b["step_is_successful"])
    fin = [float(x) for x in [l for l in out.stdout.splitlines() if l.startswith("final ")][0].split()[1:]]
    assert np.max(np.abs(np.array(fin) - sol)) < 1e-9


def test_reference_gauss_newton_runs_from_its_own_source(stba, replay_ref, tmp_path):
    """SelfGaussNewton (st17-ceres/src/include/solver.hpp:387-462) compiled from the reference header: host only."""
    s = stba.synth.pnp_scene()
    fin, fout = str(tmp_path / "pnp.bin"), str(tmp_path / "gn.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("i", len(s["points"])))
        for a in (s["points"], s["uv"], s["q_init"], s["t_init"]):
            f.write(np.ascontiguousarray(a, np.float64).tobytes())
    out = subprocess.run([replay_ref, "gn", fin, fout], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stdout + out.stderr
    pose = np.fromfile(fout)
    assert min(abs(pose[:4] - s["q_real"]).max(), abs(pose[:4] + s["q_real"]).max()) < 1e-8 and abs(pose[4:] - s["t_real"]).max() < 1e-8
    assert "iter num:" in out.stdout


@pytest.mark.gpu
def test_reference_pnp_solvers_run_on_the_gpu_from_their_own_source(stba, replay_ref, tmp_path):
    """SolvePnPWith{DynamicAutoDiff, AutoDiff, SizedCostFunction} as written in solver.hpp:247-385: every ceres::Solve
    is recognised and runs in libstba.so; all recover the pose of st17-ceres/img/release.png."""
    s = stba.synth.pnp_scene()
    fin, fout = str(tmp_path / "pnp.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("i", len(s["points"])))
        for a in (s["points"], s["uv"], s["q_init"], s["t_init"]):
            f.write(np.ascontiguousarray(a, np.float64).tobytes())
    out = subprocess.run([replay_ref, "pnp", fin, fout], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "gpu_solves=3" in out.stdout and out.stdout.count("Termination: CONVERGENCE") == 3
    res = np.fromfile(fout).reshape(4, 7)
    for v in range(4):
        assert min(abs(res[v, :4] - s["q_real"]).max(), abs(res[v, :4] + s["q_real"]).max()) < 1e-5 and abs(res[v, 4:] - s["t_real"]).max() < 1e-5


@pytest.mark.gpu
def test_reference_ba_driver_runs_on_the_gpu_from_its_own_source(stba, replay_ref, tmp_path):
    """SolveWithCeresDynamicAutoDiff, LieLocalParameterization<SO3d> and ProjectFactor exactly as written in
    st20-g2o/src/include/test_ceres.h: one GPU solve, same result as the SoA engine."""
    sc = stba.synth.make_scene(20, 300, 1200)
    fin, fout = str(tmp_path / "scene.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("iii", sc.n_cam, sc.n_lm, sc.n_obs))
        for a in (sc.cam_q, sc.cam_t, sc.lm):
            f.write(np.ascontiguousarray(a, np.float64).tobytes())
        f.write(sc.obs_cam.astype(np.int32).tobytes()); f.write(sc.obs_lm.astype(np.int32).tobytes())
        f.write(np.ascontiguousarray(sc.obs_uv, np.float64).tobytes())
    out = subprocess.run([replay_ref, "ba", fin, fout], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "gpu_solves=1" in out.stdout and "Termination: CONVERGENCE" in out.stdout
    raw = np.fromfile(fout)
    q = raw[:4 * sc.n_cam].reshape(-1, 4); t = raw[4 * sc.n_cam:7 * sc.n_cam].reshape(-1, 3); lm = raw[7 * sc.n_cam:].reshape(-1, 3)
    with stba.engine.BAEngine(sc.cam_q, sc.cam_t, sc.lm, sc.obs_cam, sc.obs_lm, sc.obs_uv, sc.cam_const) as e:
        s = e.solve()
        q2, t2, l2 = e.get_state()
    assert "Iterations: %d" % len(s.iterations) in out.stdout
    assert np.max(np.abs(q - q2)) < 1e-9 and np.max(np.abs(t - t2)) < 1e-9 and np.max(np.abs(lm - l2)) < 1e-9


@pytest.mark.gpu
def test_reference_triangulation_functor_loop_and_batch(stba, replay_ref, tmp_path):
    """The `Triangulation` functor of sim_data.h:165-194: one ceres::Solve per landmark (host LM with Jet autodiff of the
    reference text) and the same problems handed over together (recognised, stba_triangulate): same points."""
    from oracle import front_oracle as fo
    sc = stba.synth.make_scene(20, 300, 1200)
    rng = np.random.default_rng(5)
    lm0 = sc.lm + rng.normal(0, 0.05, sc.lm.shape)
    fin, fout = str(tmp_path / "scene.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("iii", sc.n_cam, sc.n_lm, sc.n_obs))
        for a in (sc.cam_q, sc.cam_t, lm0):
            f.write(np.ascontiguousarray(a, np.float64).tobytes())
        f.write(sc.obs_cam.astype(np.int32).tobytes()); f.write(sc.obs_lm.astype(np.int32).tobytes())
        f.write(np.ascontiguousarray(sc.obs_uv, np.float64).tobytes())
    out = subprocess.run([replay_ref, "tri", fin, fout], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "loop: landmarks=%d gpu_solves=0" % sc.n_lm in out.stdout and "batch: landmarks=%d on_gpu=%d gpu=1" % (sc.n_lm, sc.n_lm) in out.stdout
    raw = np.fromfile(fout).reshape(2, sc.n_lm, 3)
    want, its, _, _ = fo.triangulate(sc.cam_q, sc.cam_t, lm0, sc.obs_cam, sc.obs_lm, sc.obs_uv)
    assert np.max(np.abs(raw[0] - want)) < 1e-8 and np.max(np.abs(raw[1] - want)) < 1e-8
    n_loop = int(out.stdout.split("loop:")[1].split("summary_iterations=")[1].split()[0])
    n_batch = int(out.stdout.split("batch:")[1].split("summary_iterations=")[1].split()[0])
    assert n_loop == n_batch == int(np.sum(its))


@pytest.mark.gpu
def test_ba_replay_matches_python_engine(stba, replay, tmp_path):
    sc = stba.synth.make_scene(20, 300, 1200)
    fin, fout = str(tmp_path / "scene.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("iii", sc.n_cam, sc.n_lm, sc.n_obs))
        for a in (sc.cam_q, sc.cam_t, sc.lm):
            f.write(np.ascontiguousarray(a, np.float64).tobytes())
        f.write(sc.obs_cam.astype(np.int32).tobytes()); f.write(sc.obs_lm.astype(np.int32).tobytes())
        f.write(np.ascontiguousarray(sc.obs_uv, np.float64).tobytes())
    out = subprocess.run([replay, "ba", fin, fout], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "gpu=1" in out.stdout and "live_state_changed=1" in out.stdout and "Termination: CONVERGENCE" in out.stdout
    raw = np.fromfile(fout, dtype=np.float64, count=7 * sc.n_cam + 3 * sc.n_lm + 1)
    q = raw[:4 * sc.n_cam].reshape(-1, 4); t = raw[4 * sc.n_cam:7 * sc.n_cam].reshape(-1, 3)
    lm = raw[7 * sc.n_cam:7 * sc.n_cam + 3 * sc.n_lm].reshape(-1, 3); final_cost = raw[-1]
    with stba.engine.BAEngine(sc.cam_q, sc.cam_t, sc.lm, sc.obs_cam, sc.obs_lm, sc.obs_uv, sc.cam_const) as e:
        s = e.solve()
        q2, t2, l2 = e.get_state()
    assert abs(final_cost - s.final_cost) <= 1e-12 * s.final_cost
    # the pointer front door numbers cameras by first appearance, so sums run in a different order:
    # equal to rounding, not bit for bit
    assert np.max(np.abs(q - q2)) < 1e-9 and np.max(np.abs(t - t2)) < 1e-9 and np.max(np.abs(lm - l2)) < 1e-9
    assert "Iterations: %d" % len(s.iterations) in out.stdout


@pytest.mark.gpu
def test_pnp_replays_recover_published_pose(stba, replay, tmp_path):
    s = stba.synth.pnp_scene()
    fin, fout = str(tmp_path / "pnp.bin"), str(tmp_path / "pnp_out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("i", len(s["points"])))
        for a in (s["points"], s["uv"], s["q_init"], s["t_init"]):
            f.write(np.ascontiguousarray(a, np.float64).tobytes())
    out = subprocess.run([replay, "pnp", fin, fout], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    res = np.fromfile(fout, dtype=np.float64).reshape(3, 7)
    for v in range(3):       # DynamicAutoDiff, AutoDiff, SizedCostFunction: all recover release.png's pose
        q, t = res[v, :4], res[v, 4:]
        assert min(abs(q - s["q_real"]).max(), abs(q + s["q_real"]).max()) < 1e-5
        assert abs(t - s["t_real"]).max() < 1e-5
    assert out.stdout.count("gpu=1") == 3


@pytest.mark.gpu
def test_reference_g2o_comparator_runs_on_the_gpu_from_its_own_source(stba, replay_ref, tmp_path):
    """SolveWithG2O with VertexCamera / VertexLandmark / EdgeProject exactly as written in st20-g2o/src/include/test_g2o.h,
    against include/compat/g2o: the graph is recognised by probing the user's oplusImpl / computeError and solved by the
    engine (no fixed vertex, 40 iterations at most; landmarks written back, cameras not — test_g2o.h:137-145)."""
    sc = stba.synth.make_scene(20, 300, 1200)
    fin, fout = str(tmp_path / "scene.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("iii", sc.n_cam, sc.n_lm, sc.n_obs))
        for a in (sc.cam_q, sc.cam_t, sc.lm):
            f.write(np.ascontiguousarray(a, np.float64).tobytes())
        f.write(sc.obs_cam.astype(np.int32).tobytes()); f.write(sc.obs_lm.astype(np.int32).tobytes())
        f.write(np.ascontiguousarray(sc.obs_uv, np.float64).tobytes())
    out = subprocess.run([replay_ref, "g2o", fin, fout], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    raw = np.fromfile(fout)
    q = raw[:4 * sc.n_cam].reshape(-1, 4); t = raw[4 * sc.n_cam:7 * sc.n_cam].reshape(-1, 3); lm = raw[7 * sc.n_cam:].reshape(-1, 3)
    assert np.array_equal(q, sc.cam_q) and np.array_equal(t, sc.cam_t)           # the reference copies the camera estimate into a temporary
    opt = stba.capi.Options(); opt.max_num_iterations = 40
    with stba.engine.BAEngine(sc.cam_q, sc.cam_t, sc.lm, sc.obs_cam, sc.obs_lm, sc.obs_uv, np.zeros(sc.n_cam, np.uint8)) as e:
        s = e.solve(opt)
        _, _, l2 = e.get_state()
    assert s.final_cost < 0.5 * s.initial_cost and np.max(np.abs(lm - l2)) < 1e-9
    assert out.stdout.count("cost") >= len(s.iterations)                          # setVerbose(true): one progress line per iteration
