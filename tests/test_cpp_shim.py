"""The C++ front door: include/ceres/ceres.h + examples/replay.cpp (headless replays of the
reference's Ceres call sequences).  The `bound` and `curve` replays are host plumbing
(BASELINE.json configs[0], "no GPU") and run on CPU; `ba` and `pnp` need the B200."""
import os
import struct
import subprocess

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


def test_bounds_demo_clamps_like_the_reference(replay):
    # st17-ceres/src/ceres_bound.cpp: unbounded minimum of (x-3)^2/2 is x = 3; with x in [-2, 2] it is 2
    out = subprocess.run([replay, "bound"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "unbounded: x = 3.000000" in out.stdout and "bounded: x = 2.000000000" in out.stdout
    assert "gpu=0" in out.stdout


def test_curve_fitting_plumbing_config(replay):
    # BASELINE.json configs[0]: 1 parameter block, 100 residuals, host path
    out = subprocess.run([replay, "curve"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "residual blocks=100" in out.stdout and "gpu=0" in out.stdout and "CONVERGENCE" in out.stdout


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
