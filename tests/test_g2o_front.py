"""g2o-shaped front door: replays `SolveWithG2O`, st20-g2o/src/include/test_g2o.h:94-147."""
import numpy as np
import pytest


def build(stba, sc, fix_ends):
    g2o = stba.g2o
    opt = g2o.SparseOptimizer()
    opt.setAlgorithm()
    cams, lms = [], []
    for i in range(sc.n_cam):                                   # test_g2o.h:106-113
        v = g2o.VertexCamera(); v.setId(i); v.setEstimate(sc.cam_q[i], sc.cam_t[i])
        if fix_ends and i in (0, sc.n_cam - 1):
            v.setFixed(True)
        opt.addVertex(v); cams.append(v)
    ptr = np.concatenate([[0], np.cumsum(np.bincount(sc.obs_lm, minlength=sc.n_lm))])
    for l in range(sc.n_lm):                                    # test_g2o.h:114-131
        v = g2o.VertexLandmark(); v.setId(l + sc.n_cam); v.setEstimate(sc.lm[l]); v.setMarginalized(True)
        opt.addVertex(v); lms.append(v)
        for o in range(ptr[l], ptr[l + 1]):
            e = g2o.EdgeProject(); e.setVertex(0, cams[sc.obs_cam[o]]); e.setVertex(1, v)
            e.setMeasurement(sc.obs_uv[o]); e.setInformation(np.eye(2)); opt.addEdge(e)
    return opt, cams, lms


def test_front_door_bookkeeping_on_cpu(stba, scene_small):
    opt, cams, lms = build(stba, scene_small, True)
    # chi2 = sum |Project(landmark) - measurement|^2 (EdgeProject::computeError, test_g2o.h:73-80) = 2 x the Ceres cost
    from oracle import ba_oracle as bo
    sc = scene_small
    assert abs(opt.activeChi2() - 2 * bo.cost_of(sc.cam_q, sc.cam_t, sc.lm, sc.obs_cam, sc.obs_lm, sc.obs_uv)) < 1e-9
    with pytest.raises(RuntimeError):
        opt.optimize(40)                                        # initializeOptimization() missing


@pytest.mark.gpu
def test_g2o_replay_matches_the_engine(stba, scene_small):
    sc = scene_small
    opt, cams, lms = build(stba, sc, True)
    opt.initializeOptimization()
    chi0 = opt.activeChi2()
    n_it = opt.optimize(40)                                     # test_g2o.h:135
    with stba.engine.BAEngine(sc.cam_q, sc.cam_t, sc.lm, sc.obs_cam, sc.obs_lm, sc.obs_uv, sc.cam_const) as e:
        s = e.solve(stba.capi.Options(max_num_iterations=40))
        q, t, p = e.get_state()
    assert n_it == len(s.iterations) - 1 and abs(opt.summary.final_cost - s.final_cost) <= 1e-12 * s.final_cost
    assert np.allclose(np.array([v.estimate()[1] for v in cams]), t, atol=1e-12)
    assert np.allclose(np.array([v.estimate() for v in lms]), p, atol=1e-12)
    assert abs(opt.activeChi2() - 2 * s.final_cost) <= 1e-9 * s.final_cost and opt.activeChi2() < 1e-2 * chi0


@pytest.mark.gpu
def test_gauge_free_graph_as_in_the_reference(stba, scene_small):
    opt, cams, lms = build(stba, scene_small, False)            # the reference fixes no vertex (test_g2o.h:106-113)
    opt.initializeOptimization()
    chi0 = opt.activeChi2()
    opt.optimize(40)
    assert opt.summary.termination_type in ("CONVERGENCE", "NO_CONVERGENCE") and opt.activeChi2() < 1e-2 * chi0
