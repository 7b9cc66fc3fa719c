"""Multi-GPU host logic on CPU: landmark sharding + the allreduce structure of SURVEY.md §8e,
exercised with torch.distributed/gloo at world_size 2 (the oracle is only the checker)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_landmark_ranges_cover_and_balance(stba, scene_B):
    sc = scene_B
    for n in (1, 2, 4, 8):
        b = stba.shard.landmark_ranges(sc.obs_lm, sc.n_lm, n)
        assert b[0] == 0 and b[-1] == sc.n_lm and np.all(np.diff(b) >= 0)
        tot = 0
        for r in range(n):
            lm, oc, ol, uv, lc, (lo, hi) = stba.shard.shard_scene(sc.lm, sc.obs_cam, sc.obs_lm, sc.obs_uv, r, n)
            assert len(lm) == hi - lo and (len(ol) == 0 or (ol.min() >= 0 and ol.max() < hi - lo))
            assert np.all(np.diff(ol) >= 0)
            tot += len(oc)
            assert abs(len(oc) - sc.n_obs / n) <= 12       # balanced to within one landmark's degree
        assert tot == sc.n_obs


def test_ragged_and_empty_shards(stba):
    obs_lm = np.array([0, 0, 0, 0, 0, 0, 2, 2], dtype=np.int32)       # landmark 1 and 3 unobserved
    b = stba.shard.landmark_ranges(obs_lm, 4, 3)
    assert b[0] == 0 and b[-1] == 4 and np.all(np.diff(b) >= 0)
    seen = 0
    for r in range(3):
        out = stba.shard.shard_scene(np.zeros((4, 3)), np.zeros(8, np.int32), obs_lm, np.zeros((8, 2)), r, 3)
        seen += len(out[1])
    assert seen == 8
    b = stba.shard.landmark_ranges(np.zeros(0, np.int32), 0, 2)
    assert b.tolist() == [0, 0, 0]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import stba
    from oracle import ba_oracle as bo
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    sc = stba.synth.make_scene(8, 120, 480)
    lm, oc, ol, uv, _, (lo, hi) = stba.shard.shard_scene(sc.lm, sc.obs_cam, sc.obs_lm, sc.obs_uv, rank, world)
    r, Jc, Jl = bo.residual_jacobian(sc.cam_q, sc.cam_t, lm, oc, ol, uv)
    Hcc, gc, Hll, gl, W = bo.normal_blocks(r, Jc, Jl, oc, ol, sc.n_cam, len(lm), sc.cam_const)
    buf = torch.from_numpy(np.concatenate([Hcc.ravel(), gc.ravel(), [0.5 * np.sum(r * r)]]))
    dist.all_reduce(buf)                       # the [H_cc | g_c | cost] exchange of SURVEY.md §8e
    q.put((rank, buf.numpy().copy(), lo, hi, Hll, gl))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_allreduce_of_camera_blocks_equals_single_rank(stba):
    import torch.multiprocessing as mp
    from oracle import ba_oracle as bo
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    outs = sorted([q.get(timeout=180) for _ in range(2)], key=lambda o: o[0])
    [p.join(60) for p in procs]
    sc = stba.synth.make_scene(8, 120, 480)
    r, Jc, Jl = bo.residual_jacobian(sc.cam_q, sc.cam_t, sc.lm, sc.obs_cam, sc.obs_lm, sc.obs_uv)
    Hcc, gc, Hll, gl, W = bo.normal_blocks(r, Jc, Jl, sc.obs_cam, sc.obs_lm, sc.n_cam, sc.n_lm, sc.cam_const)
    want = np.concatenate([Hcc.ravel(), gc.ravel(), [0.5 * np.sum(r * r)]])
    assert np.array_equal(outs[0][1], outs[1][1])                     # identical on both ranks
    assert np.allclose(outs[0][1], want, rtol=1e-12, atol=1e-12)
    Hll_cat = np.concatenate([outs[0][4], outs[1][4]]); gl_cat = np.concatenate([outs[0][5], outs[1][5]])
    assert outs[0][3] == outs[1][2] and np.allclose(Hll_cat, Hll, rtol=1e-13) and np.allclose(gl_cat, gl, rtol=1e-13, atol=1e-15)
