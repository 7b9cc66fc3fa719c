"""Multi-GPU path (landmark-sharded, one NCCL all-reduce of the reduced camera system per
linearisation): 2 ranks must reproduce the single-GPU solve.  Needs >= 2 GPUs;
skipped otherwise (the driver's 1-GPU box runs the rest of the suite)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, size, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import stba
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    sc = stba.synth.make_scene(*size)
    lm, oc, ol, uv, _, (lo, hi) = stba.shard.shard_scene(sc.lm, sc.obs_cam, sc.obs_lm, sc.obs_uv, rank, world)
    eng = stba.engine.BAEngine(sc.cam_q, sc.cam_t, lm, oc, ol, uv, sc.cam_const, device=rank)
    ids = [stba.engine.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    eng.comm_init(rank, world, ids[0])
    s = eng.solve()
    cq, ct, l = eng.get_state()
    q.put((rank, lo, hi, cq, ct, l, [it["cost"] for it in s.iterations], s.termination_type))
    dist.barrier()
    eng.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("size", [(20, 300, 1200), (50, 5000, 50000)])
def test_two_rank_solve_equals_single_gpu(stba, size):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    sc = stba.synth.make_scene(*size)
    with stba.engine.BAEngine(sc.cam_q, sc.cam_t, sc.lm, sc.obs_cam, sc.obs_lm, sc.obs_uv, sc.cam_const) as e:
        s1 = e.solve()
        q1, t1, l1 = e.get_state()
    sk = socket.socket(); sk.bind(("127.0.0.1", 0)); port = sk.getsockname()[1]; sk.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, size, q)) for r in range(2)]
    [p.start() for p in procs]
    outs = sorted([q.get(timeout=300) for _ in range(2)], key=lambda o: o[0])
    [p.join(60) for p in procs]
    assert outs[0][7] == outs[1][7] == s1.termination_type
    assert outs[0][6] == outs[1][6]                                   # identical control flow on both ranks
    assert len(outs[0][6]) == len(s1.iterations)
    for a, b in zip(outs[0][6], s1.iterations):
        assert abs(a - b["cost"]) <= 1e-9 * b["cost"]
    assert np.array_equal(outs[0][3], outs[1][3]) and np.array_equal(outs[0][4], outs[1][4])   # replicated cameras
    assert np.max(np.abs(outs[0][3] - q1)) < 1e-9 and np.max(np.abs(outs[0][4] - t1)) < 1e-9
    lm = np.concatenate([outs[0][5], outs[1][5]])
    assert outs[0][2] == outs[1][1] and np.max(np.abs(lm - l1)) < 1e-9
