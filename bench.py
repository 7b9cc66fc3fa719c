#!/usr/bin/env python
"""Headline benchmark: LM iterations/sec on the synthetic bundle adjustment of BASELINE.json
configs[2] (1k cameras / 100k landmarks / 1M observations, Schur complement), see DESIGN.md §6.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl stba|reference] [--workload C|B|PG|CALIB]

BA (workloads C, B).  One "step" = one full Ceres-style LM solve from the fixed initial guess x0
(IterationZero + every trust-region iteration until the reference's own convergence test fires).
metric value = LM iterations (trust-region step computations) per second over the K timed solves.
  value : state and observations already resident in HBM (x0 restored device-to-device per step)
  e2e   : the same through the reference-facing C ABI with HOST buffers: stba_ba_create (H2D of
          every input + index preprocessing) + solve + stba_ba_get_state (D2H) inside the timed region;
          with N > 1 GPUs every rank cuts its shard out of the host arrays, and rank 0 gathers the
          landmarks at the end (a NCCL communicator created once is re-attached to every problem)
The roofline object is for the kernel BASELINE.json names — the fused residual + Jacobian + J^T J
accumulation (lin_lm + lin_cam) — timed with CUDA events on the engine's stream, L2 flushed
between repetitions.  cpu_baseline / --impl reference time the oracle's C twin (a Ceres-equivalent
restatement: Ceres itself cannot be built offline) on the host cores — all of them, and ONE thread
(what the reference configures, st20-g2o/src/include/test_ceres.h:143).
--workload PG / CALIB: the same contract for BASELINE.json configs[4] (SE(3) pose graph, 10k poses /
40k edges) and configs[3] (Zhang calibration, 20 views x 88 corners).
"""
import os
import sys

if "reference" in sys.argv:
    # the reference arm owns its thread count: torchrun exports OMP_NUM_THREADS=1 when nproc > 1, which would
    # throttle OpenBLAS (the Cholesky of the CPU restatement) to one core
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_k] = str(os.cpu_count() or 1)

import argparse
import json
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "LM iterations/sec on synthetic BA (1k cam/100k pts/1M obs)"
METRIC_PG = "LM iterations/sec on SE(3) pose graph (10k poses/40k edges)"
METRIC_CALIB = "Gauss-Newton iterations/sec on Zhang calibration (20 views x 88 corners)"
UNIT = "iterations/s"
PG_N, PG_OFFSETS = 10000, (1, 2, 3, 4)


def load_scene(name):
    import stba
    cache = os.path.join("/tmp", "stba_scene_%s_%d.npz" % (name, stba.synth.SEED_DATA))
    keys = ("cam_q", "cam_t", "lm", "obs_cam", "obs_lm", "obs_uv", "cam_const")
    if os.path.exists(cache):
        try:
            z = np.load(cache)
            return {k: z[k] for k in keys}
        except Exception:
            pass
    sc = stba.synth.make_scene(*stba.synth.CONFIGS[name])
    d = {k: getattr(sc, k) for k in keys}
    try:
        np.savez(cache + ".tmp%d.npz" % os.getpid(), **d)
        os.replace(cache + ".tmp%d.npz" % os.getpid(), cache)
    except Exception:
        pass
    return d


def load_pose_graph():
    import stba
    cache = "/tmp/stba_pg_%d.npz" % PG_N
    keys = ("q0", "t0", "ei", "ej", "zq", "zt", "q_truth", "t_truth")
    if os.path.exists(cache):
        try:
            z = np.load(cache)
            return {k: z[k] for k in keys}
        except Exception:
            pass
    G = stba.synth.pose_graph(PG_N, offsets=PG_OFFSETS)
    try:
        np.savez(cache + ".tmp%d.npz" % os.getpid(), **G)
        os.replace(cache + ".tmp%d.npz" % os.getpid(), cache)
    except Exception:
        pass
    return G


def workload_string(name, d=None):
    if name == "PG":
        return "SE(3) pose graph: %d poses / %d edges (band offsets 1-4), Ceres-faithful LM, seed 20221108" % (PG_N, sum(PG_N - o for o in PG_OFFSETS))
    if name == "CALIB":
        return "Zhang calibration: 20 views x 88 corners, intrinsics + distortion + poses, Gauss-Newton (calib.cpp:282-422), seed 20221107"
    return ("synthetic BA %s: %d cameras / %d landmarks / %d observations, Schur, seeds 20221105/20221106"
            % (name, len(d["cam_q"]), len(d["lm"]), len(d["obs_cam"])))


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md recipe), sampled through
    NVML every 10 ms (the timed region is ~0.2 s; nvidia-smi takes longer than that to start)."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt, self.max_mhz = index, [], threading.Event(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def run(self):
        nv = self.nv
        if nv is None:
            return
        bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._halt.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((float(mhz), [n for n, b in bits.items() if r & b]))
            except Exception:
                pass
            self._halt.wait(0.01)

    def finish(self):
        self._halt.set()
        self.join(timeout=2)
        sm = [r[0] for r in self.rows]
        reasons = sorted({n for r in self.rows for n in r[1]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(self.rows), "how": "NVML, 10 ms period, during the timed region"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def measured_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the contract kernels, from the committed
    `ncu --set full` capture (profiles/lin_traffic.json, written by tools/ncu_summary.py traffic)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "lin_traffic.json")))
    except Exception:
        return None


def algorithmic_bytes(n_cam, n_lm, n_obs):
    # 24 B/obs + 96 B/landmark + 272 B/camera
    return 24 * n_obs + 96 * n_lm + 272 * n_cam


def threads_limited(n):
    """BLAS / OpenMP pools of this process at n threads (OpenBLAS does the Cholesky of the CPU restatement)."""
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=n)
    except Exception:
        import contextlib
        return contextlib.nullcontext()


# ------------------------------------------------------------------------------------------ CPU arms
def cpu_solve(d, threads):
    from oracle import ba_fast, ba_oracle
    ba_fast.set_num_threads(threads)
    with threads_limited(threads):
        t0 = time.perf_counter()
        out = ba_oracle.solve(d["cam_q"], d["cam_t"], d["lm"], d["obs_cam"], d["obs_lm"], d["obs_uv"], d["cam_const"], backend="c")
        dt = time.perf_counter() - t0
    s = out[3]
    # trust-region step computations: every recorded iteration after 0, plus the one that met the tolerance
    n_it = len(s.iterations) - 1 + (1 if "tolerance reached" in s.message and "Gradient" not in s.message else 0)
    return dt, n_it, s


def cpu_solve_pg(G, threads):
    from oracle import pg_oracle
    with threads_limited(threads):
        t0 = time.perf_counter()
        _, _, s = pg_oracle.solve(G["q0"], G["t0"], G["ei"], G["ej"], G["zq"], G["zt"])
        dt = time.perf_counter() - t0
    n_it = len(s.iterations) - 1 + (1 if "tolerance reached" in s.message and "Gradient" not in s.message else 0)
    return dt, n_it, s


def cpu_solve_calib(objs, imgs, threads, reps=1):
    from oracle import calib_oracle as co
    with threads_limited(threads):
        Hs = [co.homography(im, ob) for im, ob in zip(imgs, objs)]
        K4 = co.intrinsics_from_homographies(Hs)
        poses = co.extrinsics(K4, Hs)
        t0 = time.perf_counter()
        for _ in range(reps):
            out = co.total_optimization(K4, poses, objs, imgs)
        dt = (time.perf_counter() - t0) / reps
    return dt, out[3]["iterations"], out


def cpu_baseline_ba(d):
    """Rank 0, N = 1: the CPU restatement on all host cores AND on one thread (what the reference configures)."""
    threads = os.cpu_count() or 1
    dt, n_it, s = cpu_solve(d, threads)
    dt1, n_it1, s1 = cpu_solve(d, 1)
    return {"value": n_it / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "1 full LM solve of the same workload (%d iterations, %.1f s) with the oracle's C twin + SciPy/OpenBLAS Cholesky" % (n_it, dt),
            "phase_s": s.phase_times, "final_cost": s.iterations[-1]["cost"],
            "one_thread": {"value": n_it1 / dt1, "unit": UNIT, "cores": 1, "seconds": dt1, "phase_s": s1.phase_times,
                           "why": "the reference sets options.num_threads = 1 (st20-g2o/src/include/test_ceres.h:143)"},
            "note": "Ceres-equivalent CPU restatement; Ceres is not installable offline"}


def run_reference(args):
    """--impl reference: the CPU restatement (kind "port"), all host threads, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    w = args.workload
    if w in ("B", "C"):
        d = load_scene(w)
        fn, metric, cfg = (lambda: cpu_solve(d, threads)), METRIC, workload_string(w, d)
        what = "the oracle's C twin + SciPy/OpenBLAS Cholesky"
    elif w == "PG":
        G = load_pose_graph()
        fn, metric, cfg = (lambda: cpu_solve_pg(G, threads)), METRIC_PG, workload_string(w)
        what = "oracle/pg_oracle.py (NumPy + SciPy sparse Cholesky-free direct solve)"
    else:
        import stba
        objs, imgs, _, _ = stba.synth.calib_views()

        def fn():
            dt, its, _ = cpu_solve_calib(objs, imgs, threads)
            return dt, its, None
        metric, cfg, what = METRIC_CALIB, workload_string(w), "oracle/calib_oracle.py (literal NumPy restatement of calib.cpp:282-422)"
    for _ in range(args.warmup if w != "PG" else min(args.warmup, 1)):
        fn()
    steps = args.steps if w != "PG" else min(args.steps, 3)
    tot_t, tot_it, last = 0.0, 0, None
    for _ in range(steps):
        dt, n_it, last = fn()
        tot_t += dt; tot_it += n_it
    v = tot_it / tot_t
    sample = "%d full solves of the workload (%d iterations each) with %s" % (steps, tot_it // max(steps, 1), what)
    cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
           "phase_s": getattr(last, "phase_times", {}), "note": "Ceres-equivalent CPU restatement; Ceres is not installable offline",
           "threads_env": {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS")}}
    if w in ("B", "C"):
        dt1, n1, _ = cpu_solve(d, 1)
        cpu["one_thread"] = {"value": n1 / dt1, "unit": UNIT, "cores": 1, "seconds": dt1,
                             "why": "the reference sets options.num_threads = 1 (st20-g2o/src/include/test_ceres.h:143)"}
    line = {"impl": "reference", "metric": metric, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": cfg}, "cpu_baseline": cpu,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU arm: BA
def n_iters(s):
    return s.num_iterations_run - 1 + (1 if "tolerance reached" in s.message and "Gradient" not in s.message else 0)


def run_ba(args):
    import torch
    import stba
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    d = load_scene(args.workload)
    n_cam, n_lm_total, n_obs_total = len(d["cam_q"]), len(d["lm"]), len(d["obs_cam"])
    if world > 1:
        lm, oc, ol, uv, _, _ = stba.shard.shard_scene(d["lm"], d["obs_cam"], d["obs_lm"], d["obs_uv"], rank, world)
    else:
        lm, oc, ol, uv = d["lm"], d["obs_cam"], d["obs_lm"], d["obs_uv"]

    opt = stba.capi.Options()
    if args.dense != "default":
        opt.dense_backend = {"own": stba.capi.DENSE_OWN, "cusolver": stba.capi.DENSE_CUSOLVER, "hybrid": stba.capi.DENSE_HYBRID}[args.dense]
    eng = stba.engine.BAEngine(d["cam_q"], d["cam_t"], lm, oc, ol, uv, d["cam_const"], device=local)
    comm = None
    if world > 1:
        ids = [stba.engine.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        comm = stba.engine.Comm(rank, world, ids[0], device=local)      # created once, attached to every problem
        eng.use_comm(comm)
    eng.save_state()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def rank_max(x):
        if world == 1:
            return x
        tt = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    def step():
        eng.restore_state()
        return eng.solve(opt)

    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = eng.launch_count()
    barrier()
    t0 = time.perf_counter()
    it_total, dev_ms, last = 0, 0.0, None
    for _ in range(args.steps):
        last = step()
        it_total += n_iters(last)
        dev_ms += last.total_time_ms
    barrier()
    elapsed = rank_max(time.perf_counter() - t0)
    launches = eng.launch_count() - launches0
    clocks = sampler.finish() if sampler else None
    value = it_total / elapsed

    # ---- e2e through the C ABI with host buffers (create = H2D + preprocessing, solve, get_state = D2H) ----
    h2d_keys = ("cam_q", "cam_t", "lm", "obs_cam", "obs_lm", "obs_uv", "cam_const")
    e2e_steps = max(20, args.steps)
    # the step's inputs live in PINNED host memory (the contract's e2e definition); outputs come back into fresh NumPy arrays
    keep = {k: torch.from_numpy(np.ascontiguousarray(d[k])).pin_memory() for k in h2d_keys}
    hp = {k: v.numpy() for k, v in keep.items()}
    bounds = [stba.shard.shard_scene(d["lm"], d["obs_cam"], d["obs_lm"], d["obs_uv"], r, world)[5] for r in range(world)] if world > 1 else [(0, n_lm_total)]
    max_lm = max(hi - lo for lo, hi in bounds)
    te, ite, h2d, d2h = 0.0, 0, 0, 0
    for i in range(2 + e2e_steps):
        barrier()
        t1 = time.perf_counter()
        if world > 1:
            lm_i, oc_i, ol_i, uv_i, _, _ = stba.shard.shard_scene(hp["lm"], hp["obs_cam"], hp["obs_lm"], hp["obs_uv"], rank, world)
        else:
            lm_i, oc_i, ol_i, uv_i = hp["lm"], hp["obs_cam"], hp["obs_lm"], hp["obs_uv"]
        with stba.engine.BAEngine(hp["cam_q"], hp["cam_t"], lm_i, oc_i, ol_i, uv_i, hp["cam_const"], device=local) as e2:
            if comm is not None:
                e2.use_comm(comm)
            s2 = e2.solve(opt)
            q_out, t_out, l_out = e2.get_state()
        if world > 1:      # rank 0 ends up with every landmark (cameras are replicated)
            pad = torch.zeros((max_lm, 3), dtype=torch.float64, device="cuda")
            pad[:len(l_out)] = torch.from_numpy(l_out).cuda()
            got = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
            dist.gather(pad, got, dst=0)
            if rank == 0:
                l_all = np.concatenate([g[:hi - lo].cpu().numpy() for g, (lo, hi) in zip(got, bounds)])
                assert l_all.shape == (n_lm_total, 3)
        dt = rank_max(time.perf_counter() - t1)
        if i >= 2:
            te += dt; ite += n_iters(s2)
    h2d = sum(x.nbytes for x in (hp["cam_q"], hp["cam_t"], lm_i, oc_i, ol_i, uv_i, hp["cam_const"]))
    d2h = q_out.nbytes + t_out.nbytes + l_out.nbytes
    e2e = {"value": ite / te, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps, "ms_per_step": 1e3 * te / e2e_steps,
           "bytes_are": "per rank" if world > 1 else "total",
           "includes": ("shard of the host arrays + " if world > 1 else "") + "stba_ba_create (H2D from pinned host memory + validation + index preprocessing + pair structure) + stba_ba_solve + stba_ba_get_state (D2H)"
                       + (" + gather of the landmarks on rank 0" if world > 1 else "") + " + stba_ba_destroy"}

    # ---- roofline of the contract kernel (linearise = lin_lm + lin_cam), L2 flushed between reps ----
    roofline = None; extra = {}
    if rank == 0:
        pk, how = peaks()
        ms = eng.time_phase("linearize", reps=20, flush_l2=True)[3:]
        ab = algorithmic_bytes(n_cam, len(lm), len(oc))
        ach = ab / (ms.mean() * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "k_lin3 (one launch: landmark-major + camera-major pass side by side)", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": ach / pk["hbm_gbs"], "traffic": (measured_traffic() or {}).get("bytes_per_linearisation") if (args.workload == "C" and world == 1) else None,
                    "traffic_source": (measured_traffic() or {}).get("source"), "peak_source": how + " (MEASURED_PEAKS.json hbm_gbs, burst copy)", "algorithmic_bytes": ab,
                    "launch_ms": float(ms.mean()), "l2": "flushed between repetitions (192 MiB write sweep)"}
    if rank == 0 and world == 1 and args.scaled > 1:
        # the same kernels on `scaled` side-by-side copies of the workload: the observation stream no
        # longer fits the 126 MB L2, which is where an HBM-roofline fraction is meaningful
        k = args.scaled
        nc, nl = n_cam, n_lm_total
        rep = lambda a, shift: (a[None, :] + (np.arange(k) * shift)[:, None]).astype(np.int32).ravel()
        big = stba.engine.BAEngine(np.tile(d["cam_q"], (k, 1)), np.tile(d["cam_t"], (k, 1)), np.tile(d["lm"], (k, 1)),
                                   rep(d["obs_cam"], nc), rep(d["obs_lm"], nl), np.tile(d["obs_uv"], (k, 1)),
                                   np.tile(d["cam_const"], k), device=local, linearize_only=True)
        msb = big.time_phase("linearize", reps=12, flush_l2=True)[2:]
        abb = algorithmic_bytes(k * nc, k * nl, k * n_obs_total)
        extra["roofline_scaled"] = {"copies": k, "n_obs": k * n_obs_total, "algorithmic_bytes": abb, "launch_ms": float(msb.mean()),
                                    "achieved": abb / (msb.mean() * 1e-3) / 1e9, "unit": "GB/s",
                                    "frac": abb / (msb.mean() * 1e-3) / 1e9 / pk["hbm_gbs"],
                                    "three_launch_yardstick_ms": {"lin_lm2": float(big.time_phase("lin_lm", reps=6, flush_l2=True)[1:].mean()),
                                                                  "lin_cam2": float(big.time_phase("lin_cam", reps=6, flush_l2=True)[1:].mean())},
                                    "note": "camera table (%d x 112 B) exceeds shared memory: k_lin3 stages a window of 1024 cameras per landmark range" % (k * nc)}
        big.close()
    dense_names = {stba.capi.DENSE_OWN: "own", stba.capi.DENSE_CUSOLVER: "cusolver", stba.capi.DENSE_HYBRID: "hybrid"}
    if rank == 0 and world == 1:      # the remaining phases contain collectives when world > 1: single-GPU only
        fp64 = stba.engine.peak_fp64(local)
        n = 6 * int((d["cam_const"] == 0).sum())
        dense_phase = "dense_" + dense_names[opt.dense_backend]
        dms = eng.time_phase(dense_phase, reps=5)[1:]
        extra["roofline_dense"] = {"bound": "fp64", "kernel": "reduced-camera Cholesky + solve (n=%d)" % n, "achieved": (n ** 3 / 3 + 2 * n * n) / (dms.mean() * 1e-3) / 1e12,
                                   "peak": fp64, "unit": "TFLOP/s", "peak_source": "stba_peak_fp64 (DFMA chains, measured in this run)",
                                   "launch_ms": float(dms.mean())}
        extra["roofline_dense"]["frac"] = extra["roofline_dense"]["achieved"] / fp64 if fp64 else None
        lin_ms = float(eng.time_phase("linearize", reps=8, flush_l2=True)[2:].mean())
        extra["roofline_fp64_linearise"] = {"note": "the same launches against the FP64 pipe: ~135 DFMA per observation (DESIGN.md §3.1)",
                                            "achieved": 2 * 135 * len(oc) / (lin_ms * 1e-3) / 1e12, "peak": fp64, "unit": "TFLOP/s",
                                            "frac": (2 * 135 * len(oc) / (lin_ms * 1e-3) / 1e12) / fp64 if fp64 else None}
        extra["phase_ms_isolated"] = {ph: float(eng.time_phase(ph, reps=5)[1:].mean()) for ph in ("linearize", "lin_lm", "lin_cam", "schur", "dense_own", "dense_cusolver", "dense_hybrid", "backsub", "cost")}
    if rank == 0 and world == 1 and args.workload == "C":
        # the rows either side of the path through their C-ABI calls with host buffers
        t1 = time.perf_counter()
        vis = stba.front.visibility(d["cam_q"], d["cam_t"], d["lm"])
        t2 = time.perf_counter()
        _, its, _, _, tri_ms = stba.front.triangulate(d["cam_q"], d["cam_t"], d["lm"], d["obs_cam"], d["obs_lm"], d["obs_uv"])
        t3 = time.perf_counter()
        extra["front"] = {"visibility": {"pairs_tested": n_cam * n_lm_total, "visible": int(len(vis["obs_cam"])), "e2e_ms": 1e3 * (t2 - t1),
                                         "note": "stba_visibility: predicate + ordered compaction, host buffers in and out; the time is dominated by the device-to-host copy of the lists (28 B per visible pair into pageable memory); the kernels take 3.8 ms in total (ncu: profiles/r1_launches_front.md)"},
                          "triangulate": {"landmarks": n_lm_total, "observations": n_obs_total, "kernel_ms": tri_ms, "e2e_ms": 1e3 * (t3 - t2),
                                          "mean_lm_iterations": float(its.mean())}}
        # the same three stages chained through DEVICE memory (the C entry points take host or device pointers):
        # visibility -> triangulation -> stba_ba_create -> solve, nothing but counts and the summary crosses PCIe
        dq = torch.as_tensor(d["cam_q"], device="cuda"); dt = torch.as_tensor(d["cam_t"], device="cuda"); dp = torch.as_tensor(d["lm"], device="cuda")
        doc = torch.as_tensor(d["obs_cam"], device="cuda"); dol = torch.as_tensor(d["obs_lm"], device="cuda"); duv = torch.as_tensor(d["obs_uv"], device="cuda")
        stba.front.visibility_device(dq, dt, dp)                       # warm-up (stream-ordered pool, attributes)
        torch.cuda.synchronize()
        t4 = time.perf_counter()
        vis_d = stba.front.visibility_device(dq, dt, dp)
        torch.cuda.synchronize()
        t5 = time.perf_counter()
        lm_d, _, _, _, tri_ms_d = stba.front.triangulate_device(dq, dt, dp, doc, dol, duv)
        torch.cuda.synchronize()
        t6 = time.perf_counter()
        e_d = stba.engine.BAEngine.from_device(dq, dt, lm_d, doc, dol, duv, cam_const=d["cam_const"])
        torch.cuda.synchronize()
        t7 = time.perf_counter()
        e_d.close()
        extra["front"]["device_hand_off"] = {"visibility_ms": 1e3 * (t5 - t4), "visible": int(vis_d["obs_cam"].shape[0]), "triangulate_ms": 1e3 * (t6 - t5),
                                              "triangulate_kernel_ms": tri_ms_d, "ba_create_ms": 1e3 * (t7 - t6),
                                              "note": "torch CUDA tensors in and out of stba_visibility / stba_triangulate / stba_ba_create (cudaMemcpyDefault inside): "
                                                      "the 590 MB of lists never leave HBM; visibility runs its counting pass twice (count, then fill)"}
        del vis_d, lm_d
    if rank == 0:
        extra["phase_ms_per_solve"] = {k: v for k, v in last.phase_ms.items()}
        extra["iterations_per_solve"] = n_iters(last)
        extra["termination"] = last.termination_type
        extra["final_cost"] = last.final_cost

    cpu = cpu_baseline_ba(d) if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": {"workload": workload_string(args.workload, d)},
                "details": {"step": "one full LM solve from x0 (Ceres defaults, stops on its own tolerance tests)",
                            "parallelism": ("landmark-sharded x%d, ONE all-reduce of [S packed lower | rhs | H_cc | g_c | scalars] per linearisation, replicated reduced solve" % world) if world > 1 else "single GPU",
                            "dense_backend": {"own": "own (persistent DAG-scheduled DMMA Cholesky + one-launch substitution)", "cusolver": "cusolver (library potrf + potrs)",
                                              "hybrid": "hybrid (cusolverDnDpotrf + own one-launch forward/backward substitutions)"}[dense_names[opt.dense_backend]],
                            "l2": "working set (E 144 MB + S 287 MB) exceeds L2; inputs re-read every iteration"},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
                "device_ms_per_step": dev_ms / args.steps}
        line.update(extra)
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.barrier()
        comm.close()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------ GPU arm: pose graph
def run_pg(args):
    """BASELINE.json configs[4].  The banded direct solve does not shard: with N > 1 GPUs every rank
    runs an independent replica and the value is the aggregate ("replicas only")."""
    import torch
    import stba
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    G = load_pose_graph()
    n, m = len(G["q0"]), len(G["ei"])
    pgr = stba.posegraph.PoseGraph(G["q0"], G["t0"], G["ei"], G["ej"], G["zq"], G["zt"], device=local)
    pgr.save_state()

    def step():
        pgr.restore_state()
        return pgr.solve()

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(max(args.warmup, 3)):
        last = step()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    sync()
    t0 = time.perf_counter()
    it_total, launches = 0, 0
    for _ in range(args.steps):
        last = step()
        it_total += n_iters(last); launches += last.gpu_launches
    sync()
    elapsed = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([elapsed], dtype=torch.float64, device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX); elapsed = float(tt.item())
    clocks = sampler.finish() if sampler else None
    value = world * it_total / elapsed
    keys = ("q0", "t0", "ei", "ej", "zq", "zt")
    keep = {k: torch.from_numpy(np.ascontiguousarray(G[k])).pin_memory() for k in keys}
    hp = {k: v.numpy() for k, v in keep.items()}
    e2e_steps = max(20, args.steps)
    te, ite = 0.0, 0
    for i in range(2 + e2e_steps):
        sync()
        t1 = time.perf_counter()
        with stba.posegraph.PoseGraph(hp["q0"], hp["t0"], hp["ei"], hp["ej"], hp["zq"], hp["zt"], device=local) as p2:
            s2 = p2.solve()
            q_out, t_out = p2.get_state()
        if i >= 2:
            te += time.perf_counter() - t1; ite += n_iters(s2)
    e2e = {"value": world * ite / te, "unit": UNIT, "h2d_bytes_per_step": sum(hp[k].nbytes for k in keys), "d2h_bytes_per_step": q_out.nbytes + t_out.nbytes,
           "steps": e2e_steps, "ms_per_step": 1e3 * te / e2e_steps,
           "includes": "stba_pg_create (H2D from pinned host memory + incidence lists) + stba_pg_solve + stba_pg_get_state (D2H) + stba_pg_destroy"}
    if rank == 0:
        pk, how = peaks()
        lms = pgr.time_linearize(reps=20)[3:]
        # algorithmic bytes of the linearisation: per edge 2 x i32 + measurement 7 f64; per pose 7 f64 read + gradient 6 +
        # diagonal block 36 + (B) band blocks 36 each written
        B = max(PG_OFFSETS)
        ab = m * (8 + 56) + n * (56 + 48 + 288 * (1 + B))
        ach = ab / (lms.mean() * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "k_pg_linearize", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                    "traffic": None, "algorithmic_bytes": ab, "launch_ms": float(lms.mean()), "peak_source": how + " (MEASURED_PEAKS.json hbm_gbs)",
                    "note": "3 MB working set: L2-resident and launch-bound; the solve time is in the block-banded Cholesky (profiles/r1_launches_posegraph.md)"}
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            dt, n_it, _ = cpu_solve_pg(G, threads)
            cpu = {"value": n_it / dt, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": "1 full LM solve of the same graph (%d iterations, %.1f s) with oracle/pg_oracle.py (NumPy + SciPy sparse)" % (n_it, dt),
                   "note": "the reference has no pose-graph solver (SURVEY.md §8d): the oracle is the specification"}
        # the same graph with 1 % loop closures (100 edges between poses far apart along the chain): closure endpoints
        # become separators of the partitioned solve, the reduced system goes to the dense DAG Cholesky
        closures = None
        if world == 1:
            Gc = stba.synth.pose_graph(PG_N, offsets=PG_OFFSETS, closures=PG_N // 100)
            with stba.posegraph.PoseGraph(Gc["q0"], Gc["t0"], Gc["ei"], Gc["ej"], Gc["zq"], Gc["zt"], device=local) as pc:
                pc.save_state()
                for _ in range(2):
                    pc.restore_state(); sc = pc.solve()
                torch.cuda.synchronize()
                tc = time.perf_counter()
                itc = 0
                for _ in range(5):
                    pc.restore_state(); sc = pc.solve(); itc += n_iters(sc)
                torch.cuda.synchronize()
                tc = time.perf_counter() - tc
            closures = {"edges": int(len(Gc["ei"])), "loop_closures": PG_N // 100, "value": itc / tc, "unit": UNIT, "ms_per_solve": 1e3 * tc / 5,
                        "iterations_per_solve": n_iters(sc), "termination": sc.termination_type, "final_cost": sc.final_cost}
        line = {"metric": METRIC_PG, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True, "scaling": "weak" if world > 1 else "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": {"workload": workload_string("PG")},
                "details": {"step": "one full LM solve from the drifted odometry initial guess", "parallelism": "replicas only (x%d)" % world if world > 1 else "single GPU"},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
                "iterations_per_solve": n_iters(last), "termination": last.termination_type, "final_cost": last.final_cost,
                "with_loop_closures": closures}
        print(json.dumps(line))
    pgr.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------ GPU arm: Zhang calibration
def run_calib(args):
    """BASELINE.json configs[3] (single GPU; with N > 1 every rank runs a replica)."""
    import torch
    import stba
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    objs, imgs, _, _ = stba.synth.calib_views()
    views = list(zip(objs, imgs))
    base = stba.calib.CalibSolver(views=views, device=local).initialize()
    K0, P0 = base.intrinsics.copy(), base.imgPos.copy()

    def step():
        c = base
        c.intrinsics[:] = K0; c.distortion[:] = 0.0; c.imgPos[:] = P0
        c.totalOptimization()
        return c

    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    its, loop_ms, acc_ms, launches = 0, 0.0, 0.0, 0
    for _ in range(args.steps):
        c = step()
        its += len(c.update_norms); loop_ms += c.loop_ms; acc_ms += c.accumulate_ms; launches += c.gpu_launches
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    clocks = sampler.finish() if sampler else None
    if rank != 0:
        return
    pk, how = peaks()
    n_obs = int(base.view_ptr[-1]); V = base.cbsCount
    # accumulation kernel: per corner 2 + 2 f64 read, per view 12 f64 pose read + 136 block entries written
    ab = (n_obs * 32 + V * (96 + 136 * 8)) * its / args.steps
    ach = ab / (acc_ms / args.steps * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "k_calib_accumulate", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"], "traffic": None,
                "algorithmic_bytes": ab, "launch_ms": acc_ms / max(its, 1), "peak_source": how + " (MEASURED_PEAKS.json hbm_gbs)",
                "note": "80 KB problem: one CTA per view, latency-bound by construction (DESIGN.md §3.7)"}
    cpu = None
    if not args.no_cpu_baseline:
        dt, cpu_its, _ = cpu_solve_calib(objs, imgs, os.cpu_count() or 1, reps=3)
        cpu = {"value": cpu_its / dt, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
               "sample": "3 runs of totalOptimization on the same 20 x 88 input with oracle/calib_oracle.py (literal NumPy restatement of calib.cpp:282-422)"}
    line = {"metric": METRIC_CALIB, "value": its / (loop_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": loop_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_string("CALIB")},
            "details": {"step": "totalOptimization from the closed-form initialisation until |update| < 1e-8", "value_is": "Gauss-Newton loop with corners, parameters and poses resident in HBM (CUDA events)"},
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": its / wall, "unit": UNIT, "h2d_bytes_per_step": int(base.obj.nbytes + base.img.nbytes + base.view_ptr.nbytes + 9 * 8 + 96 * V),
                    "d2h_bytes_per_step": int(9 * 8 + 96 * V), "steps": args.steps, "ms_per_step": 1e3 * wall / args.steps,
                    "includes": "stba_calib_optimize with host buffers: allocation, H2D of corners / parameters / poses, Gauss-Newton loop, D2H, se3 logarithms"},
            "gpu_launches": launches, "clocks": clocks, "iterations_per_solve": len(c.update_norms), "final_cost": c.costs[-1] if c.costs else None}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="stba", choices=["stba", "reference"])
    ap.add_argument("--workload", default="C", choices=["B", "C", "PG", "CALIB"])
    ap.add_argument("--dense", default="default", choices=["default", "own", "cusolver", "hybrid"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaled", type=int, default=10, help="copies of the workload for the streaming-size roofline (0/1 = skip)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        return run_reference(args)
    import torch
    import stba
    if not torch.cuda.is_available() or stba.capi.device_count() == 0:
        raise SystemExit("bench.py needs a B200: libstba has no CPU path")
    if args.workload == "PG":
        return run_pg(args)
    if args.workload == "CALIB":
        return run_calib(args)
    return run_ba(args)


if __name__ == "__main__":
    main()
