#!/usr/bin/env python
"""Headline benchmark: LM iterations/sec on the synthetic bundle adjustment of BASELINE.json
configs[2] (1k cameras / 100k landmarks / 1M observations, Schur complement), see DESIGN.md §6.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl stba|reference] [--workload C|B]

One "step" = one full Ceres-style LM solve from the fixed initial guess x0 (IterationZero + every
trust-region iteration until the reference's own convergence test fires).  metric value = LM
iterations (trust-region step computations) per second over the K timed solves.
  value : state and observations already resident in HBM (x0 restored device-to-device per step)
  e2e   : the same through the reference-facing C ABI with HOST buffers: stba_ba_create (H2D of
          every input + index preprocessing) + solve + stba_ba_get_state (D2H) inside the timed region
The roofline object is for the kernel BASELINE.json names — the fused residual + Jacobian + J^T J
accumulation (lin_lm + lin_cam) — timed with CUDA events on the engine's stream, L2 flushed
between repetitions.  cpu_baseline / --impl reference time the oracle's C twin (a Ceres-equivalent
restatement: Ceres itself cannot be built offline) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "LM iterations/sec on synthetic BA (1k cam/100k pts/1M obs)"
UNIT = "iterations/s"


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
        np.savez(cache + ".tmp.npz", **d)
        os.replace(cache + ".tmp.npz", cache)
    except Exception:
        pass
    return d


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


def cpu_solve(d, threads):
    from oracle import ba_fast, ba_oracle
    ba_fast.set_num_threads(threads)
    t0 = time.perf_counter()
    out = ba_oracle.solve(d["cam_q"], d["cam_t"], d["lm"], d["obs_cam"], d["obs_lm"], d["obs_uv"], d["cam_const"], backend="c")
    dt = time.perf_counter() - t0
    s = out[3]
    # trust-region step computations: every recorded iteration after 0, plus the one that met the tolerance
    n_it = len(s.iterations) - 1 + (1 if "tolerance reached" in s.message and "Gradient" not in s.message else 0)
    return dt, n_it, s


def run_reference(args):
    """--impl reference: the CPU restatement (kind "port"), all host threads, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    d = load_scene(args.workload)
    threads = os.cpu_count() or 1
    for _ in range(args.warmup):
        cpu_solve(d, threads)
    tot_t, tot_it, last = 0.0, 0, None
    for _ in range(args.steps):
        dt, n_it, last = cpu_solve(d, threads)
        tot_t += dt; tot_it += n_it
    v = tot_it / tot_t
    sample = "%d full LM solves of the workload (%d iterations each) with the oracle's C twin + SciPy/OpenBLAS Cholesky" % (args.steps, tot_it // max(args.steps, 1))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "synthetic BA %s: %d cameras / %d landmarks / %d observations, Schur, seeds 20221105/20221106"
                       % (args.workload, len(d["cam_q"]), len(d["lm"]), len(d["obs_cam"]))},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                             "phase_s": last.phase_times, "note": "Ceres-equivalent CPU restatement; Ceres is not installable offline"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="stba", choices=["stba", "reference"])
    ap.add_argument("--workload", default="C", choices=["B", "C"])
    ap.add_argument("--dense", default="default", choices=["default", "own", "cusolver", "hybrid"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaled", type=int, default=10, help="copies of the workload for the streaming-size roofline (0/1 = skip)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import stba
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or stba.capi.device_count() == 0:
        raise SystemExit("bench.py needs a B200: libstba has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    d = load_scene(args.workload)
    n_cam, n_lm_total, n_obs_total = len(d["cam_q"]), len(d["lm"]), len(d["obs_cam"])
    if world > 1:
        lm, oc, ol, uv, _, _ = stba.shard.shard_scene(d["lm"], d["obs_cam"], d["obs_lm"], d["obs_uv"], rank, world)
    else:
        lm, oc, ol, uv = d["lm"], d["obs_cam"], d["obs_lm"], d["obs_uv"]

    def make_engine():
        return stba.engine.BAEngine(d["cam_q"], d["cam_t"], lm, oc, ol, uv, d["cam_const"], device=local)

    opt = stba.capi.Options()
    if args.dense != "default":
        opt.dense_backend = {"own": stba.capi.DENSE_OWN, "cusolver": stba.capi.DENSE_CUSOLVER, "hybrid": stba.capi.DENSE_HYBRID}[args.dense]
    eng = make_engine()
    if world > 1:
        ids = [stba.engine.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        eng.comm_init(rank, world, ids[0])
    eng.save_state()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        eng.restore_state()
        return eng.solve(opt)

    def n_iters(s):
        return s.num_iterations_run - 1 + (1 if "tolerance reached" in s.message and "Gradient" not in s.message else 0)

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
    elapsed = time.perf_counter() - t0
    launches = eng.launch_count() - launches0
    if world > 1:
        tt = torch.tensor([elapsed], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        elapsed = float(tt.item())
    clocks = sampler.finish() if sampler else None
    value = it_total / elapsed

    # ---- e2e through the C ABI with host buffers (create = H2D + preprocessing, solve, get_state = D2H) ----
    e2e = None
    if world == 1:
        h2d = sum(d[k].nbytes for k in ("cam_q", "cam_t", "lm", "obs_cam", "obs_lm", "obs_uv", "cam_const"))
        d2h = d["cam_q"].nbytes + d["cam_t"].nbytes + d["lm"].nbytes
        e2e_steps = max(2, min(args.steps, 5))
        te, ite = 0.0, 0
        # the step's inputs live in PINNED host memory (the contract's e2e definition); outputs come back into fresh NumPy arrays
        keep = {k: torch.from_numpy(np.ascontiguousarray(d[k])).pin_memory() for k in ("cam_q", "cam_t", "lm", "obs_cam", "obs_lm", "obs_uv", "cam_const")}
        hp = {k: v.numpy() for k, v in keep.items()}

        def make_engine_pinned():
            return stba.engine.BAEngine(hp["cam_q"], hp["cam_t"], hp["lm"], hp["obs_cam"], hp["obs_lm"], hp["obs_uv"], hp["cam_const"], device=local)

        for i in range(1 + e2e_steps):
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            with make_engine_pinned() as e2:
                s2 = e2.solve(opt)
                e2.get_state()
            if i:
                te += time.perf_counter() - t1; ite += n_iters(s2)
        e2e = {"value": ite / te, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "steps": e2e_steps, "ms_per_step": 1e3 * te / e2e_steps,
               "includes": "stba_ba_create (H2D from pinned host memory + validation + index preprocessing + pair structure) + stba_ba_solve + stba_ba_get_state (D2H) + stba_ba_destroy"}

    # ---- roofline of the contract kernel (linearise = lin_lm + lin_cam), L2 flushed between reps ----
    roofline = None; extra = {}
    if rank == 0:
        pk, how = peaks()
        ms = eng.time_phase("linearize", reps=20, flush_l2=True)[3:]
        ab = algorithmic_bytes(n_cam, len(lm), len(oc))
        ach = ab / (ms.mean() * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "k_lin_lm+k_lin_cam(+finish)", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
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
                                    "lin_lm_ms": float(big.time_phase("lin_lm", reps=6, flush_l2=True)[1:].mean()),
                                    "lin_cam_ms": float(big.time_phase("lin_cam", reps=6, flush_l2=True)[1:].mean()),
                                    "note": "camera table (%d x 112 B) exceeds shared memory: tiles come from L1/L2" % (k * nc)}
        big.close()
    if rank == 0 and world == 1:      # the remaining phases contain collectives when world > 1: single-GPU only
        fp64 = stba.engine.peak_fp64(local)
        n = 6 * int((d["cam_const"] == 0).sum())
        dense_name = {stba.capi.DENSE_OWN: "own", stba.capi.DENSE_CUSOLVER: "cusolver", stba.capi.DENSE_HYBRID: "hybrid"}[opt.dense_backend]
        dense_phase = "dense_" + dense_name
        dms = eng.time_phase(dense_phase, reps=5)[1:]
        extra["roofline_dense"] = {"bound": "fp64", "kernel": "reduced-camera Cholesky + solve (n=%d)" % n, "achieved": (n ** 3 / 3 + 2 * n * n) / (dms.mean() * 1e-3) / 1e12,
                                   "peak": fp64, "unit": "TFLOP/s", "peak_source": "stba_peak_fp64 (DFMA chains, measured in this run)",
                                   "launch_ms": float(dms.mean())}
        extra["roofline_dense"]["frac"] = extra["roofline_dense"]["achieved"] / fp64 if fp64 else None
        extra["phase_ms_isolated"] = {ph: float(eng.time_phase(ph, reps=5)[1:].mean()) for ph in ("lin_lm", "lin_cam", "schur", "dense_own", "dense_cusolver", "dense_hybrid", "backsub", "cost")}
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
    if rank == 0:
        extra["phase_ms_per_solve"] = {k: v for k, v in last.phase_ms.items()}
        extra["iterations_per_solve"] = n_iters(last)
        extra["termination"] = last.termination_type
        extra["final_cost"] = last.final_cost

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        dt, n_it, s = cpu_solve(d, threads)
        cpu = {"value": n_it / dt, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "1 full LM solve of the same workload (%d iterations, %.1f s) with the oracle's C twin + SciPy/OpenBLAS Cholesky" % (n_it, dt),
               "phase_s": s.phase_times, "final_cost": s.iterations[-1]["cost"]}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": "synthetic BA %s: %d cameras / %d landmarks / %d observations, Schur, seeds 20221105/20221106"
                           % (args.workload, n_cam, n_lm_total, n_obs_total),
                           "step": "one full LM solve from x0 (Ceres defaults, stops on its own tolerance tests)",
                           "parallelism": "landmark-sharded x%d, replicated reduced solve" % world if world > 1 else "single GPU",
                           "dense_backend": {stba.capi.DENSE_OWN: "own (hand-written DMMA Cholesky + substitutions)", stba.capi.DENSE_CUSOLVER: "cusolver (library potrf + potrs)",
                                             stba.capi.DENSE_HYBRID: "hybrid (cusolverDnDpotrf + own one-launch forward/backward substitutions)"}[opt.dense_backend],
                           "l2": "working set (E 144 MB + S 287 MB) exceeds L2; inputs re-read every iteration"},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
                "device_ms_per_step": dev_ms / args.steps}
        line.update(extra)
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
