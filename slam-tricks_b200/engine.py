"""SoA bundle-adjustment engine handle (`stba_ba_*`): what `ceres::Solve` does internally for the
reprojection problem of st20-g2o/src/include/test_ceres.h:98-152, resident on one B200."""
import ctypes as C

import numpy as np

from . import capi


class Summary:
    """`ceres::Solver::Summary` as the reference uses it (BriefReport, solver.hpp:290)."""

    def __init__(self, s, iterations):
        self.termination_type = capi.TERMINATION.get(s.termination_type, str(s.termination_type))
        self.message = s.message.decode(errors="replace")
        self.initial_cost = s.initial_cost
        self.final_cost = s.final_cost
        self.num_successful_steps = s.num_successful_steps
        self.num_unsuccessful_steps = s.num_unsuccessful_steps
        self.total_time_ms = s.total_time_ms
        self.phase_ms = dict(linearize=s.time_linearize_ms, schur=s.time_schur_ms, dense=s.time_dense_ms,
                             backsub=s.time_backsub_ms, cost=s.time_cost_ms)
        self.gpu_launches = s.gpu_launches
        self.iterations = iterations
        self.num_iterations_run = s.reserved

    def BriefReport(self):
        # the format printed in st17-ceres/img/release.png
        return ("Ceres Solver Report: Iterations: %d, Initial cost: %e, Final cost: %e, Termination: %s"
                % (len(self.iterations), self.initial_cost, self.final_cost, self.termination_type))

    brief_report = BriefReport


def run_solve(fn, handle, options, callback, max_records=512):
    """Shared driver of stba_ba_solve / stba_problem_solve."""
    opt = options if options is not None else capi.Options()
    recs = (capi.Iteration * max_records)()
    s = capi.SummaryStruct()
    s.iterations = C.cast(recs, C.POINTER(capi.Iteration))
    s.iterations_capacity = max_records
    if callback is not None:
        def _cb(it_ptr, _user):
            r = callback(it_ptr.contents.as_dict())
            if r is None or r is True:
                return capi.SOLVER_CONTINUE
            if r is False:
                return capi.SOLVER_ABORT
            return int(r)
        cfn = capi.ITERATION_CALLBACK(_cb)
    else:
        cfn = C.cast(None, capi.ITERATION_CALLBACK)
    capi.check(fn(handle, C.byref(opt), C.byref(s), cfn, None), fn.__name__)
    return Summary(s, [recs[i].as_dict() for i in range(s.num_iterations)])


class BAEngine:
    """Device-resident bundle-adjustment problem in the SoA layout of SURVEY.md §8d."""

    def __init__(self, cam_q, cam_t, lm, obs_cam, obs_lm, obs_uv, cam_const=None, lm_const=None, device=0,
                 linearize_only=False):
        L = capi.lib()
        self.n_cam, self.n_lm, self.n_obs = len(cam_q), len(lm), len(obs_cam)
        q = capi.as_f64(cam_q, (self.n_cam, 4)); t = capi.as_f64(cam_t, (self.n_cam, 3)); p = capi.as_f64(lm, (self.n_lm, 3))
        oc = np.ascontiguousarray(obs_cam, dtype=np.int32); ol = np.ascontiguousarray(obs_lm, dtype=np.int32)
        uv = capi.as_f64(obs_uv, (self.n_obs, 2))
        cc = np.ascontiguousarray(cam_const if cam_const is not None else np.zeros(self.n_cam), dtype=np.uint8)
        lc = np.ascontiguousarray(lm_const, dtype=np.uint8) if lm_const is not None else None
        self.cam_const = cc.astype(bool)
        self._h = C.c_void_p()
        capi.check(L.stba_ba_create_ex(C.byref(self._h), device, self.n_cam, self.n_lm, self.n_obs, capi.dptr(q), capi.dptr(t),
                                       capi.dptr(p), capi.iptr(oc), capi.iptr(ol), capi.dptr(uv), capi.bptr(cc), capi.bptr(lc),
                                       capi.CREATE_LINEARIZE_ONLY if linearize_only else 0), "stba_ba_create_ex")
        self._L = L

    @classmethod
    def from_device(cls, cam_q, cam_t, lm, obs_cam, obs_lm, obs_uv, cam_const=None, lm_const=None):
        """Device hand-off: the arrays are torch CUDA tensors (e.g. straight from `front.visibility_device` /
        `front.triangulate_device`); `stba_ba_create` copies them device-to-device.  cam_const / lm_const stay small
        host arrays (the free-camera map is built on the host)."""
        L = capi.lib()
        self = cls.__new__(cls)
        dev = cam_q.device
        self.n_cam, self.n_lm, self.n_obs = cam_q.shape[0], lm.shape[0], obs_cam.shape[0]
        cc = np.ascontiguousarray(cam_const if cam_const is not None else np.zeros(self.n_cam), dtype=np.uint8)
        lc = np.ascontiguousarray(lm_const, dtype=np.uint8) if lm_const is not None else None
        self.cam_const = cc.astype(bool)
        tp = lambda t, ct: C.cast(C.c_void_p(t.contiguous().data_ptr()), C.POINTER(ct))
        keep = [cam_q.contiguous(), cam_t.contiguous(), lm.contiguous(), obs_cam.contiguous(), obs_lm.contiguous(), obs_uv.contiguous()]
        self._h = C.c_void_p()
        capi.check(L.stba_ba_create_ex(C.byref(self._h), dev.index or 0, self.n_cam, self.n_lm, self.n_obs, tp(keep[0], C.c_double),
                                       tp(keep[1], C.c_double), tp(keep[2], C.c_double), tp(keep[3], C.c_int32), tp(keep[4], C.c_int32),
                                       tp(keep[5], C.c_double), capi.bptr(cc), capi.bptr(lc), 0), "stba_ba_create_ex")
        self._L = L
        return self

    def get_state_device(self):
        """State as torch CUDA tensors (device-to-device copy): (cam_q [n,4], cam_t [n,3], lm [n,3])."""
        import torch
        dev = torch.device("cuda", torch.cuda.current_device())
        q = torch.empty((self.n_cam, 4), dtype=torch.float64, device=dev); t = torch.empty((self.n_cam, 3), dtype=torch.float64, device=dev)
        p = torch.empty((self.n_lm, 3), dtype=torch.float64, device=dev)
        tp = lambda x: C.cast(C.c_void_p(x.data_ptr()), C.POINTER(C.c_double))
        capi.check(self._L.stba_ba_get_state(self._h, tp(q), tp(t), tp(p)), "stba_ba_get_state")
        return q, t, p

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.stba_ba_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- state ----
    def set_state(self, cam_q, cam_t, lm):
        q = capi.as_f64(cam_q, (self.n_cam, 4)); t = capi.as_f64(cam_t, (self.n_cam, 3)); p = capi.as_f64(lm, (self.n_lm, 3))
        capi.check(self._L.stba_ba_set_state(self._h, capi.dptr(q), capi.dptr(t), capi.dptr(p)), "stba_ba_set_state")

    def get_state(self):
        q = np.empty((self.n_cam, 4)); t = np.empty((self.n_cam, 3)); p = np.empty((self.n_lm, 3))
        capi.check(self._L.stba_ba_get_state(self._h, capi.dptr(q), capi.dptr(t), capi.dptr(p)), "stba_ba_get_state")
        return q, t, p

    def save_state(self):
        capi.check(self._L.stba_ba_save_state(self._h), "stba_ba_save_state")

    def restore_state(self):
        capi.check(self._L.stba_ba_restore_state(self._h), "stba_ba_restore_state")

    # ---- integer structure (bit-exact contract) ----
    def index_structures(self):
        out = dict(lm_deg=np.empty(self.n_lm, np.int32), cam_deg=np.empty(self.n_cam, np.int32),
                   lm_ptr=np.empty(self.n_lm + 1, np.int32), cam_ptr=np.empty(self.n_cam + 1, np.int32),
                   cam_perm=np.empty(self.n_obs, np.int32))
        capi.check(self._L.stba_ba_get_index(self._h, capi.iptr(out["lm_deg"]), capi.iptr(out["cam_deg"]),
                                             capi.iptr(out["lm_ptr"]), capi.iptr(out["cam_ptr"]),
                                             capi.iptr(out["cam_perm"])), "stba_ba_get_index")
        n = C.c_int64(0)
        capi.check(self._L.stba_ba_get_covis(self._h, None, C.byref(n)), "stba_ba_get_covis")
        keys = np.empty(max(n.value, 1), np.int64)
        capi.check(self._L.stba_ba_get_covis(self._h, keys.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(n)), "stba_ba_get_covis")
        out["covis"] = keys[:n.value]
        return out

    # ---- one linearisation ----
    def linearize(self):
        capi.check(self._L.stba_ba_linearize(self._h), "stba_ba_linearize")

    def blocks(self):
        """(Hcc [n_cam,6,6], gc [n_cam,6], Hll [n_lm,3,3], gl [n_lm,3], cost) of the last linearisation."""
        hcc = np.empty((self.n_cam, 21)); gc = np.empty((self.n_cam, 6)); hll = np.empty((self.n_lm, 6)); gl = np.empty((self.n_lm, 3))
        cost = C.c_double(0)
        capi.check(self._L.stba_ba_get_blocks(self._h, capi.dptr(hcc), capi.dptr(gc), capi.dptr(hll), capi.dptr(gl),
                                              C.byref(cost)), "stba_ba_get_blocks")
        iu = np.triu_indices(6)
        H = np.zeros((self.n_cam, 6, 6)); H[:, iu[0], iu[1]] = hcc; H[:, iu[1], iu[0]] = hcc
        iu3 = np.triu_indices(3)
        Hl = np.zeros((self.n_lm, 3, 3)); Hl[:, iu3[0], iu3[1]] = hll; Hl[:, iu3[1], iu3[0]] = hll
        return H, gc, Hl, gl, cost.value

    def reduced_system(self, radius, options=None, fetch=True):
        opt = options if options is not None else capi.Options()
        n = C.c_int32(0)
        capi.check(self._L.stba_ba_reduced_system(self._h, radius, C.byref(opt), None, None, C.byref(n)), "stba_ba_reduced_system")
        if not fetch:
            return n.value
        S = np.empty((n.value, n.value)); rhs = np.empty(n.value)
        capi.check(self._L.stba_ba_reduced_system(self._h, radius, C.byref(opt), capi.dptr(S), capi.dptr(rhs), C.byref(n)),
                   "stba_ba_reduced_system")
        return S.T.copy(), rhs     # column-major on the device -> row-major view

    def solve_step(self, dense_backend=capi.DENSE_OWN):
        yc = np.empty((self.n_cam, 6)); yl = np.empty((self.n_lm, 3)); mcc = C.c_double(0)
        capi.check(self._L.stba_ba_solve_step(self._h, dense_backend, capi.dptr(yc), capi.dptr(yl), C.byref(mcc)),
                   "stba_ba_solve_step")
        return yc, yl, mcc.value

    def solve(self, options=None, callback=None):
        return run_solve(self._L.stba_ba_solve, self._h, options, callback)

    # ---- timing ----
    PHASES = dict(linearize=0, lin_lm=1, lin_cam=2, schur=3, dense=4, backsub=5, cost=6, dense_own=7, dense_cusolver=8, dense_hybrid=9)

    def time_phase(self, phase, reps=10, flush_l2=False):
        ms = (C.c_float * reps)()
        capi.check(self._L.stba_ba_time_phase(self._h, self.PHASES[phase], reps, int(flush_l2), ms), "stba_ba_time_phase")
        return np.array(list(ms), dtype=np.float64)

    def launch_count(self):
        return int(self._L.stba_ba_launch_count(self._h))

    def use_comm(self, comm):
        capi.check(self._L.stba_ba_use_comm(self._h, comm._h), "stba_ba_use_comm")

    def comm_init(self, rank, nranks, unique_id):
        capi.check(self._L.stba_ba_comm_init(self._h, rank, nranks, unique_id), "stba_ba_comm_init")


def dense_cholesky_solve(S, rhs, backend=capi.DENSE_OWN, reps=1, device=0):
    """Solve S x = rhs on the GPU with one of the dense back ends (S symmetric: lower triangle read).
    Returns (x, info, ms[reps])."""
    S = np.asfortranarray(S, dtype=np.float64); rhs = capi.as_f64(rhs)
    n = S.shape[0]
    x = np.empty(n); info = C.c_int32(0); ms = (C.c_float * reps)()
    capi.check(capi.lib().stba_dense_cholesky_solve(device, backend, n, S.ctypes.data_as(C.POINTER(C.c_double)), capi.dptr(rhs),
                                                    capi.dptr(x), C.byref(info), reps, ms), "stba_dense_cholesky_solve")
    return x, info.value, np.array(list(ms))


def peak_fp64(device=0, reps=5):
    v = C.c_double(0)
    capi.check(capi.lib().stba_peak_fp64(device, reps, C.byref(v)), "stba_peak_fp64")
    return v.value


class Comm:
    """A NCCL communicator that outlives one problem (`stba_comm_create`); attach with `BAEngine.use_comm`."""

    def __init__(self, rank, nranks, unique_id, device=0):
        self._h = C.c_void_p()
        self._L = capi.lib()
        capi.check(self._L.stba_comm_create(C.byref(self._h), device, rank, nranks, unique_id), "stba_comm_create")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.stba_comm_destroy(self._h)
            self._h = None


def comm_unique_id():
    buf = C.create_string_buffer(capi.UNIQUE_ID_BYTES)
    capi.check(capi.lib().stba_comm_unique_id(buf), "stba_comm_unique_id")
    return buf.raw
