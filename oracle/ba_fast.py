"""ctypes driver of the C twin (oracle/ba_oracle.c).  TEST INFRASTRUCTURE / CPU BASELINE ONLY.

`CSystem` has the interface of `ba_oracle.SchurSystem`, so `ba_oracle.solve(..., backend="c")`
runs the very same Ceres-faithful LM loop with the linearise / Schur / back-substitution steps
in compiled, multi-threaded C and the dense Cholesky in SciPy's LAPACK (OpenBLAS).
PARITY UNPINNED — see oracle/__init__.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import scipy.linalg

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_bp = C.POINTER(C.c_uint8)


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libba_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.ba_cost.restype = C.c_double
        L.ba_cost.argtypes = [C.c_int, C.c_int64, _dp, _dp, _dp, _ip, _ip, _dp]
        L.ba_linearize.restype = C.c_double
        L.ba_linearize.argtypes = [C.c_int, C.c_int, C.c_int64, _dp, _dp, _dp, _ip, _ip, _dp, _bp, _ip, _ip, _ip] + [_dp] * 7
        L.ba_schur.restype = None
        L.ba_schur.argtypes = [C.c_int, C.c_int, C.c_int, _ip, _ip, _ip] + [_dp] * 14
        L.ba_backsub.restype = C.c_double
        L.ba_backsub.argtypes = [C.c_int, _ip, _ip, _bp] + [_dp] * 9
        L.ba_num_threads.restype = C.c_int
        L.ba_set_num_threads.argtypes = [C.c_int]
        _LIB = L
    return _LIB


def set_num_threads(n):
    lib().ba_set_num_threads(int(n))


def num_threads():
    return int(lib().ba_num_threads())


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _b(a):
    return a.ctypes.data_as(_bp)


class Structure:
    """Integer structure shared by all linearisations of one problem."""

    def __init__(self, obs_cam, obs_lm, n_cam, n_lm, cam_const):
        self.obs_cam = np.ascontiguousarray(obs_cam, dtype=np.int32)
        self.obs_lm = np.ascontiguousarray(obs_lm, dtype=np.int32)
        self.n_cam, self.n_lm, self.n_obs = n_cam, n_lm, len(self.obs_cam)
        self.cam_const = np.ascontiguousarray(np.asarray(cam_const).astype(bool), dtype=np.uint8)
        self.lm_ptr = np.concatenate([[0], np.cumsum(np.bincount(self.obs_lm, minlength=n_lm))]).astype(np.int32)
        self.cam_ptr = np.concatenate([[0], np.cumsum(np.bincount(self.obs_cam, minlength=n_cam))]).astype(np.int32)
        self.cam_perm = np.argsort(self.obs_cam, kind="stable").astype(np.int32)
        free = self.cam_const == 0
        self.free_idx = np.nonzero(free)[0]
        self.free_of = -np.ones(n_cam, dtype=np.int32)
        self.free_of[self.free_idx] = np.arange(len(self.free_idx), dtype=np.int32)
        self.n_free = len(self.free_idx)


def cost(st, cam_q, cam_t, lm, obs_uv):
    q = np.ascontiguousarray(cam_q); t = np.ascontiguousarray(cam_t); p = np.ascontiguousarray(lm)
    return float(lib().ba_cost(st.n_cam, st.n_obs, _d(q), _d(t), _d(p), _i(st.obs_cam), _i(st.obs_lm), _d(obs_uv)))


class CSystem:
    """Linearisation at one point + scaled, damped Schur solve (interface of SchurSystem)."""

    def __init__(self, st, cam_q, cam_t, lm, obs_uv, sc=None, sl=None, jacobi_scaling=True):
        L = lib()
        self.st = st
        n_obs = st.n_obs
        q = np.ascontiguousarray(cam_q); t = np.ascontiguousarray(cam_t); p = np.ascontiguousarray(lm)
        self.r = np.empty((n_obs, 2)); self.Jc = np.empty((n_obs, 2, 6)); self.Jl = np.empty((n_obs, 2, 3))
        self.Hcc = np.empty((st.n_cam, 6, 6)); self.gc = np.empty((st.n_cam, 6))
        self.Hll = np.empty((st.n_lm, 3, 3)); self.gl = np.empty((st.n_lm, 3))
        self.cost = float(L.ba_linearize(st.n_cam, st.n_lm, n_obs, _d(q), _d(t), _d(p), _i(st.obs_cam), _i(st.obs_lm),
                                         _d(obs_uv), _b(st.cam_const), _i(st.lm_ptr), _i(st.cam_ptr), _i(st.cam_perm),
                                         _d(self.r), _d(self.Jc), _d(self.Jl), _d(self.Hcc), _d(self.gc), _d(self.Hll),
                                         _d(self.gl)))
        if sc is None:
            if jacobi_scaling:
                sc = 1.0 / (1.0 + np.sqrt(np.einsum("nii->ni", self.Hcc)))
                sl = 1.0 / (1.0 + np.sqrt(np.einsum("nii->ni", self.Hll)))
            else:
                sc = np.ones((st.n_cam, 6)); sl = np.ones((st.n_lm, 3))
        self.sc = np.ascontiguousarray(sc); self.sl = np.ascontiguousarray(sl)
        self._mcc = None
        self.times = {}

    def diagonal(self, opt):
        dc = np.clip(np.einsum("nii->ni", self.Hcc) * self.sc ** 2, opt.min_lm_diagonal, opt.max_lm_diagonal)
        dl = np.clip(np.einsum("nii->ni", self.Hll) * self.sl ** 2, opt.min_lm_diagonal, opt.max_lm_diagonal)
        return dc, dl

    def reduced_system(self, dc2, dl2):
        st = self.st
        n = 6 * st.n_free
        S = np.empty((n, n)); rhs = np.empty(n); Minv = np.empty((st.n_lm, 3, 3))
        dc2 = np.ascontiguousarray(dc2); dl2 = np.ascontiguousarray(dl2)
        lib().ba_schur(st.n_cam, st.n_lm, st.n_free, _i(st.obs_cam), _i(st.lm_ptr), _i(st.free_of), _d(self.r), _d(self.Jc),
                       _d(self.Jl), _d(self.Hcc), _d(self.gc), _d(self.Hll), _d(self.gl), _d(self.sc), _d(self.sl),
                       _d(dc2), _d(dl2), _d(S), _d(rhs), _d(Minv))
        return S, rhs, Minv

    def solve(self, dc2, dl2):
        import time
        st = self.st
        t0 = time.perf_counter()
        S, rhs, Minv = self.reduced_system(dc2, dl2)
        t1 = time.perf_counter()
        yc = np.zeros((st.n_cam, 6))
        if st.n_free:
            cf = scipy.linalg.cho_factor(S, lower=True, overwrite_a=True, check_finite=False)
            yc[st.free_idx] = scipy.linalg.cho_solve(cf, rhs, check_finite=False).reshape(-1, 6)
        t2 = time.perf_counter()
        yl = np.empty((st.n_lm, 3))
        self._mcc = float(lib().ba_backsub(st.n_lm, _i(st.obs_cam), _i(st.lm_ptr), _b(st.cam_const), _d(self.r), _d(self.Jc),
                                           _d(self.Jl), _d(self.gl), _d(self.sc), _d(self.sl), _d(Minv), _d(yc), _d(yl)))
        t3 = time.perf_counter()
        self.times = dict(schur=t1 - t0, dense=t2 - t1, backsub=t3 - t2)
        return yc, yl

    def model_cost_change(self, step_c, step_l):
        return self._mcc   # computed with the step of the last solve()
