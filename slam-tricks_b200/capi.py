"""ctypes binding of libstba.so (the C ABI in include/stba.h).

There is no fallback: if the shared library is missing, or no B200 is visible when a
compute entry point is called, this module raises.  Nothing here imports `oracle/`.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("STBA_LIB") or os.path.join(_HERE, "libstba.so")   # STBA_LIB: instrumented builds (profiling only)

OK = 0
ERR_NO_DEVICE = 3
TERMINATION = {0: "CONVERGENCE", 1: "NO_CONVERGENCE", 2: "FAILURE", 3: "USER_SUCCESS", 4: "USER_FAILURE"}
SOLVER_CONTINUE, SOLVER_ABORT, SOLVER_TERMINATE_SUCCESSFULLY = 0, 1, 2
DENSE_QR, SPARSE_SCHUR = 0, 1
MANIFOLD_EUCLIDEAN, MANIFOLD_SO3_QUAT, MANIFOLD_SO3_LOG = 0, 1, 2
DENSE_OWN, DENSE_CUSOLVER, DENSE_HYBRID = 0, 1, 2
CREATE_LINEARIZE_ONLY = 1
UNIQUE_ID_BYTES = 128


class StbaError(RuntimeError):
    def __init__(self, status, where):
        self.status = status
        super().__init__("%s failed: status %d (%s)" % (where, status, status_string(status)))


class Options(C.Structure):
    """`stba_options` == the `ceres::Solver::Options` fields the reference touches
    (test_ceres.h:133-145, solver.hpp:272-282) + the Ceres defaults."""
    _fields_ = [
        ("max_num_iterations", C.c_int32), ("max_num_consecutive_invalid_steps", C.c_int32),
        ("jacobi_scaling", C.c_int32), ("linear_solver_type", C.c_int32),
        ("update_state_every_iteration", C.c_int32), ("minimizer_progress_to_stdout", C.c_int32),
        ("num_threads", C.c_int32), ("dense_backend", C.c_int32),
        ("initial_trust_region_radius", C.c_double), ("max_trust_region_radius", C.c_double),
        ("min_trust_region_radius", C.c_double), ("min_relative_decrease", C.c_double),
        ("min_lm_diagonal", C.c_double), ("max_lm_diagonal", C.c_double),
        ("function_tolerance", C.c_double), ("gradient_tolerance", C.c_double),
        ("parameter_tolerance", C.c_double),
    ]

    def __init__(self, **kw):
        super().__init__()
        lib().stba_options_init(C.byref(self))
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError("stba_options has no field %r" % k)
            setattr(self, k, v)


class Iteration(C.Structure):
    _fields_ = [
        ("iteration", C.c_int32), ("step_is_valid", C.c_int32), ("step_is_successful", C.c_int32),
        ("reserved", C.c_int32), ("cost", C.c_double), ("cost_change", C.c_double),
        ("gradient_max_norm", C.c_double), ("gradient_norm", C.c_double), ("step_norm", C.c_double),
        ("relative_decrease", C.c_double), ("trust_region_radius", C.c_double),
        ("iteration_time_ms", C.c_double),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if n != "reserved"}


class SummaryStruct(C.Structure):
    _fields_ = [
        ("termination_type", C.c_int32), ("num_iterations", C.c_int32),
        ("num_successful_steps", C.c_int32), ("num_unsuccessful_steps", C.c_int32),
        ("initial_cost", C.c_double), ("final_cost", C.c_double), ("total_time_ms", C.c_double),
        ("time_linearize_ms", C.c_double), ("time_schur_ms", C.c_double), ("time_dense_ms", C.c_double),
        ("time_backsub_ms", C.c_double), ("time_cost_ms", C.c_double), ("gpu_launches", C.c_int64),
        ("message", C.c_char * 192), ("iterations", C.POINTER(Iteration)),
        ("iterations_capacity", C.c_int32), ("reserved", C.c_int32),
    ]


ITERATION_CALLBACK = C.CFUNCTYPE(C.c_int32, C.POINTER(Iteration), C.c_void_p)

_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_bp = C.POINTER(C.c_uint8)
_lp = C.POINTER(C.c_int64)

# every symbol include/stba.h declares: (restype, argtypes)
SIGNATURES = {
    "stba_options_init": (None, [C.POINTER(Options)]),
    "stba_version": (C.c_char_p, []),
    "stba_status_string": (C.c_char_p, [C.c_int]),
    "stba_device_count": (C.c_int, []),
    "stba_ba_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int32, C.c_int32, C.c_int64, _dp, _dp, _dp,
                                 _ip, _ip, _dp, _bp, _bp]),
    "stba_ba_create_ex": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int32, C.c_int32, C.c_int64, _dp, _dp, _dp,
                                    _ip, _ip, _dp, _bp, _bp, C.c_uint32]),
    "stba_ba_destroy": (None, [C.c_void_p]),
    "stba_ba_set_state": (C.c_int, [C.c_void_p, _dp, _dp, _dp]),
    "stba_ba_get_state": (C.c_int, [C.c_void_p, _dp, _dp, _dp]),
    "stba_ba_save_state": (C.c_int, [C.c_void_p]),
    "stba_ba_restore_state": (C.c_int, [C.c_void_p]),
    "stba_peak_fp64": (C.c_int, [C.c_int, C.c_int, _dp]),
    "stba_ba_get_index": (C.c_int, [C.c_void_p, _ip, _ip, _ip, _ip, _ip]),
    "stba_ba_get_covis": (C.c_int, [C.c_void_p, _lp, _lp]),
    "stba_ba_linearize": (C.c_int, [C.c_void_p]),
    "stba_ba_get_blocks": (C.c_int, [C.c_void_p, _dp, _dp, _dp, _dp, _dp]),
    "stba_ba_reduced_system": (C.c_int, [C.c_void_p, C.c_double, C.POINTER(Options), _dp, _dp, _ip]),
    "stba_ba_solve_step": (C.c_int, [C.c_void_p, C.c_int, _dp, _dp, _dp]),
    "stba_ba_solve": (C.c_int, [C.c_void_p, C.POINTER(Options), C.POINTER(SummaryStruct), ITERATION_CALLBACK, C.c_void_p]),
    "stba_ba_time_phase": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "stba_ba_launch_count": (C.c_int64, [C.c_void_p]),
    "stba_dense_cholesky_solve": (C.c_int, [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _ip, C.c_int, C.POINTER(C.c_float)]),
    "stba_comm_unique_id": (C.c_int, [C.c_char_p]),
    "stba_ba_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_char_p]),
    "stba_comm_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_char_p]),
    "stba_comm_destroy": (None, [C.c_void_p]),
    "stba_ba_use_comm": (C.c_int, [C.c_void_p, C.c_void_p]),
    "stba_visibility": (C.c_int, [C.c_int, C.c_int32, C.c_int32, _dp, _dp, _dp, C.c_double, C.c_double, C.c_int32, C.c_int64, _lp,
                                  _ip, _ip, _ip, _ip, _dp, _ip]),
    "stba_triangulate": (C.c_int, [C.c_int, C.c_int32, C.c_int32, C.c_int64, _dp, _dp, _dp, _ip, _ip, _dp, C.POINTER(Options),
                                   _ip, _dp, _ip, C.POINTER(C.c_float)]),
    "stba_pnp_gauss_newton": (C.c_int, [C.c_int, C.c_int32, _ip, _dp, _dp, _dp, _dp, C.c_int32, C.c_double, C.c_int32, _ip, _dp, C.POINTER(C.c_float)]),
    "stba_calib_initialize": (C.c_int, [C.c_int32, _ip, _dp, _dp, _dp, _dp, _dp]),
    "stba_calib_optimize": (C.c_int, [C.c_int, C.c_int32, _ip, _dp, _dp, _dp, _dp, _dp, C.c_int32, C.c_double, _ip, _dp, _dp, _lp]),
    "stba_calib_optimize_timed": (C.c_int, [C.c_int, C.c_int32, _ip, _dp, _dp, _dp, _dp, _dp, C.c_int32, C.c_double, _ip, _dp, _dp, _lp, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "stba_pg_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int32, C.c_int64, _dp, _dp, _ip, _ip, _dp, _dp]),
    "stba_pg_destroy": (None, [C.c_void_p]),
    "stba_pg_get_state": (C.c_int, [C.c_void_p, _dp, _dp]),
    "stba_pg_set_state": (C.c_int, [C.c_void_p, _dp, _dp]),
    "stba_pg_save_state": (C.c_int, [C.c_void_p]),
    "stba_pg_restore_state": (C.c_int, [C.c_void_p]),
    "stba_pg_time_linearize": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float)]),
    "stba_pg_linearize": (C.c_int, [C.c_void_p, _dp, _dp, _dp, _ip]),
    "stba_pg_solve": (C.c_int, [C.c_void_p, C.POINTER(Options), C.POINTER(SummaryStruct), ITERATION_CALLBACK, C.c_void_p]),
    "stba_problem_create": (C.c_int, [C.POINTER(C.c_void_p)]),
    "stba_problem_destroy": (None, [C.c_void_p]),
    "stba_problem_add_parameter_block": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "stba_problem_set_parameter_block_constant": (C.c_int, [C.c_void_p, C.c_void_p]),
    "stba_problem_set_parameter_lower_bound": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_double]),
    "stba_problem_set_parameter_upper_bound": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_double]),
    "stba_problem_add_reprojection": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                                C.POINTER(C.c_void_p), _dp]),
    "stba_problem_add_pnp": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, _dp, _dp]),
    "stba_problem_num_residual_blocks": (C.c_int, [C.c_void_p, _lp]),
    "stba_problem_num_parameter_blocks": (C.c_int, [C.c_void_p, _lp]),
    "stba_problem_solve": (C.c_int, [C.c_void_p, C.POINTER(Options), C.POINTER(SummaryStruct), ITERATION_CALLBACK, C.c_void_p]),
}


def lib():
    """Load libstba.so (once).  Raises if it has not been built — there is no Python/CPU path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libstba.so not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "or slam-tricks_b200/csrc/build.sh (expected at %s)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def status_string(status):
    return lib().stba_status_string(int(status)).decode()


def check(status, where):
    if status != OK:
        raise StbaError(status, where)


def device_count():
    return int(lib().stba_device_count())


def as_f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def dptr(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def iptr(a):
    return a.ctypes.data_as(_ip) if a is not None else None


def bptr(a):
    return a.ctypes.data_as(_bp) if a is not None else None
