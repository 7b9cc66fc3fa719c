"""Python mirror of the `ceres::` API subset the reference calls, bound to the C ABI.

Same names, argument meaning and error behaviour as the reference's call sites so the parity
tests read like `SolveWithCeresDynamicAutoDiff` (st20-g2o/src/include/test_ceres.h:98-152) and
`SolvePnPWith*` (st17-ceres/src/include/solver.hpp:247-385).  Parameter blocks are NumPy float64
arrays; as in Ceres, a block IS its memory (`arr.ctypes.data`), values are read at Solve() and
written back in place.  Everything numeric happens in libstba.so on the GPU.
"""
import ctypes as C

import numpy as np

from . import capi
from .engine import run_solve

DENSE_QR = capi.DENSE_QR
SPARSE_SCHUR = capi.SPARSE_SCHUR
SOLVER_CONTINUE = capi.SOLVER_CONTINUE
SOLVER_ABORT = capi.SOLVER_ABORT
SOLVER_TERMINATE_SUCCESSFULLY = capi.SOLVER_TERMINATE_SUCCESSFULLY


class LieLocalParameterization:
    """`LieLocalParameterization<Sophus::SO3d>` (test_ceres.h:14-45): q <- q*exp(d); 4 -> 3."""
    manifold = capi.MANIFOLD_SO3_QUAT

    def GlobalSize(self):
        return 4

    def LocalSize(self):
        return 3


class LieR3LocalParameterization:
    """`LieR3LocalParameterization` (solver.hpp:63-94): x <- log(exp(x) exp(d)); 3 -> 3."""
    manifold = capi.MANIFOLD_SO3_LOG

    def GlobalSize(self):
        return 3

    def LocalSize(self):
        return 3


class ProjectFactor:
    """`ProjectFactor` (test_ceres.h:47-81): blocks (so3[4], pos[3], landmark[3]) -> 2 residuals."""
    block_sizes = (4, 3, 3)

    def __init__(self, feature):
        self.feature = np.asarray(feature, dtype=np.float64).reshape(2)

    @staticmethod
    def Create(feature):
        return ProjectFactor(feature)


class PnPFactor:
    """`PnPDynamicAutoDiffFunctor` / `PnPAutoDiffFunctor` / `PnPSizedCostFunction`
    (solver.hpp:96-212): known 3-D point + feature; blocks (rotation, position) -> 2 residuals.
    The rotation block is a quaternion (4) or, for the Sized variant, so3.log() (3)."""

    def __init__(self, point, feature, rotation_size=4):
        self.point = np.asarray(point, dtype=np.float64).reshape(3)
        self.feature = np.asarray(feature, dtype=np.float64).reshape(2)
        self.block_sizes = (rotation_size, 3)


def _block(a):
    if not isinstance(a, np.ndarray) or a.dtype != np.float64 or not a.flags["C_CONTIGUOUS"]:
        raise TypeError("parameter blocks must be C-contiguous float64 NumPy arrays (their memory is the block)")
    return a


class Problem:
    """`ceres::Problem` subset: AddResidualBlock, AddParameterBlock, SetParameterBlockConstant,
    SetParameterLowerBound/UpperBound."""

    def __init__(self):
        self._L = capi.lib()
        self._h = C.c_void_p()
        capi.check(self._L.stba_problem_create(C.byref(self._h)), "stba_problem_create")
        self._keep = {}          # address -> array (keeps user memory alive, like Ceres assumes)
        self._pending = []       # batched reprojection factors

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.stba_problem_destroy(self._h)
            self._h = None

    def _addr(self, a):
        a = _block(a)
        self._keep[a.ctypes.data] = a
        return a.ctypes.data

    def AddParameterBlock(self, values, size, local_parameterization=None):
        self._flush()
        m = local_parameterization.manifold if local_parameterization is not None else capi.MANIFOLD_EUCLIDEAN
        capi.check(self._L.stba_problem_add_parameter_block(self._h, self._addr(values), int(size), m),
                   "stba_problem_add_parameter_block")

    def SetParameterBlockConstant(self, values):
        self._flush()
        capi.check(self._L.stba_problem_set_parameter_block_constant(self._h, self._addr(values)),
                   "stba_problem_set_parameter_block_constant")

    def SetParameterLowerBound(self, values, index, bound):
        self._flush()
        capi.check(self._L.stba_problem_set_parameter_lower_bound(self._h, self._addr(values), index, bound),
                   "stba_problem_set_parameter_lower_bound")

    def SetParameterUpperBound(self, values, index, bound):
        self._flush()
        capi.check(self._L.stba_problem_set_parameter_upper_bound(self._h, self._addr(values), index, bound),
                   "stba_problem_set_parameter_upper_bound")

    def AddResidualBlock(self, cost_function, loss_function, parameter_blocks):
        if loss_function is not None:
            raise NotImplementedError("the reference always passes nullptr as the loss function")
        blocks = [_block(b) for b in parameter_blocks]
        if [b.size for b in blocks] != list(cost_function.block_sizes):
            raise ValueError("parameter block sizes %s do not match the cost function %s"
                             % ([b.size for b in blocks], list(cost_function.block_sizes)))
        if isinstance(cost_function, ProjectFactor):
            self._pending.append((self._addr(blocks[0]), self._addr(blocks[1]), self._addr(blocks[2]),
                                  cost_function.feature))
        elif isinstance(cost_function, PnPFactor):
            self._flush()
            m = capi.MANIFOLD_SO3_LOG if blocks[0].size == 3 else capi.MANIFOLD_SO3_QUAT
            pt = capi.as_f64(cost_function.point, (1, 3)); uv = capi.as_f64(cost_function.feature, (1, 2))
            capi.check(self._L.stba_problem_add_pnp(self._h, 1, self._addr(blocks[0]), self._addr(blocks[1]), m,
                                                    capi.dptr(pt), capi.dptr(uv)), "stba_problem_add_pnp")
        else:
            raise NotImplementedError("only the reprojection factor family of the reference runs on the GPU")

    def AddReprojectionBlocks(self, so3_blocks, pos_blocks, landmark_blocks, uv):
        """Batched form of n x AddResidualBlock(ProjectFactor::Create(uv[i]), nullptr, {...})."""
        for a, b, c, f in zip(so3_blocks, pos_blocks, landmark_blocks, np.asarray(uv, dtype=np.float64)):
            self._pending.append((self._addr(a), self._addr(b), self._addr(c), f))

    def _flush(self):
        if not self._pending:
            return
        n = len(self._pending)
        arr = lambda k: (C.c_void_p * n)(*[p[k] for p in self._pending])
        uv = capi.as_f64(np.stack([p[3] for p in self._pending]), (n, 2))
        capi.check(self._L.stba_problem_add_reprojection(self._h, n, arr(0), arr(1), arr(2), capi.dptr(uv)),
                   "stba_problem_add_reprojection")
        self._pending = []

    def NumResidualBlocks(self):
        self._flush()
        n = C.c_int64(0)
        capi.check(self._L.stba_problem_num_residual_blocks(self._h, C.byref(n)), "stba_problem_num_residual_blocks")
        return n.value

    def NumParameterBlocks(self):
        self._flush()
        n = C.c_int64(0)
        capi.check(self._L.stba_problem_num_parameter_blocks(self._h, C.byref(n)), "stba_problem_num_parameter_blocks")
        return n.value


class SolverOptions(capi.Options):
    """`ceres::Solver::Options`; `callbacks` holds callables taking an iteration-summary dict and
    returning SOLVER_CONTINUE / SOLVER_ABORT / SOLVER_TERMINATE_SUCCESSFULLY (test_ceres.h:83-96)."""

    def __init__(self, **kw):
        super().__init__(**kw)
        self.callbacks = []


def Solve(options, problem, summary=None):
    """`ceres::Solve(options, &problem, &summary)`; returns the summary."""
    problem._flush()
    cbs = list(getattr(options, "callbacks", []) or [])

    def cb(it):
        for f in cbs:
            r = f(it)
            if r not in (None, True, SOLVER_CONTINUE):
                return SOLVER_ABORT if r is False else r
        return SOLVER_CONTINUE

    return run_solve(problem._L.stba_problem_solve, problem._h, options, cb if cbs else None)
