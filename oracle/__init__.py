"""CPU oracle for the bundle-adjustment hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in NumPy fp64 (and a plain-C twin for speed), the
algorithm the reference runs for its bundle-adjustment path
(`st20-g2o/src/include/test_ceres.h:98-152` → `ceres::Solve`).  Nothing under
`oracle/` is product code: only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s CPU-baseline legs may import, link or execute it, and there only
as the checker / the reported CPU baseline.

PARITY UNPINNED.  The arithmetic of the hot path lives in Ceres Solver, which
the reference pulls in with an un-versioned `find_package(Ceres)`
(`st17-ceres/src/CMakeLists.txt:5`, `st20-g2o/src/CMakeLists.txt:9`) and which
is neither vendored under `/root/reference` nor installable here (no Eigen,
Sophus, Ceres, SuiteSparse; no network).  The reference has no tests and no
golden vectors for this path.  The oracle therefore restates Ceres' published
trust-region Levenberg-Marquardt algorithm (2.0/2.1 behaviour, bounded by the
reference's use of `ceres::LocalParameterization`) and is pinned only to what
the reference does record:
  * the PnP ground-truth / initial poses and convergence behaviour printed in
    `st17-ceres/img/release.png` (`st17-ceres/src/main.cpp:14-35`),
  * the closed form of the SO(3) plus-Jacobian (`st17-ceres/docs/notes.tex:131-144`,
    `test_ceres.h:32-38`),
  * the reference's own hand Gauss-Newton (`st17-ceres/src/include/solver.hpp:387-462`),
    whose analytic Jacobians agree with the exact ones when t = 0.
See `tests/test_oracle_*.py` for those checks.
"""
