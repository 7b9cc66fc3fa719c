"""slam-tricks_b200 — B200-native bundle-adjustment hot path.

The directory name carries a hyphen, so import it through the ``stba`` alias module at the
repo root (``import stba``) or ``importlib.import_module("slam-tricks_b200")``.

Sub-modules: ``capi`` (ctypes binding of libstba.so), ``engine`` (SoA engine handle),
``ceres`` (mirror of the reference's ceres:: API subset), ``synth`` (deterministic scenes),
``shard`` (landmark sharding for multi-GPU), ``front`` (visibility + batched triangulation), ``calib`` (Zhang calibration, mirror of ns_st3::CalibSolver), ``posegraph`` (SE(3) pose graph), ``g2o`` (g2o-shaped vertex / edge front door of test_g2o.h).  Nothing in this package imports ``oracle/``.
"""
from . import synth  # noqa: F401
from . import capi, engine, ceres, shard, front, calib, posegraph, g2o  # noqa: F401
