"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol that
include/stba.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "stba.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(stba_[a-z0-9_]+)\s*\(", src)) - {"stba_iteration_callback"})


def test_library_exports_every_declared_symbol(stba):
    L = stba.capi.lib()
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), "libstba.so does not export %s" % n
    assert set(names) == set(stba.capi.SIGNATURES), "capi.SIGNATURES out of sync with include/stba.h"


def test_options_defaults_are_ceres_defaults(stba):
    o = stba.capi.Options()
    assert (o.max_num_iterations, o.max_num_consecutive_invalid_steps, o.jacobi_scaling) == (50, 5, 1)
    assert o.initial_trust_region_radius == 1e4 and o.max_trust_region_radius == 1e16
    assert o.min_relative_decrease == 1e-3 and o.min_lm_diagonal == 1e-6 and o.max_lm_diagonal == 1e32
    assert (o.function_tolerance, o.gradient_tolerance, o.parameter_tolerance) == (1e-6, 1e-10, 1e-8)
    assert o.linear_solver_type == stba.capi.SPARSE_SCHUR and o.num_threads == 1


def test_struct_sizes_match_the_header(stba):
    assert ctypes.sizeof(stba.capi.Options) == 8 * 4 + 9 * 8
    assert ctypes.sizeof(stba.capi.Iteration) == 4 * 4 + 8 * 8
    assert ctypes.sizeof(stba.capi.SummaryStruct) == 4 * 4 + 8 * 8 + 8 + 192 + 8 + 8


def test_no_cpu_fallback(stba):
    if stba.capi.device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(stba.capi.StbaError) as e:
        stba.engine.BAEngine(np.array([[0, 0, 0, 1.0]]), np.zeros((1, 3)), np.array([[0, 0, 1.0]]), [0], [0], np.zeros((1, 2)))
    assert e.value.status == stba.capi.ERR_NO_DEVICE


def test_invalid_arguments_are_rejected_before_touching_the_device(stba):
    with pytest.raises(stba.capi.StbaError) as e:    # not landmark-major
        stba.engine.BAEngine(np.array([[0, 0, 0, 1.0]]), np.zeros((1, 3)), np.ones((2, 3)), [0, 0], [1, 0], np.zeros((2, 2)))
    assert e.value.status == 1
    with pytest.raises(stba.capi.StbaError) as e:    # camera index out of range
        stba.engine.BAEngine(np.array([[0, 0, 0, 1.0]]), np.zeros((1, 3)), np.ones((1, 3)), [3], [0], np.zeros((1, 2)))
    assert e.value.status == 1


def test_problem_front_door_bookkeeping(stba):
    ceres = stba.ceres
    p = ceres.Problem()
    so3 = np.array([0, 0, 0, 1.0]); pos = np.zeros(3); lms = [np.array([0.1 * i, 0, 4.0]) for i in range(3)]
    for i, l in enumerate(lms):
        p.AddResidualBlock(ceres.ProjectFactor.Create([0.0, 0.0]), None, [so3, pos, l])
        p.AddParameterBlock(so3, 4, ceres.LieLocalParameterization())   # idempotent, test_ceres.h:124
    assert p.NumResidualBlocks() == 3 and p.NumParameterBlocks() == 5
    p.SetParameterBlockConstant(so3)
    with pytest.raises(stba.capi.StbaError):
        p.SetParameterBlockConstant(np.zeros(3))     # unknown block
    with pytest.raises(ValueError):
        p.AddResidualBlock(ceres.ProjectFactor.Create([0, 0]), None, [pos, so3, lms[0]])
    with pytest.raises(TypeError):
        p.AddResidualBlock(ceres.ProjectFactor.Create([0, 0]), None, [[0, 0, 0, 1.0], pos, lms[0]])


def test_no_product_module_imports_the_oracle():
    pkg = os.path.join(ROOT, "slam-tricks_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "ba_oracle" not in txt, f
