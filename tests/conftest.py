import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def stba():
    import stba as pkg
    return pkg


@pytest.fixture(scope="session")
def scene_small(stba):
    """20 cameras / 300 landmarks / 1200 observations (ragged degrees 2..5) — seconds for the NumPy oracle."""
    return stba.synth.make_scene(20, 300, 1200)


@pytest.fixture(scope="session")
def scene_B(stba):
    """BASELINE.json configs[1]: 50 cameras / 5k landmarks / 50k observations."""
    return stba.synth.make_scene(*stba.synth.CONFIGS["B"])


def scene_args(sc):
    return (sc.cam_q, sc.cam_t, sc.lm, sc.obs_cam, sc.obs_lm, sc.obs_uv, sc.cam_const)
