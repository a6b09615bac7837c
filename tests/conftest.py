import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def default_float_loops():
    """The oracle runs the shaders' literal float-counter window loops (geometry.glsl:198-212, depth_curvature_gradient.frag:62-75),
    like the CUDA kernels; 0 (round 1's intended integer windows) is only set locally by the test that measures the difference."""
    return 1


@pytest.fixture(scope="session")
def orc():
    from oracle import orc_py
    orc_py.lib().orc_set_float_loops(default_float_loops())
    return orc_py


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from hrbffusion3d_b200 import lib
    lib()   # raises loudly if the CUDA library is missing: no fallback
    return torch
