import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def default_float_loops():
    """HRBF_LITERAL=1 (development): the oracle runs the shaders' literal float-counter window loops for the
    whole session -- to be used together with HRBF_B200_LIB=build/libhrbf_literal.so (scripts/run_literal_variant.sh)"""
    return 1 if os.environ.get("HRBF_LITERAL") == "1" else 0


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
