"""Pins row 5 of the CPU oracle (map copies, pyramids, rigid transforms, image pyramids: SURVEY 8a row 5) to the REFERENCE's own
kernels: Core/src/Cuda/cudafuncs.cu compiled unmodified into oracle/_ref/libref_cudafuncs.so (oracle/build_ref.sh).

  * CPU: oracle vs the golden vectors the reference kernels produced on a B200 (tests/golden/ref_cudafuncs.npz, written by
    oracle/gen_ref5_golden.py; generated in round 2).  Every output agrees under tests/ref5_cases.compare; the one stated
    difference is the u8 truncation of pyrDownUcharGauss under the reference build's approximate division (see compare()).
  * GPU: oracle vs the reference kernels live, same cases."""
import os

import numpy as np
import pytest

from tests import ref5_cases

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_cudafuncs.npz")


def _stored(g, W, H, name, a):
    """the part of `a` the golden file holds for this case"""
    if (W, H) == ref5_cases.SIZES[0]:
        return a
    shape = g[f"{W}x{H}/{name}"].shape
    planes = shape[0]
    rows = a.shape[0] // planes
    return np.ascontiguousarray(a.reshape((planes, rows) + a.shape[1:])[:, ::5, ::5])


@pytest.mark.skipif(not os.path.exists(GOLD), reason="tests/golden/ref_cudafuncs.npz not generated yet (needs a GPU run of oracle/gen_ref5_golden.py)")
@pytest.mark.parametrize("W,H", ref5_cases.SIZES)
def test_oracle_row5_matches_reference_golden(orc, W, H):
    g = np.load(GOLD)
    out = ref5_cases.run_all(orc, orc, W, H)
    for name, a in out.items():
        ref5_cases.compare(name, _stored(g, W, H, name, a), g[f"{W}x{H}/{name}"])
        if f"{W}x{H}/{name}/nan" in g.files and a.dtype.kind == "f":
            assert int(np.isnan(a).sum()) == int(g[f"{W}x{H}/{name}/nan"][0]), name


@pytest.mark.gpu
@pytest.mark.skipif(not (os.path.exists(GOLD) or os.environ.get("HRBF_REF5_LIVE") == "1"),
                    reason="enabled once tests/golden/ref_cudafuncs.npz has been generated and checked (or HRBF_REF5_LIVE=1)")
@pytest.mark.parametrize("W,H", ref5_cases.SIZES)
def test_oracle_row5_matches_reference_kernels_live(orc, cuda, W, H):
    from oracle import ref5_py
    if not ref5_py.available():
        pytest.skip("oracle/_ref/libref_cudafuncs.so not built")
    ref = ref5_cases.run_all(ref5_py, orc, W, H)
    out = ref5_cases.run_all(orc, orc, W, H)
    for name, a in out.items():
        ref5_cases.compare(name, a, ref[name])
