"""Pins the oracle's restatement of the GLSL passes to the REFERENCE's own shader text: the shaders under
/root/reference/Core/src/Shaders are compiled for the CPU (oracle/build_ref_glsl.py: GLSL types / built-ins as C++ in
oracle/glsl_cpu.h, qualifiers stripped mechanically, nothing else touched) and run fragment by fragment on the same inputs as
the oracle.  CPU-only; skipped where the library cannot be built (no reference sources on the machine) and was not shipped.

Tolerance: both sides are fp32, but the oracle mirrors the operation order of the CUDA kernels' bit-exact parts, the shader
text has its own; discrete outcomes (which pixels find a surface, which neighbour is nearest) must agree, continuous ones to
a few ulps of the quantities involved."""
import numpy as np
import pytest

from hrbffusion3d_b200 import synth
from oracle import refglsl_py as rg

pytestmark = pytest.mark.skipif(not rg.available(), reason="oracle/_ref/libref_glsl.so not built and /root/reference absent")


def _scene(W, H, kind, stride):
    """a confident map seen from a nearby pose (the helper of tests/test_gpu_indexmap.py, without the GPU)"""
    from tests.test_gpu_indexmap import _scene as f
    return f(W, H, kind, stride)


@pytest.mark.parametrize("W,H,kind,stride,kw", [
    (160, 120, "room", 1, {}),
    (320, 240, "plane", 1, {}),
    (320, 240, "room", 2, {}),                                              # sparse map: min-neighbour rejections
    (320, 240, "room", 1, dict(win=2, minNeighbors=4, maxNeighbors=16, confThreshold=4.5)),
    (640, 480, "room", 1, {}),                                              # BASELINE resolution, reference defaults
])
def test_predict_hrbf_oracle_matches_reference_shader(orc, W, H, kind, stride, kw):
    """row 6: Shaders/predict_hrbf.frag + hrbfbase.glsl + color.glsl + utils.glsl"""
    s, pose, cam = _scene(W, H, kind, stride)
    idx = orc.predictIndices(pose, s, cam, W, H)
    a, b = orc.predictHRBF(idx, cam, W, H, **kw), rg.predictHRBF(idx, cam, W, H, **kw)
    fa, fb = a["vertex"][..., 2] > 0, b["vertex"][..., 2] > 0
    assert fa.sum() > 0.5 * W * H
    assert np.mean(fa != fb) <= 1e-5, "found-flag flips"
    both = fa & fb
    dv = (a["vertex"][..., :3].astype(np.float64) - b["vertex"][..., :3])[both]
    assert np.sqrt((dv ** 2).sum(-1).mean()) <= 2e-6 and np.abs(dv).max() <= 5e-5      # metres; the bisection ends on a 6-um interval
    assert np.mean(np.abs(dv).max(-1) == 0) >= 0.9                                        # most roots are bit-identical
    dn = np.abs((a["normal"][..., :3] - b["normal"][..., :3])[both]).max(-1)              # gradient at the root: sensitive at depth edges
    assert np.mean(dn > 1e-4) <= 1e-4 and dn.max() <= 5e-3
    for k in ("curvk1", "curvk2", "image", "time"):                                      # copied from the nearest neighbour: same pick
        assert np.array_equal(a[k][both], b[k][both]), k
    assert np.array_equal(a["vertex"][..., 3][both], b["vertex"][..., 3][both]) and np.array_equal(a["normal"][..., 3][both], b["normal"][..., 3][both])
    assert np.abs((a["icpw"] - b["icpw"])[both]).max() <= 1e-5
    none = ~fa & ~fb                                                                      # no surface: the cleared values + curvature w = 1000
    for k in ("vertex", "normal"):
        assert not a[k][none].any() and not b[k][none].any()
    assert np.array_equal(a["curvk1"][none], b["curvk1"][none]) and np.array_equal(a["icpw"][none], b["icpw"][none])
