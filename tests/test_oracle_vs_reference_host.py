"""Pins the pose update of the Gauss-Newton loop (OdometryProvider::rodrigues / computeUpdateSE3) to the REFERENCE's own
header, Core/src/Utils/OdometryProvider.h, compiled unmodified against a minimal Eigen stand-in (oracle/eigen_mini: Eigen is not
installed) into oracle/_ref/libref_host.so by oracle/build_ref_host.sh.  CPU-only; skipped where the library cannot be built.
The rest of row 4 (the loop of RGBDOdometry.cpp around it and Eigen's ldlt) stays restated: "parity unpinned"."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE = os.path.join(os.path.dirname(HERE), "oracle")
PATH = os.path.join(ORACLE, "_ref", "libref_host.so")


def _lib():
    subprocess.call(["bash", os.path.join(ORACLE, "build_ref_host.sh")], stdout=subprocess.DEVNULL)
    return C.CDLL(PATH) if os.path.exists(PATH) else None


REF = _lib()
pytestmark = pytest.mark.skipif(REF is None, reason="oracle/_ref/libref_host.so not built and /root/reference absent")
dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))


def test_rodrigues_matches_reference_header(orc):
    rng = np.random.default_rng(11)
    cases = [np.zeros(3), np.array([1e-17, 0, 0]), np.array([2.2e-16, 0, 0]), np.array([0, 0, np.pi]), np.array([1e-9, -1e-9, 1e-9])]
    cases += [rng.standard_normal(3) * s for s in (1e-6, 1e-3, 0.05, 1.0, 3.0) for _ in range(20)]
    for w in cases:
        w = np.ascontiguousarray(w, np.float64)
        a, b = np.zeros(9), np.zeros(9)
        orc.lib().orc_rodrigues(dp(w), dp(a))
        REF.refh_rodrigues(dp(w), dp(b))
        assert np.array_equal(a, b), w                      # same fp64 expressions in the same order: bit-identical


def test_compute_update_se3_matches_reference_header(orc):
    """a chain of 19 Gauss-Newton updates (the iteration count of one frame), accumulated in resultRt like the reference does"""
    rng = np.random.default_rng(12)
    Ra, Rb = np.eye(4).ravel().copy(), np.eye(4).ravel().copy()
    for it in range(19):
        xi = np.ascontiguousarray(rng.standard_normal(6) * (1e-2 if it < 5 else 1e-4), np.float64)
        ia, ib = np.zeros(16, np.float32), np.zeros(16, np.float32)
        orc.lib().orc_computeUpdateSE3(dp(Ra), dp(xi), fp(ia))
        REF.refh_computeUpdateSE3(dp(Rb), dp(xi), fp(ib))
        assert np.array_equal(Ra, Rb) and np.array_equal(ia, ib), it
    R = Ra.reshape(4, 4)[:3, :3]
    np.testing.assert_allclose(R @ R.T, np.eye(3), atol=1e-14)
    assert np.array_equal(ia.reshape(4, 4)[3], [0, 0, 0, 1])


# ---- .klg wire format against the reference's own reader ---------------------------------------------------------
KLG_PATH = os.path.join(ORACLE, "_ref", "libref_klg.so")
KLG_W, KLG_H = 80, 60        # ONE size per process: the reference's Resolution singleton keeps the first size it is given


@pytest.mark.skipif(not os.path.exists(KLG_PATH), reason="oracle/_ref/libref_klg.so not built and /root/reference absent")
@pytest.mark.parametrize("compress", [True, False])
def test_klg_written_here_is_read_by_the_reference_reader(tmp_path, compress):
    """hrbffusion3d_b200.klg.write_klg -> GUI/src/Tools/RawLogReader.cpp (compiled unmodified): frame count, timestamps, zlib and raw depth,
    raw RGB, and flipColors; and our KlgReader returns what the reference reader returns."""
    from hrbffusion3d_b200 import klg
    L = C.CDLL(KLG_PATH)
    rng = np.random.default_rng(21)
    n = 5
    frames = [(33333 * i + 7, rng.integers(0, 9000, (KLG_H, KLG_W)).astype(np.uint16), rng.integers(0, 256, (KLG_H, KLG_W, 3)).astype(np.uint8)) for i in range(n)]
    frames[2] = (frames[2][0], np.zeros((KLG_H, KLG_W), np.uint16), frames[2][2])      # an all-zero depth image compresses to a few bytes
    path = tmp_path / "log.klg"
    path.write_bytes(klg.write_klg(frames, KLG_W, KLG_H, compress_depth=compress))
    assert L.refk_num_frames(str(path).encode(), KLG_W, KLG_H) == n
    for flip in (0, 1):
        ts = np.zeros(n + 2, np.int64)
        depth = np.zeros((n + 2, KLG_H, KLG_W), np.uint16)
        rgb = np.zeros((n + 2, KLG_H, KLG_W, 3), np.uint8)
        got = L.refk_read_klg(str(path).encode(), KLG_W, KLG_H, flip, n + 2, ts.ctypes.data_as(C.POINTER(C.c_longlong)),
                              depth.ctypes.data_as(C.POINTER(C.c_ushort)), rgb.ctypes.data_as(C.POINTER(C.c_ubyte)))
        assert got == n
        ours = list(klg.KlgReader(str(path), KLG_W, KLG_H, flip_colors=bool(flip)))
        assert len(ours) == n
        for i, (t, d, c) in enumerate(frames):
            assert ts[i] == t and np.array_equal(depth[i], d)
            assert np.array_equal(rgb[i], c[..., ::-1] if flip else c)
            assert ours[i][0] == ts[i] and np.array_equal(ours[i][1], depth[i]) and np.array_equal(ours[i][2], rgb[i])
