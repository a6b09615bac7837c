"""Shared helpers for the parity tests."""
import numpy as np

from hrbffusion3d_b200 import synth

CAM640 = synth.default_camera(640, 480)


def pair(width=160, height=120, kind="room", seed=0, delta=None):
    """Two views of a scene: (maps0, pose0), (maps1, pose1), cam."""
    cam = synth.default_camera(width, height)
    sc = synth.Scene(kind)
    pose0 = synth.make_pose(0.02, -0.03, 0.01, (0.05, -0.02, 0.0))
    if delta is None:
        delta = synth.make_pose(0.004, -0.006, 0.003, (0.008, -0.005, 0.006))
    pose1 = (pose0.astype(np.float64) @ delta.astype(np.float64)).astype(np.float32)
    m0 = synth.ideal_maps(sc, pose0, width, height, cam, seed=seed)
    m1 = synth.ideal_maps(sc, pose1, width, height, cam, seed=seed + 1)
    return m0, pose0, m1, pose1, cam


def nan_eq_planes(a, b, rows, atol=1e-6, rtol=1e-6):
    """Compare two SoA [4*rows, cols] maps the way the reference consumes them: plane x decides
    validity; the other planes are only compared where plane x is valid."""
    a = np.asarray(a); b = np.asarray(b)
    ax, bx = a[:rows], b[:rows]
    assert np.array_equal(np.isnan(ax), np.isnan(bx)), "validity masks differ"
    ok = ~np.isnan(ax)
    for p in range(a.shape[0] // rows):
        pa, pb = a[p * rows:(p + 1) * rows][ok], b[p * rows:(p + 1) * rows][ok]
        np.testing.assert_allclose(pa, pb, rtol=rtol, atol=atol, err_msg=f"plane {p}")


def pose_err(R0, t0, R1, t1):
    dR = np.asarray(R0, np.float64).T @ np.asarray(R1, np.float64)
    # arccos((tr-1)/2) loses half the digits near 0; use the skew part instead
    ang = np.arcsin(min(1.0, np.linalg.norm(dR - dR.T) / (2.0 * np.sqrt(2.0))))
    return float(ang), float(np.linalg.norm(np.asarray(t0, np.float64) - np.asarray(t1, np.float64)))
