"""CPU tests (no GPU): the oracle against closed-form facts, its own invariants and the committed
golden vectors.  SURVEY.md section 8 rows 1-5."""
import os

import numpy as np
import pytest

from tests.util import pair, pose_err

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_ldlt_matches_numpy(orc):
    import ctypes as C
    rng = np.random.default_rng(1)
    for _ in range(20):
        M = rng.normal(size=(6, 6)); A = M @ M.T + 1e-3 * np.eye(6); b = rng.normal(size=6)
        x = np.zeros(6)
        orc.lib().orc_ldlt_solve6(A.ctypes.data_as(C.POINTER(C.c_double)), b.ctypes.data_as(C.POINTER(C.c_double)), x.ctypes.data_as(C.POINTER(C.c_double)))
        np.testing.assert_allclose(x, np.linalg.solve(A, b), rtol=1e-9, atol=1e-12)
    # singular (all-zero) system -> zero update, like Eigen's ldlt().solve()
    x = np.ones(6); Z = np.zeros((6, 6)); z = np.zeros(6)
    orc.lib().orc_ldlt_solve6(Z.ctypes.data_as(C.POINTER(C.c_double)), z.ctypes.data_as(C.POINTER(C.c_double)), x.ctypes.data_as(C.POINTER(C.c_double)))
    assert np.all(x == 0)


def test_rodrigues(orc):
    import ctypes as C
    w = np.array([0.1, -0.2, 0.3]); R = np.zeros(9)
    orc.lib().orc_rodrigues(w.ctypes.data_as(C.POINTER(C.c_double)), R.ctypes.data_as(C.POINTER(C.c_double)))
    R = R.reshape(3, 3)
    th = np.linalg.norm(w); k = w / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    np.testing.assert_allclose(R, np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K, atol=1e-14)


def test_copy_resize_nan_semantics(orc):
    m0, pose0, m1, pose1, cam = pair(64, 48)
    v, n = orc.copyMaps(m0["vertex"], m0["normal"])
    rows = 48
    invalid = (m0["vertex"][..., 2] == 0) | (m0["normal"][..., 3] <= 0)
    assert np.array_equal(np.isnan(v[:rows]), invalid)
    assert np.array_equal(np.isnan(n[3 * rows:]), invalid)
    v1 = orc.resizeMap(v, 0)
    # a level-1 pixel is NaN iff any of its four sources is
    inv1 = invalid.reshape(24, 2, 32, 2).any(axis=(1, 3))
    assert np.array_equal(np.isnan(v1[:24]), inv1)
    ok = ~inv1
    mean_z = m0["vertex"][..., 2].reshape(24, 2, 32, 2).mean(axis=(1, 3))
    np.testing.assert_allclose(v1[48:72][ok], mean_z[ok], rtol=1e-6)
    n1 = orc.resizeMap(n, 1)
    nn = np.sqrt(n1[:24] ** 2 + n1[24:48] ** 2 + n1[48:72] ** 2)
    np.testing.assert_allclose(nn[ok], 1.0, atol=1e-5)


def test_icp_step_zero_residual_at_truth(orc):
    """At the true relative pose on noise-free maps the point-to-plane residual vanishes."""
    m0, pose0, m1, pose1, cam = pair(160, 120, kind="plane")
    o = orc.Odometry(160, 120, cam[2], cam[3], cam[0], cam[1])
    o.initICPModel(m0["vertex"], m0["normal"], 20.0, pose0)
    o.initICP(m1["vertex"], m1["normal"])
    o.fillNeutralCurvature()
    Rprev, tprev = pose0[:3, :3], pose0[:3, 3]
    A, b, res, sums, _ = orc.icpStep(pose1[:3, :3], pose1[:3, 3], o.map("vmap_curr", 0), o.map("nmap_curr", 0), o.map("ck1_curr", 0),
                                    o.map("ck2_curr", 0), np.linalg.inv(Rprev), tprev, cam, o.map("vmap_g_prev", 0), o.map("nmap_g_prev", 0),
                                    o.map("ck1_g_prev", 0), o.map("ck2_g_prev", 0), o.map("icpWeight", 0), use_weight=0)
    assert res[1] > 0.5 * 160 * 120
    assert np.sqrt(res[0] / res[1]) < 1e-5          # rms point-to-plane distance
    assert np.allclose(A, A.T)
    assert np.all(np.linalg.eigvalsh(A.astype(np.float64)) > -1e-3)


@pytest.mark.parametrize("icpWeight,so3", [(100.0, False), (10.0, True)])
def test_tracking_recovers_motion(orc, icpWeight, so3):
    m0, pose0, m1, pose1, cam = pair(320, 240, kind="room")
    o = orc.Odometry(320, 240, cam[2], cam[3], cam[0], cam[1])
    o.initFirstRGB(m0["rgba"])
    o.initICPModel(m0["vertex"], m0["normal"], 20.0, pose0)
    o.initRGBModel(m0["rgba"])
    o.initCurvatureModel(m0["k1"], m0["k2"], pose0)
    o.initICP(m1["vertex"], m1["normal"])
    o.initRGB(m1["rgba"])
    o.initCurvature(m1["k1"], m1["k2"])
    o.initICPweight(m0["icpw"])
    t, R, st = o.getIncrementalTransformation(pose0[:3, 3], pose0[:3, :3], icpWeight=icpWeight, so3=so3)
    ang, dt = pose_err(R, t, pose1[:3, :3], pose1[:3, 3])
    ang0, dt0 = pose_err(pose0[:3, :3], pose0[:3, 3], pose1[:3, :3], pose1[:3, 3])
    assert st.icp_iterations_run == 19
    tol = 0.15 if icpWeight >= 100 else 0.5   # the photometric term is coarser than ICP on this texture
    assert dt < tol * dt0 and ang < tol * ang0, (ang, dt, ang0, dt0)


def test_golden_gputest_pair(orc):
    """The oracle's pose / A / b on the reference's GPUTest frame pair is the committed golden vector
    (the reference records none)."""
    path = os.path.join(GOLD, "gputest_pair.npz")
    if not os.path.exists(path):
        pytest.skip("golden fixture not generated yet")
    g = np.load(path)
    from tests.gputest_pair import run_oracle
    out = run_oracle(orc, g)
    for k in ("icp_only_trans", "icp_only_rot", "faithful_trans", "faithful_rot"):
        np.testing.assert_allclose(out[k], g[k], atol=2e-6, err_msg=k)
    np.testing.assert_allclose(out["A0"], g["A0"], rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(out["b0"], g["b0"], rtol=1e-5, atol=1e-3)
