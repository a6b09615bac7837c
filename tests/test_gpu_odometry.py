"""GPU parity tests (run on the B200): CUDA path through the C ABI vs the CPU oracle.
SURVEY.md section 8 rows 1-5 (+ RGB / SO3)."""
import os

import numpy as np
import pytest

from tests.util import pair, nan_eq_planes, pose_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")

# tolerances (north_star: pose <= 1e-5; the sums are fp32 products accumulated fp32-per-block /
# fp64-across-blocks on the GPU and fp64 in the oracle)
SUM_RTOL = 2e-5
POSE_TOL = 1e-5


def dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def build_both(orc, torch, W, H, kind="room", seed=0):
    from hrbffusion3d_b200 import odometry as od
    m0, pose0, m1, pose1, cam = pair(W, H, kind=kind, seed=seed)
    oo = orc.Odometry(W, H, cam[2], cam[3], cam[0], cam[1])
    go = od.RGBDOdometry(W, H, cam[2], cam[3], cam[0], cam[1])
    for o, f in ((oo, lambda a: a), (go, lambda a: dev(torch, a))):
        o.initFirstRGB(f(m0["rgba"]))
        o.initICPModel(f(m0["vertex"]), f(m0["normal"]), 20.0, pose0)
        o.initRGBModel(f(m0["rgba"]))
        o.initCurvatureModel(f(m0["k1"]), f(m0["k2"]), pose0)
        o.initICP(f(m1["vertex"]), f(m1["normal"]), 20.0)
        o.initRGB(f(m1["rgba"]))
        o.initCurvature(f(m1["k1"]), f(m1["k2"]))
        o.initICPweight(f(m0["icpw"]))
    return oo, go, (m0, pose0, m1, pose1, cam)


@pytest.mark.parametrize("W,H", [(160, 120), (640, 480)])
def test_pyramid_maps_match_oracle(orc, cuda, W, H):
    oo, go, _ = build_both(orc, cuda, W, H)
    for lvl in range(3):
        rows = H >> lvl
        for name in ("vmap_g_prev", "nmap_g_prev", "vmap_curr", "nmap_curr"):
            nan_eq_planes(go.map(name, lvl).cpu().numpy(), oo.map(name, lvl), rows, atol=2e-6, rtol=2e-6)
        for name in ("ck1_g_prev", "ck2_g_prev", "ck1_curr", "ck2_curr"):
            a, b = go.map(name, lvl).cpu().numpy(), oo.map(name, lvl)
            # curvature validity lives in plane w
            assert np.array_equal(np.isnan(a[3 * rows:]), np.isnan(b[3 * rows:]))
            ok = ~np.isnan(b[3 * rows:])
            for p in range(4):
                np.testing.assert_allclose(a[p * rows:(p + 1) * rows][ok], b[p * rows:(p + 1) * rows][ok], rtol=2e-6, atol=2e-6)
        a, b = go.map("icpWeight", lvl).cpu().numpy(), oo.map("icpWeight", lvl)
        assert np.array_equal(np.isnan(a), np.isnan(b))
        np.testing.assert_allclose(a[~np.isnan(b)], b[~np.isnan(b)], rtol=2e-6)
        # RGB branch prep is integer / exact
        for which in (0, 1, 2):
            assert np.array_equal(go.image(which, lvl).cpu().numpy(), oo.image(which, lvl)), (which, lvl)
        for which in (0, 1):
            a, b = go.depth(which, lvl).cpu().numpy(), oo.depth(which, lvl)
            assert np.array_equal(np.isnan(a), np.isnan(b))
            np.testing.assert_allclose(a[~np.isnan(b)], b[~np.isnan(b)], rtol=1e-6)


def test_row5_single_kernels_match_oracle(orc, cuda):
    """The one-to-one cudafuncs.cuh replacements (pitched destination, like DeviceArray2D)."""
    import ctypes as C
    from hrbffusion3d_b200._lib import lib, check, ptr, stream_ptr
    torch = cuda
    m0, pose0, m1, pose1, cam = pair(160, 120)
    rows, cols, pitch = 120, 160, 192          # pitch in elements (> cols)
    v = torch.full((4 * rows, pitch), 7.0, device="cuda"); n = torch.full((4 * rows, pitch), 7.0, device="cuda")
    d_v, d_n, d_k1, d_w = dev(torch, m0["vertex"]), dev(torch, m0["normal"]), dev(torch, m0["k1"]), dev(torch, m0["icpw"])   # keep alive
    check(lib().hrbf_copy_maps(ptr(d_v), ptr(d_n), ptr(v), C.c_size_t(pitch * 4), ptr(n), C.c_size_t(pitch * 4), rows, cols, stream_ptr()))
    ov, on = orc.copyMaps(m0["vertex"], m0["normal"])
    np.testing.assert_array_equal(v[:, :cols].cpu().numpy(), ov)
    np.testing.assert_array_equal(n[:, :cols].cpu().numpy(), on)
    assert float(v[:, cols:].min()) == 7.0      # padding untouched
    # resize (vmap / nmap) with stale-plane semantics: only plane x is written for NaN pixels
    for mode, fn in ((0, lib().hrbf_resize_vmap), (1, lib().hrbf_resize_nmap)):
        src = v if mode == 0 else n
        dst = torch.full((4 * (rows // 2), 96), 3.0, device="cuda")
        check(fn(ptr(src), C.c_size_t(pitch * 4), ptr(dst), C.c_size_t(96 * 4), rows, cols, stream_ptr()))
        ref = orc.resizeMap(ov if mode == 0 else on, mode, init=np.full((4 * (rows // 2), cols // 2), 3.0, np.float32))
        np.testing.assert_allclose(dst[:, :cols // 2].cpu().numpy(), ref, rtol=1e-6, atol=1e-7, equal_nan=True)
    # curvature copy / resize, weight copy / resize
    c = torch.zeros((4 * rows, pitch), device="cuda")
    check(lib().hrbf_copy_curvature_map(ptr(d_k1), ptr(c), C.c_size_t(pitch * 4), rows, cols, C.c_float(300.0), stream_ptr()))
    oc = orc.copyCurvatureMap(m0["k1"], 300.0)
    np.testing.assert_array_equal(c[:, :cols].cpu().numpy(), oc)
    c1 = torch.full((4 * (rows // 2), cols // 2), np.nan, device="cuda")
    check(lib().hrbf_resize_cmap(ptr(c), C.c_size_t(pitch * 4), ptr(c1), C.c_size_t(cols // 2 * 4), rows, cols, stream_ptr()))
    np.testing.assert_allclose(c1.cpu().numpy(), orc.resizeCMap(oc), rtol=1e-6, equal_nan=True)
    w = torch.zeros((rows, pitch), device="cuda")
    check(lib().hrbf_copy_icpweight_map(ptr(d_w), ptr(w), C.c_size_t(pitch * 4), rows, cols, stream_ptr()))
    ow = orc.copyicpWeightMap(m0["icpw"])
    np.testing.assert_array_equal(w[:, :cols].cpu().numpy(), ow)
    w1 = torch.zeros((rows // 2, cols // 2), device="cuda")
    check(lib().hrbf_resize_icpweight_map(ptr(w), C.c_size_t(pitch * 4), ptr(w1), C.c_size_t(cols // 2 * 4), rows, cols, stream_ptr()))
    np.testing.assert_allclose(w1.cpu().numpy(), orc.resizeicpWeightMap(ow), rtol=1e-6, equal_nan=True)
    # rigid transforms, in place
    R = np.ascontiguousarray(pose0[:3, :3]); t = np.ascontiguousarray(pose0[:3, 3])
    hp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    check(lib().hrbf_transform_maps(ptr(v), C.c_size_t(pitch * 4), ptr(n), C.c_size_t(pitch * 4), hp(R), hp(t), ptr(v), C.c_size_t(pitch * 4), ptr(n), C.c_size_t(pitch * 4), rows, cols, stream_ptr()))
    tv, tn = orc.tranformMaps(ov, on, R, t)
    nan_eq_planes(v[:, :cols].cpu().numpy(), tv, rows, atol=1e-6)
    nan_eq_planes(n[:, :cols].cpu().numpy(), tn, rows, atol=1e-6)


@pytest.mark.parametrize("W,H,use_weight", [(160, 120, 1), (640, 480, 1), (640, 480, 0), (1280, 960, 1)])
def test_icp_step_matches_oracle(orc, cuda, W, H, use_weight):
    from hrbffusion3d_b200 import odometry as od
    oo, go, (m0, pose0, m1, pose1, cam) = build_both(orc, cuda, W, H)
    Rp, tp = pose0[:3, :3], pose0[:3, 3]
    Rpi = np.linalg.inv(Rp).astype(np.float32)
    for lvl in range(3):
        camL = tuple(np.float32(c) / np.float32(1 << lvl) for c in cam)
        names = ("vmap_curr", "nmap_curr", "ck1_curr", "ck2_curr")
        gnames = ("vmap_g_prev", "nmap_g_prev", "ck1_g_prev", "ck2_g_prev", "icpWeight")
        A, b, res, sums, cg = od.icpStep(Rp, tp, *[go.map(k, lvl) for k in names], Rpi, tp, camL, *[go.map(k, lvl) for k in gnames],
                                         use_weight=bool(use_weight), want_corres=True)
        Ao, bo, reso, sumso, co = orc.icpStep(Rp, tp, *[oo.map(k, lvl) for k in names], Rpi, tp, camL, *[oo.map(k, lvl) for k in gnames],
                                              use_weight=use_weight, want_corres=True)
        # index work is bit-exact up to pixels whose projection sits on a rounding boundary
        mism = np.mean(np.any(cg.cpu().numpy() != co, axis=-1))
        assert mism < 2e-4, mism
        assert abs(res[1] - reso[1]) <= max(2.0, 2e-4 * reso[1])
        scale = np.abs(sumso[:27]).max()
        np.testing.assert_allclose(sums[:27], sumso[:27], rtol=SUM_RTOL * 50, atol=SUM_RTOL * scale)
        np.testing.assert_allclose(A, Ao, rtol=1e-3, atol=SUM_RTOL * np.abs(Ao).max())


@pytest.mark.parametrize("W,H", [(96, 72), (160, 120), (640, 480), (1280, 960)])
def test_icp_tile_reduction_matches_oracle(orc, cuda, W, H):
    """The TMA-staged tile form of the ICP reduction (csrc/icp_tile.cuh: bulk-copied current-frame rows, bounding box of the
    associations, bulk-copied model window, gather from shared memory, __ldg for associations outside the window) against the
    oracle's icpStep and against the per-pixel-gather kernel, on every pyramid level, for the pose tracking starts from (shift 0),
    the true pose (a few pixels), and a pose far outside the window (3 degrees / 5 cm: every association takes the fallback)."""
    from hrbffusion3d_b200 import synth
    oo, go, (m0, pose0, m1, pose1, cam) = build_both(orc, cuda, W, H)
    Rp, tp = pose0[:3, :3], pose0[:3, 3]
    Rpi = np.linalg.inv(Rp).astype(np.float32)
    far = (pose0.astype(np.float64) @ synth.make_pose(0.05, -0.03, 0.02, (0.05, -0.03, 0.02)).astype(np.float64)).astype(np.float32)
    names = ("vmap_curr", "nmap_curr", "ck1_curr", "ck2_curr")
    gnames = ("vmap_g_prev", "nmap_g_prev", "ck1_g_prev", "ck2_g_prev", "icpWeight")
    for pose in (pose0, pose1, far):
        Rc, tc = np.ascontiguousarray(pose[:3, :3]), np.ascontiguousarray(pose[:3, 3])
        for lvl in range(3):
            if (H >> lvl) < 16:
                continue
            camL = tuple(np.float32(c) / np.float32(1 << lvl) for c in cam)
            for use_weight in (1, 0):
                Ao, bo, reso, sumso, _ = orc.icpStep(Rc, tc, *[oo.map(k, lvl) for k in names], Rpi, tp, camL, *[oo.map(k, lvl) for k in gnames], use_weight=use_weight)
                At, bt, rest, sumst = go.icpStepLevel(lvl, Rc, tc, Rpi, tp, use_weight=bool(use_weight), tiled=True)
                Ag, bg, resg, sumsg = go.icpStepLevel(lvl, Rc, tc, Rpi, tp, use_weight=bool(use_weight), tiled=False)
                assert rest[1] == resg[1], (lvl, rest[1], resg[1])                           # same associations as the gather kernel, exactly
                assert abs(rest[1] - reso[1]) <= max(2.0, 2e-4 * reso[1])
                scale = np.abs(sumso[:27]).max()
                np.testing.assert_allclose(sumst[:27], sumsg[:27], rtol=1e-4, atol=1e-6 * scale)     # fp32 partial sums grouped differently
                np.testing.assert_allclose(sumst[:27], sumso[:27], rtol=SUM_RTOL * 50, atol=SUM_RTOL * scale)
    assert reso[1] > 0


def test_icp_step_search_window(orc, cuda):
    from hrbffusion3d_b200 import odometry as od
    oo, go, (m0, pose0, m1, pose1, cam) = build_both(orc, cuda, 160, 120)
    Rp, tp = pose0[:3, :3], pose0[:3, 3]
    Rpi = np.linalg.inv(Rp).astype(np.float32)
    names = ("vmap_curr", "nmap_curr", "ck1_curr", "ck2_curr")
    gnames = ("vmap_g_prev", "nmap_g_prev", "ck1_g_prev", "ck2_g_prev", "icpWeight")
    A, b, res, sums, _ = od.icpStep(Rp, tp, *[go.map(k, 0) for k in names], Rpi, tp, cam, *[go.map(k, 0) for k in gnames], use_search=True, search_radius=2)
    Ao, bo, reso, sumso, _ = orc.icpStep(Rp, tp, *[oo.map(k, 0) for k in names], Rpi, tp, cam, *[oo.map(k, 0) for k in gnames], use_search=1, radius=2)
    assert res[1] == reso[1]                      # the same best candidate in every 5 x 5 window
    np.testing.assert_allclose(sums[:27], sumso[:27], rtol=2e-4, atol=1e-6 * np.abs(sumso[:27]).max())      # measured: 1.3e-5 / 6e-8


def test_rgb_and_so3_steps_match_oracle(orc, cuda):
    from hrbffusion3d_b200 import odometry as od
    torch = cuda
    W, H = 320, 240
    oo, go, (m0, pose0, m1, pose1, cam) = build_both(orc, cuda, W, H)
    for lvl in range(3):
        camL = tuple(np.float32(c) / np.float32(1 << lvl) for c in cam)
        nextI, lastI = oo.image(1, lvl), oo.image(0, lvl)
        dx, dy = orc.sobel(nextI)
        Kl = np.array([[camL[0], 0, camL[2]], [0, camL[1], camL[3]], [0, 0, 1]], np.float64)
        Rrel = np.eye(3)
        krkinv = (Kl @ Rrel @ np.linalg.inv(Kl)).astype(np.float32)
        kt = (Kl @ np.array([0.002, -0.001, 0.003])).astype(np.float32)
        minScale = [5, 3, 1][lvl] ** 2 / 0.125 ** 2
        co, sigo, cnto = orc.computeRgbResidual(minScale, dx, dy, oo.depth(0, lvl), oo.depth(1, lvl), lastI, nextI, 0.07, kt, krkinv)
        cg, sigg, cntg = od.computeRgbResidual(minScale, dev(torch, dx), dev(torch, dy), go.depth(0, lvl), go.depth(1, lvl), go.image(0, lvl), go.image(1, lvl), 0.07, kt, krkinv)
        assert cnto > 100
        assert (sigg, cntg) == (sigo, cnto)                      # integer sums: bit-exact
        assert np.array_equal(cg.cpu().numpy().view(orc.DATATERM).reshape(co.shape), co)
        cloud = orc.projectToPointCloud(oo.depth(0, lvl), camL)
        sigma = float(np.sqrt(cnto))
        Ao, bo, so = orc.rgbStep(co, sigma, cloud, camL[0], camL[1], dx, dy, 0, 0.125)
        A, b, s = od.rgbStep(cg, sigma, dev(torch, cloud), camL[0], camL[1], dev(torch, dx), dev(torch, dy), 0, 0.125)
        np.testing.assert_allclose(s[:27], so[:27], rtol=1e-4, atol=2e-5 * np.abs(so[:27]).max())
    # SO3 (level 2 images)
    lvl = 2
    camL = tuple(np.float32(c) / np.float32(1 << lvl) for c in cam)
    Kl = np.array([[camL[0], 0, camL[2]], [0, camL[1], camL[3]], [0, 0, 1]], np.float64)
    from hrbffusion3d_b200 import synth
    Rr = synth.rot_xyz(0.002, -0.003, 0.001)
    B = (Kl @ Rr @ np.linalg.inv(Kl)).astype(np.float32); kinv = np.linalg.inv(Kl).astype(np.float32); krlr = (Kl @ Rr).astype(np.float32)
    Ao, bo, ro, so = orc.so3Step(oo.image(2, lvl), oo.image(1, lvl), B, kinv, krlr)
    A, b, r, s = od.so3Step(go.image(2, lvl), go.image(1, lvl), B, kinv, krlr)
    assert r[1] == ro[1]
    np.testing.assert_allclose(s[:10], so[:10], rtol=1e-4, atol=2e-5 * np.abs(so[:10]).max())


@pytest.mark.parametrize("W,H,kw", [
    (640, 480, dict(icpWeight=100.0, so3=False)),                 # ICP only
    (640, 480, dict(icpWeight=10.0, so3=True)),                   # reference default: joint RGB-D + SO3 pre-alignment
    (640, 480, dict(icpWeight=10.0, so3=False, pyramid=False)),
    (320, 240, dict(rgbOnly=True, so3=False, pyramid=False, fastOdom=True)),
    (320, 240, dict(rgbOnly=True, so3=False, pyramid=True, fastOdom=True)),      # early break at a coarse level (RGBDOdometry.cpp:1020-1023): the next level must warp with its own K
    (640, 480, dict(icpWeight=100.0, so3=False, fastOdom=True, if_curvature_info=False)),
    (1280, 960, dict(icpWeight=100.0, so3=False)),
])
def test_tracking_pose_matches_oracle(orc, cuda, W, H, kw):
    oo, go, (m0, pose0, m1, pose1, cam) = build_both(orc, cuda, W, H)
    to, Ro, sto = oo.getIncrementalTransformation(pose0[:3, 3], pose0[:3, :3], **kw)
    tg, Rg, stg = go.getIncrementalTransformation(pose0[:3, 3], pose0[:3, :3], **kw)
    ang, dt = pose_err(Ro, to, Rg, tg)
    # rgbOnly (sigma = -1, unit weights, hard integer correspondences) is not a contraction on this
    # data: round-off differences between two correct implementations grow ~10x per iteration.
    # It is checked per step (test_rgb_and_so3_steps_match_oracle) and over a 3-iteration run here.
    tol = 2e-4 if kw.get("rgbOnly") else POSE_TOL
    assert dt <= tol and ang <= tol, (ang, dt)
    np.testing.assert_allclose(np.asarray(Rg), np.asarray(Ro), atol=tol)
    assert stg.icp_iterations_run == sto.icp_iterations_run
    if not kw.get("rgbOnly"):
        assert abs(stg.lastICPCount - sto.lastICPCount) <= max(3.0, 3e-4 * sto.lastICPCount)
    # and both actually track: closer to the true pose than the start
    ang1, dt1 = pose_err(Rg, tg, pose1[:3, :3], pose1[:3, 3])
    ang0, dt0 = pose_err(pose0[:3, :3], pose0[:3, 3], pose1[:3, :3], pose1[:3, 3])
    if kw.get("icpWeight", 0) >= 100:      # the photometric term is not guaranteed to help on this texture
        assert dt1 < dt0 and ang1 < ang0


def test_persistent_and_graph_trackers_agree(orc, cuda):
    """The persistent cooperative kernel (default) and the first-generation kernel-per-reduction graph run the same
    arithmetic with a different cross-CTA summation grouping (fp64): poses agree to round-off."""
    for kw in (dict(icpWeight=100.0, so3=False), dict(icpWeight=10.0, so3=True), dict(icpWeight=100.0, so3=False, pyramid=False, fastOdom=True)):
        res = []
        for graph in (False, True):
            oo, go, (m0, pose0, m1, pose1, cam) = build_both(orc, cuda, 320, 240)
            go.setTracker(graph)
            t, R, st = go.getIncrementalTransformation(pose0[:3, 3], pose0[:3, :3], **kw)
            res.append((t, R, st))
        ang, dt = pose_err(res[0][1], res[0][0], res[1][1], res[1][0])
        tol = 1e-6 if kw["icpWeight"] >= 100 else 2e-4
        assert ang <= tol and dt <= tol, (kw, ang, dt)
        assert res[0][2].icp_iterations_run == res[1][2].icp_iterations_run
        assert res[0][2].kernel_launches == 1 and res[1][2].kernel_launches > 1


@pytest.mark.parametrize("W,H", [(96, 72), (320, 240), (640, 480), (1280, 960)])
def test_resident_tile_tracker_matches_streaming_tracker(orc, cuda, W, H):
    """The persistent tracker with its ICP tiles resident in shared memory (TMA-staged once per level, csrc/icp_tile.cuh) against the
    same kernel re-reading the maps every iteration, and against the oracle: same associations, fp32 partial sums grouped by tile
    instead of by pixel range.  1280x960: level 0 does not fit and streams, levels 1-2 are resident.  Also with the 256-thread
    shape (half the shared-memory budget) and with a start pose far from the solution (associations outside the staged window)."""
    from hrbffusion3d_b200 import synth
    floor = None
    for kw in (dict(icpWeight=100.0, so3=False), dict(icpWeight=10.0, so3=True)):
        for threads in (512, 384, 256):
            res = {}
            for resident in (True, False):
                oo, go, (m0, pose0, m1, pose1, cam) = build_both(orc, cuda, W, H)
                go.setTrackerTiles(resident)
                go.setTrackerThreads(threads)
                res[resident] = go.getIncrementalTransformation(pose0[:3, 3], pose0[:3, :3], **kw)
            to, Ro, sto = oo.getIncrementalTransformation(pose0[:3, 3], pose0[:3, :3], **kw)
            ang, dt = pose_err(res[True][1], res[True][0], res[False][1], res[False][0])
            # ICP only: the two forms differ in how fp32 partial sums are grouped.  With the photometric term the loop amplifies such
            # differences (hard roundings of ~5 000 correspondences): bounded by the oracle's own sensitivity to its inputs' last place
            if kw["icpWeight"] >= 100:
                tol = 1e-6 if W >= 640 else 1e-5
            else:
                if floor is None:
                    from tests.util import tracker_noise_floor
                    d = dict(first=m0["rgba"], rgba=m1["rgba"], src=dict(vertex=m0["vertex"], normal=m0["normal"], image=m0["rgba"], curvk1=m0["k1"], curvk2=m0["k2"], icpw=m0["icpw"]),
                             fr=dict(vertex_filtered=m1["vertex"], normal=m1["normal"], curv1=m1["k1"], curv2=m1["k2"]))
                    floor = tracker_noise_floor(orc, W, H, cam, pose0, d, dict(icpWeight=10.0, so3=True), n=3)
                tol = max(2e-5, 4 * floor)
            print(f"{W}x{H} {kw} threads {threads}: resident vs streaming ang {ang:.1e} t {dt:.1e} (tol {tol:.1e})")
            assert ang <= tol and dt <= tol, (kw, threads, ang, dt, tol)
            assert res[True][2].icp_iterations_run == res[False][2].icp_iterations_run == sto.icp_iterations_run
            assert abs(res[True][2].lastICPCount - res[False][2].lastICPCount) <= 2
            if W >= 640 and kw["icpWeight"] >= 100:
                ang, dt = pose_err(res[True][1], res[True][0], Ro, to)
                assert ang <= POSE_TOL and dt <= POSE_TOL, (kw, threads, ang, dt)
    # far start: 2 degrees / 3 cm off -> most associations leave the window staged for the level's first pose
    oo, go, (m0, pose0, m1, pose1, cam) = build_both(orc, cuda, W, H)
    far = (pose0.astype(np.float64) @ synth.make_pose(0.03, -0.02, 0.02, (0.03, -0.02, 0.01)).astype(np.float64)).astype(np.float32)
    out = {}
    for resident in (True, False):
        go.setTrackerTiles(resident)
        out[resident] = go.getIncrementalTransformation(far[:3, 3], far[:3, :3], icpWeight=100.0, so3=False)
    to, Ro, sto = oo.getIncrementalTransformation(far[:3, 3], far[:3, :3], icpWeight=100.0, so3=False)
    ang, dt = pose_err(out[True][1], out[True][0], out[False][1], out[False][0])
    print(f"{W}x{H} far start: resident vs streaming ang {ang:.1e} t {dt:.1e}; streaming vs oracle %.1e %.1e; inliers {out[True][2].lastICPCount:.0f} / {out[False][2].lastICPCount:.0f} / {sto.lastICPCount:.0f}"
          % pose_err(out[False][1], out[False][0], Ro, to))
    if W >= 320:      # (a 96x72 frame does not converge from this far: nothing to compare)
        # from this far the first iterations take half-blind steps: round-off decides which of ~1e5 borderline associations are in,
        # and the three implementations (resident, streaming, oracle) land equally far from one another
        angs, dts = pose_err(out[False][1], out[False][0], Ro, to)
        bound = max(2e-5, 3 * max(angs, dts))
        assert ang <= bound and dt <= bound, (ang, dt, bound)
        assert abs(out[True][2].lastICPCount - out[False][2].lastICPCount) <= max(3, 5e-3 * sto.lastICPCount)


def test_tracking_two_frames_so3_swap(orc, cuda):
    """Second call exercises the lastNextImage/nextImage swap (RGBDOdometry.cpp:1239-1245)."""
    oo, go, (m0, pose0, m1, pose1, cam) = build_both(orc, cuda, 320, 240)
    for _ in range(2):
        to, Ro, _s = oo.getIncrementalTransformation(pose0[:3, 3], pose0[:3, :3], icpWeight=10.0, so3=True)
        tg, Rg, _s = go.getIncrementalTransformation(pose0[:3, 3], pose0[:3, :3], icpWeight=10.0, so3=True)
        ang, dt = pose_err(Ro, to, Rg, tg)
        assert dt <= POSE_TOL and ang <= POSE_TOL, (ang, dt)
        for o, f in ((oo, lambda a: a), (go, lambda a: dev(cuda, a))):
            o.initRGB(f(m1["rgba"]))


def test_gputest_pair_golden(orc, cuda):
    """Reference GPUTest frame pair: CUDA path vs the committed golden vector (oracle output)."""
    g = np.load(os.path.join(GOLD, "gputest_pair.npz"))
    from tests.gputest_pair import run_cuda
    out = run_cuda(g)
    for k in ("icp_only", "faithful"):
        ang, dt = pose_err(out[k + "_rot"], out[k + "_trans"], g[k + "_rot"], g[k + "_trans"])
        assert dt <= POSE_TOL and ang <= POSE_TOL, (k, ang, dt)
    np.testing.assert_allclose(out["A0"], g["A0"], rtol=1e-3, atol=SUM_RTOL * np.abs(g["A0"]).max())
    np.testing.assert_allclose(out["b0"], g["b0"], rtol=1e-3, atol=SUM_RTOL * np.abs(g["A0"]).max())
    assert abs(out["res0"][1] - g["res0"][1]) <= 30


def test_invalid_arguments_fail_loudly(cuda):
    import ctypes as C
    from hrbffusion3d_b200._lib import lib, HrbfError
    from hrbffusion3d_b200 import odometry as od
    with pytest.raises(HrbfError):
        od.RGBDOdometry(641, 480, 320, 240, 528, 528)          # width not a multiple of 8
    h = C.c_void_p()
    assert lib().hrbf_odometry_create(None, 640, 480, C.c_float(1), C.c_float(1), C.c_float(1), C.c_float(1), C.c_float(.1), C.c_float(.3)) == -1
    assert b"invalid argument" in lib().hrbf_last_error()


def _pipeline_frame1_inputs(orc, W, H):
    """the tracker inputs of the SECOND frame of the oracle pipeline (the first tracked one): model maps = fill-in of frame 0"""
    from oracle import orc_pipeline as op
    from hrbffusion3d_b200 import synth
    from tests.util import pipeline_tracker_inputs
    cam = synth.default_camera(W, H)
    sc = synth.Scene("room")
    frames = [synth.render_depth(sc, p, W, H, cam, noise=True, seed=i) for i, p in enumerate(synth.circle_trajectory(2, frames_per_rev=120))]
    f = op.HRBFFusion(W, H, cam)
    f.processFrame(frames[0][1], frames[0][0])
    return cam, f.currPose.copy(), pipeline_tracker_inputs(orc, f, frames[0][1], frames[1][1], frames[1][0])


from tests.util import init_tracker as _init_tracker  # noqa: E402


def test_default_config_frame_agrees_iteration_by_iteration(orc, cuda):
    """The benchmarked configuration (reference defaults: RGB-D + ICP + SO3 pre-alignment, 640x480) pinned reduction by reduction.
    The Gauss-Newton loop of RGBDOdometry.cpp:796-1249 is driven from Python over the oracle's step functions
    (tests/gn_loop_py.py) and, at the SAME pose in every one of its 3 SO3 + 19 SE3 iterations, the CUDA step functions are
    evaluated beside them: the integer outputs of computeRgbResidual (sigma, count) and the SO3 counts must be identical, the ICP
    inlier count may move by a correspondence on the edge of a threshold, the 27 + 27 + 9 float sums agree to that one
    correspondence.  Then both trackers run free on these (bit-identical) inputs: what separates them is bounded by what a change of
    the inputs in the last place does to the ORACLE itself (tests/util.tracker_noise_floor) -- the hard roundings of the photometric
    correspondences make this configuration sensitive (~5e-6 per unit in the last place with ~4 800 correspondences at level 0)."""
    from hrbffusion3d_b200 import odometry as od
    from tests import gn_loop_py as gn
    W, H = 640, 480
    cam, pose, d = _pipeline_frame1_inputs(orc, W, H)
    mk_o = lambda: _init_tracker(orc.Odometry(W, H, cam[2], cam[3], cam[0], cam[1]), lambda a: a, pose, d)
    mk_g = lambda: _init_tracker(od.RGBDOdometry(W, H, cam[2], cam[3], cam[0], cam[1]), lambda a: dev(cuda, a), pose, d)
    t, R, log = gn.run(gn.OracleBackend(orc, mk_o()), cam, pose[:3, 3], pose[:3, :3], shadow=gn.CudaBackend(od, mk_g(), orc, cuda))
    to, Ro, sto = mk_o().getIncrementalTransformation(pose[:3, 3], pose[:3, :3])
    ang, dt = pose_err(R, t, Ro, to)
    assert ang <= 2e-6 and dt <= 2e-6, "the Python restatement of the loop left the oracle's C loop"      # fp64 solve: numpy vs the oracle's LDLT
    n_so3 = sum(r["kind"] == "so3" for r in log)
    assert n_so3 >= 2 and len(log) - n_so3 == 19
    rel = lambda a, b, n: float(np.abs(a[:n] - b[:n]).max() / np.abs(a[:n]).max())
    for r in log:
        where = (r["kind"], r["level"], r["it"])
        if r["kind"] == "so3":
            assert r["d_res"][1] == r["s_res"][1], where
            assert rel(r["d_sums"], r["s_sums"], 9) <= 1e-6, where
            continue
        assert (r["d_sigma"], r["d_count"]) == (r["s_sigma"], r["s_count"]), where               # integer sums: bit-exact
        assert abs(r["d_icp_res"][1] - r["s_icp_res"][1]) <= 2, where
        assert rel(r["d_icp_sums"], r["s_icp_sums"], 27) <= 1e-4, where                            # one of ~50 000 (level 1) correspondences
        assert rel(r["d_rgb_sums"], r["s_rgb_sums"], 27) <= 1e-5, where
    tg, Rg, stg = mk_g().getIncrementalTransformation(pose[:3, 3], pose[:3, :3])
    ang, dt = pose_err(Ro, to, Rg, tg)
    print(f"default configuration, free-running on identical inputs: CUDA vs oracle ang {ang:.2e} t {dt:.2e}")
    from tests.util import tracker_noise_floor
    floor = tracker_noise_floor(orc, W, H, cam, pose, d, {}, n=4)
    print(f"the oracle's own sensitivity to one unit in the last place of its inputs: {floor:.2e}")
    assert stg.lastSO3Count == sto.lastSO3Count and abs(stg.lastRGBCount - sto.lastRGBCount) <= 2
    assert ang <= max(POSE_TOL, 4 * floor) and dt <= max(POSE_TOL, 4 * floor), (ang, dt, floor)


def test_rgb_prep_single_functions_match_oracle(orc, cuda):
    """The GPUTest- and RGB-branch preparation functions of cudafuncs.cuh as one-to-one C-ABI calls on PITCHED arrays (pyrDown, createVMap,
    createNMap, verticesToDepth, pyrDownGaussF, pyrDownUcharGauss, imageBGRToIntensity, computeDerivativeImages, projectToPointCloud)
    against the oracle functions that tests/test_oracle_vs_reference_row5.py pins to the reference's own kernels."""
    import ctypes as C
    from hrbffusion3d_b200._lib import lib, check, ptr, stream_ptr, Camera
    torch = cuda
    W, H = 160, 120
    m0, pose0, m1, pose1, cam = pair(W, H)
    L = lib()
    camS = Camera(*cam)
    pitched = lambda rows, cols, dt, fill: torch.full((rows, cols + 24), fill, dtype=dt, device="cuda")      # 24 elements of padding per row
    step = lambda t: C.c_size_t(t.shape[1] * t.element_size())
    d_v = dev(torch, m0["vertex"])
    # verticesToDepth
    d = pitched(H, W, torch.float32, 7.0)
    check(L.hrbf_vertices_to_depth(ptr(d_v), ptr(d), step(d), H, W, C.c_float(2.0), stream_ptr()))
    ref_d = orc.verticesToDepth(m0["vertex"], 2.0)
    np.testing.assert_array_equal(d[:, :W].cpu().numpy(), ref_d)
    assert float(d[:, W:].min()) == 7.0
    ref_d = orc.verticesToDepth(m0["vertex"], 20.0)
    check(L.hrbf_vertices_to_depth(ptr(d_v), ptr(d), step(d), H, W, C.c_float(20.0), stream_ptr()))
    # pyrDownGaussF
    g = pitched(H // 2, W // 2, torch.float32, 0.0)
    check(L.hrbf_pyr_down_gauss_f(ptr(d), step(d), ptr(g), step(g), H, W, stream_ptr()))
    np.testing.assert_array_equal(g[:, :W // 2].cpu().numpy(), orc.pyrDownGaussF(ref_d))
    # pyrDown (raw depth in mm, 0 = invalid)
    raw = np.nan_to_num(ref_d * 1000.0).astype(np.float32)
    d_raw = pitched(H, W, torch.float32, 0.0); d_raw[:, :W] = dev(torch, raw)
    pd = pitched(H // 2, W // 2, torch.float32, 0.0)
    check(L.hrbf_pyr_down(ptr(d_raw), step(d_raw), ptr(pd), step(pd), H, W, stream_ptr()))
    np.testing.assert_allclose(pd[:, :W // 2].cpu().numpy(), orc.pyrDownDepth(raw), rtol=1e-6, atol=1e-4)
    # imageBGRToIntensity, pyrDownUcharGauss, computeDerivativeImages
    img = pitched(H, W, torch.uint8, 0)
    check(L.hrbf_image_bgr_to_intensity(ptr(dev(torch, m0["rgba"])), ptr(img), step(img), H, W, stream_ptr()))
    ref_img = orc.rgbaToIntensity(m0["rgba"])
    np.testing.assert_array_equal(img[:, :W].cpu().numpy(), ref_img)
    img2 = pitched(H // 2, W // 2, torch.uint8, 0)
    check(L.hrbf_pyr_down_uchar_gauss(ptr(img), step(img), ptr(img2), step(img2), H, W, stream_ptr()))
    np.testing.assert_array_equal(img2[:, :W // 2].cpu().numpy(), orc.pyrDownUcharGauss(ref_img))
    dx, dy = pitched(H, W, torch.int16, 0), pitched(H, W, torch.int16, 0)
    check(L.hrbf_compute_derivative_images(ptr(img), step(img), ptr(dx), step(dx), ptr(dy), step(dy), H, W, stream_ptr()))
    rdx, rdy = orc.sobel(ref_img)
    np.testing.assert_array_equal(dx[:, :W].cpu().numpy(), rdx)
    np.testing.assert_array_equal(dy[:, :W].cpu().numpy(), rdy)
    # projectToPointCloud at pyramid level 1 of a dense half-size depth
    g_dense = g[:, :W // 2].contiguous()
    cloud = torch.zeros((H // 2, W // 2, 3), device="cuda")
    check(L.hrbf_project_to_point_cloud(ptr(g_dense), C.c_size_t(W // 2 * 4), ptr(cloud), C.c_size_t(W // 2 * 12), camS, 1, H // 2, W // 2, stream_ptr()))
    camL = tuple(np.float32(c) / np.float32(2) for c in cam)
    np.testing.assert_allclose(cloud.cpu().numpy(), orc.projectToPointCloud(orc.pyrDownGaussF(ref_d), camL), rtol=1e-6, atol=1e-7, equal_nan=True)
    # createVMap / createNMap (TUM-style raw depth, 1/5000 m)
    raw5 = np.nan_to_num(ref_d * 5000.0).astype(np.float32)
    vm, nm = pitched(4 * H, W, torch.float32, 0.0), pitched(4 * H, W, torch.float32, 0.0)
    d5 = dev(torch, raw5)
    check(L.hrbf_create_vmap(camS, ptr(d5), C.c_size_t(W * 4), ptr(vm), step(vm), H, W, C.c_float(20.0), C.c_float(1.0 / 5000.0), stream_ptr()))
    check(L.hrbf_create_nmap(ptr(vm), step(vm), ptr(nm), step(nm), H, W, stream_ptr()))
    rv = orc.createVMap(cam, raw5, 20.0, 1.0 / 5000.0)
    nan_eq_planes(vm[:, :W].cpu().numpy(), rv, H, atol=1e-6)
    nan_eq_planes(nm[:, :W].cpu().numpy(), orc.createNMap(rv), H, atol=1e-5)
