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


def pipeline_tracker_inputs(orc, ref, prev_rgb, rgb, depth):
    """what HRBFFusion::processFrame (HRBFFusion.cpp:1063-1100) is about to hand to the tracker for the frame (rgb, depth), given the
    oracle pipeline `ref` (oracle/orc_pipeline.py) as it stands after the previous frame"""
    fr = orc.preprocess(ref.pp, depth)
    src = ref.fill if not orc.denseEnough(ref.pred["vertex"]) else ref.pred
    return dict(first=ref.rgba(prev_rgb), rgba=ref.rgba(rgb), src=src, fr=fr)


def init_tracker(o, up, pose, d):
    """the seven init* calls of a frame on an oracle or CUDA RGBDOdometry; `up` uploads (identity for the oracle)"""
    o.initFirstRGB(up(d["first"]))
    o.initICPModel(up(d["src"]["vertex"]), up(d["src"]["normal"]), 20.0, pose)
    o.initRGBModel(up(d["src"]["image"]))
    o.initCurvatureModel(up(d["src"]["curvk1"]), up(d["src"]["curvk2"]), pose)
    o.initICP(up(d["fr"]["vertex_filtered"]), up(d["fr"]["normal"]), 20.0)
    o.initRGB(up(d["rgba"]))
    o.initCurvature(up(d["fr"]["curv1"]), up(d["fr"]["curv2"]))
    o.initICPweight(up(d["src"]["icpw"]))
    return o


def tracker_noise_floor(orc, W, H, cam, pose, d, kw, n=3, seed=0):
    """How far the ORACLE's own result moves when its inputs change by one unit in the last place: the vertex maps of the model and of
    the current frame are multiplied by (1 + e), e drawn from {-2^-23, 0, +2^-23} per component, n times; returns the largest pose
    change (max of angle [rad] and translation [m]).  Two correct fp32 implementations of the tracker cannot be expected to agree
    better than this: the loop contains hard roundings (projective association, photometric correspondences) that such an input
    change flips for a few pixels."""
    import copy
    o0 = init_tracker(orc.Odometry(W, H, cam[2], cam[3], cam[0], cam[1]), lambda a: a, pose, d)
    t0, R0, _ = o0.getIncrementalTransformation(pose[:3, 3], pose[:3, :3], **kw)
    rng = np.random.default_rng(seed)
    worst = 0.0
    for _ in range(n):
        dd = dict(d)
        dd["src"] = dict(d["src"]); dd["fr"] = dict(d["fr"])
        for grp, key in (("src", "vertex"), ("fr", "vertex_filtered")):
            v = dd[grp][key].copy()
            v[..., :3] *= (1.0 + np.float32(2.0 ** -23) * rng.integers(-1, 2, v[..., :3].shape)).astype(np.float32)
            dd[grp][key] = v
        o1 = init_tracker(orc.Odometry(W, H, cam[2], cam[3], cam[0], cam[1]), lambda a: a, pose, dd)
        t1, R1, _ = o1.getIncrementalTransformation(pose[:3, 3], pose[:3, :3], **kw)
        worst = max(worst, *pose_err(R0, t0, R1, t1))
    return worst
