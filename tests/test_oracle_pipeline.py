"""CPU: the oracle's own processFrame pipeline (oracle/orc_pipeline.py) behaves like a tracker -- it follows a synthetic camera
to millimetres and grows a map -- and its pieces honour the reference's stated rules.  Independent of the CUDA path."""
import numpy as np
import pytest

from hrbffusion3d_b200 import synth
from tests.util import pose_err


@pytest.fixture(scope="module")
def seq():
    W, H = 320, 240
    cam = synth.default_camera(W, H)
    sc = synth.Scene("room")
    poses = synth.circle_trajectory(4, frames_per_rev=120)
    return W, H, cam, poses, [synth.render_depth(sc, p, W, H, cam, noise=True, seed=i) for i, p in enumerate(poses)]


@pytest.mark.parametrize("kw", [dict(icpWeight=100.0, so3=False), dict()])
def test_oracle_pipeline_tracks_ground_truth(orc, seq, kw):
    from oracle import orc_pipeline as op
    W, H, cam, poses, frames = seq
    f = op.HRBFFusion(W, H, cam, **kw)
    P0inv = np.linalg.inv(poses[0].astype(np.float64))
    counts = []
    for i, (depth, rgb) in enumerate(frames):
        T = f.processFrame(rgb, depth)
        gt = P0inv @ poses[i].astype(np.float64)
        ang, dt = pose_err(T[:3, :3], T[:3, 3], gt[:3, :3], gt[:3, 3])
        tol = 5e-3 if kw else 2e-2                                   # Kinect-noise depth: millimetres with ICP only; the photometric term on the
        assert ang < tol and dt < tol, (kw, i, ang, dt)              # procedural texture at quarter resolution pulls the pose by up to a centimetre
        counts.append(f.surfels.shape[0])
    assert counts[0] > 0.25 * W * H                                  # first frame initialises from (nearly) every valid pixel
    assert counts[-1] >= counts[0]                                   # fuse/clean keep the map and add the newly seen border
    # surfels younger than preictionConfThreshold (3) are not predicted from: after 4 frames the HRBF prediction is still (nearly)
    # empty and the tracker runs on the fill-in maps (HRBFFusion.cpp:1069-1086)
    assert (f.pred["vertex"][..., 2] > 0).mean() < 0.75


def test_oracle_preprocess_rules(orc, seq):
    """depth cut-offs, NaN-free outputs, unit normals, curvature default 1000 where undefined (depth_*.frag)"""
    W, H, cam, poses, frames = seq
    depth = frames[0][0].copy()
    depth[:4, :] = 0                       # no measurement
    depth[-4:, :] = 40000                  # 8 m: beyond globalDepthCutoff 3.5 m
    pp = orc.prep_params(cam, W, H)
    out = orc.preprocess(pp, depth)
    assert np.all(out["metric"][:4] == 0) and np.all(out["metric"][-4:] == 0)
    assert np.all(out["vertex_filtered"][:3, :, 2] == 0)
    n = out["normal"][..., :3]
    ln = np.linalg.norm(n, axis=-1)
    valid = ln > 0
    assert valid.mean() > 0.5
    np.testing.assert_allclose(ln[valid], 1.0, atol=1e-5)
    assert np.all(n[valid][:, 2] >= 0) or np.mean(n[valid][:, 2] >= 0) > 0.99        # PCA normals are flipped to n.z >= 0 (geometry.glsl:241-242)
    k1 = out["curv1"][..., 3]
    assert np.all(k1[~valid] == 1000.0)
    # (the principal DIRECTIONS may be NaN where the Weingarten system degenerates -- planar / umbilic points -- exactly as in the shader;
    # the curvature values, vertices, normals and depths never are)
    for k in ("filtered", "metric", "metric_filtered", "vertex_raw", "vertex_filtered", "normal", "gradient_mag"):
        assert not np.isnan(np.asarray(out[k], np.float64)).any(), k
    assert not np.isnan(out["curv1"][..., 3]).any() and not np.isnan(out["curv2"][..., 3]).any()
