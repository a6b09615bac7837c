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
    """a confident map seen from a nearby pose (the same construction as tests/test_gpu_indexmap.py)"""
    from tests.util import pair
    m0, pose0, m1, pose1, cam = pair(W, H, kind=kind)
    return synth.surfels_from_maps(m0, pose0, stride=stride), pose1, cam


@pytest.mark.parametrize("W,H,kind,stride,kw", [
    (160, 120, "room", 1, {}),
    (320, 240, "plane", 1, {}),
    (320, 240, "room", 2, {}),                                              # sparse map: min-neighbour rejections
    (320, 240, "room", 1, dict(win=2, minNeighbors=4, maxNeighbors=16, confThreshold=4.5)),
    (640, 480, "room", 1, {}),                                              # BASELINE resolution, reference defaults
    (1280, 960, "room", 1, dict(maxNeighbors=16)),                          # BASELINE config 4: the float counters of the ring loops (:74-80) at this texture size
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


def _frame(W, H, kind, seed=3):
    cam = synth.default_camera(W, H)
    depth, rgb = synth.render_depth(synth.Scene(kind), synth.circle_trajectory(1, frames_per_rev=120)[0], W, H, cam, noise=True, seed=seed)
    return cam, depth, rgb


def _identical(x, y):
    return np.mean((x == y) | (np.isnan(x) & np.isnan(y)))


@pytest.fixture
def literal_windows(orc):
    """the oracle with the shaders' literal float-counter window loops (orc_set_float_loops, oracle/orc_prep.c)"""
    from tests.conftest import default_float_loops
    orc.lib().orc_set_float_loops(1)
    yield
    orc.lib().orc_set_float_loops(default_float_loops())


@pytest.mark.parametrize("W,H,kind", [(160, 120, "room"), (320, 240, "plane"), (640, 480, "room"), (1280, 960, "room")])
def test_preprocess_oracle_matches_reference_shaders(orc, literal_windows, W, H, kind):
    """row 10: depth_bilateral.frag, depth_metric_raw/filtered.frag, depth_vertex_normal_radius.frag (+ geometry.glsl PCA normals,
    surfels.glsl radius / confidence), depth_curvature_gradient.frag (+ hrbfbase.glsl gradient / Hessian).
    With the literal window loops the oracle must reproduce the shader text BIT FOR BIT on the same inputs (vertex, PCA normal,
    radius, curvature, gradient magnitude, optimised normal); the bilateral chain to fp32 round-off (its exp() is the oracle's own
    deterministic one, the CPU-compiled shader calls expf)."""
    cam, depth, _ = _frame(W, H, kind)
    pp = orc.prep_params(cam, W, H)
    a = orc.preprocess(pp, depth)
    b = rg.preprocess(pp, depth)                         # the shader chain end to end
    assert np.array_equal(a["metric"], b["metric"])
    for k in ("filtered", "metric_filtered", "vertex_raw", "vertex_filtered"):
        assert np.array_equal(a[k] == 0, b[k] == 0), k
        np.testing.assert_allclose(b[k], a[k], rtol=2e-6, atol=0, err_msg=k)
    # one shader pass at a time on the oracle's own inputs: identical, not just close
    s1 = rg.preprocess_stage(pp, "vertex_normal_radius", a)
    s2 = rg.preprocess_stage(pp, "curvature_gradient", a)
    assert _identical(a["vertex_raw"], s1["vertex_raw"]) >= 0.999 and np.abs(a["vertex_raw"] - s1["vertex_raw"]).max() <= 1e-6     # confidence: exp
    for k in ("vertex_filtered", "normal_pca", "radius"):
        assert _identical(a[k], s1[k]) == 1.0, k
    for k in ("curv1", "curv2", "gradient_mag", "normal_opt"):
        assert _identical(a[k], s2[k]) == 1.0, k
    assert (np.linalg.norm(a["normal_pca"][..., :3], axis=-1) > 0).mean() > 0.8          # the comparison is not vacuous
    assert (np.abs(a["curv1"][..., 3]) < 300).mean() > 0.8


def test_intended_vs_literal_windows_deviation_is_bounded(orc):
    """What the oracle's (and the CUDA kernels') integer windows change against the shaders' float-counter loops, which drop the
    last column / row of the 7x7 window for about 40 % of the columns (stated deviation): measured and bounded here."""
    W, H = 320, 240
    cam, depth, _ = _frame(W, H, "room")
    pp = orc.prep_params(cam, W, H)
    orc.lib().orc_set_float_loops(0)
    try:
        ideal = orc.preprocess(pp, depth)
        orc.lib().orc_set_float_loops(1)
        lit = orc.preprocess(pp, depth)
    finally:
        from tests.conftest import default_float_loops
        orc.lib().orc_set_float_loops(default_float_loops())
    for k in ("filtered", "metric", "metric_filtered", "vertex_raw", "vertex_filtered"):          # no window loops of that kind there
        assert np.array_equal(ideal[k], lit[k]), k
    n0, n1 = ideal["normal_pca"][..., :3], lit["normal_pca"][..., :3]
    v0, v1 = np.linalg.norm(n0, axis=-1) > 0, np.linalg.norm(n1, axis=-1) > 0
    assert np.mean(v0 != v1) <= 1e-3
    ang = np.degrees(np.arccos(np.clip((n0 * n1).sum(-1)[v0 & v1], -1, 1)))
    assert np.median(ang) <= 0.5 and np.percentile(ang, 99) <= 3.0, (np.median(ang), np.percentile(ang, 99))
    k0, k1 = ideal["curv1"][..., 3], lit["curv1"][..., 3]
    assert np.mean((np.abs(k0) < 300) != (np.abs(k1) < 300)) <= 2e-3


@pytest.mark.parametrize("useConfEval", [0, 1])
def test_vertex_confidence_and_fill_in_oracle_matches_reference_shaders(orc, useConfEval):
    """depth_confidence_evaluation.frag and the four FillIn shaders (fill_vertex / fill_normal / fill_curvature / fill_rgb.frag) on the
    state of the oracle pipeline after a few frames: a prediction with holes (young map: confidence below the prediction threshold
    in places) + the current frame.  Pass-through copies: bit-exact; the icp weight (exp) to fp32 round-off."""
    from oracle import orc_pipeline as op
    W, H = 320, 240
    cam = synth.default_camera(W, H)
    sc = synth.Scene("room")
    f = op.HRBFFusion(W, H, cam, icpWeight=100.0, so3=False)
    for i, p in enumerate(synth.circle_trajectory(6, frames_per_rev=120)):      # confidence passes the prediction threshold (3) after 4 frames
        depth, rgb = synth.render_depth(sc, p, W, H, cam, noise=True, seed=i)
        f.processFrame(rgb, depth)
    fr, pred, pp = f.last["frame"], f.pred, f.pp
    # VertexConfidence
    for weighting in (1.0, 0.5):
        a = orc.vertexConfidence(pp, fr["gradient_mag"], weighting, useConfEval, 1000.0)
        b = rg.vertexConfidence(pp, fr["gradient_mag"], fr["metric"], weighting, useConfEval, 1000.0)
        assert a.max() > 0
        np.testing.assert_allclose(b, a, rtol=2e-6, atol=1e-12)
    conf = orc.vertexConfidence(pp, fr["gradient_mag"], 1.0, useConfEval, 1000.0)
    holes = pred["vertex"][..., 2] == 0
    assert holes.sum() > 1000 and (~holes).sum() > 1000   # both branches of every fill shader are exercised
    for passthrough in (0, 1):
        a = orc.fillIn(pp, pred, fr, conf, rgb, passthrough, 10.0, 300.0)
        b = rg.fillIn(pp, pred, fr, conf, rgb, passthrough, 10.0, 300.0)
        for k in ("vertex", "normal", "curvk1", "curvk2", "image"):
            assert _identical(a[k].astype(np.float32), b[k].astype(np.float32)) == 1.0, (k, passthrough)
        assert np.array_equal(a["icpw"] == 0, b["icpw"] == 0)
        np.testing.assert_allclose(b["icpw"], a["icpw"], rtol=3e-6, atol=0)


@pytest.mark.parametrize("W,H,kind", [(160, 120, "room"), (640, 480, "room"), (640, 480, "plane"), (1280, 960, "room")])
def test_predict_indices_oracle_matches_reference_vertex_shader(orc, W, H, kind):
    """row 7: Shaders/index_map.vert per surfel (projection, depth / sub-map culling, normal rotation); the point rasterisation and
    the depth test are fixed-function GL, restated in the driver with the rules the oracle states.  The shader goes through
    normalised device coordinates, the oracle projects directly: a surfel within an ulp of a pixel boundary may land next door."""
    s, pose, cam = _scene(W, H, kind, 1)
    s = s.copy()
    s[::7, 5] = 3.0                                        # sub-map 3 is not active: culled by the key-frame mask
    ak = np.zeros(19200, np.float32); ak[0] = 1.0
    a = orc.predictIndices(pose, s, cam, W, H, maxDepth=2.5, active_kf=ak)      # a depth cut-off inside the scene
    b = rg.predictIndices(pose, s, cam, W, H, maxDepth=2.5, active_kf=ak)
    same = a["index"] == b["index"]
    assert (a["index"] > 0).mean() > 0.5 and same.mean() >= 0.9999
    assert np.array_equal(a["index"] > 0, b["index"] > 0) or np.mean((a["index"] > 0) != (b["index"] > 0)) < 1e-4
    assert not np.isin(a["index"][a["index"] > 0], np.arange(0, len(s), 7)).any()            # none of the culled sub-map
    for k in ("colorTime", "curvMax", "curvMin"):
        assert np.array_equal(a[k][same], b[k][same]), k
    for k in ("vertConf", "normRad"):
        assert np.abs(a[k] - b[k])[same].max() <= 1e-6, k


@pytest.mark.parametrize("W,H", [(320, 240), (640, 480), (1280, 960)])
def test_fuse_and_clean_oracle_match_reference_vertex_shaders(orc, literal_windows, W, H):
    """rows 8-9: GlobalModel::fuse = data.vert per pixel (association, candidate record) + the first-fragment-wins scatter into the
    update textures (fixed function, restated in oracle/refglsl_py.py) + update.vert per surfel (merge); GlobalModel::clean =
    copy_unstable.vert / .geom over the model and the recorded vertices.  On the state of the oracle pipeline after several
    frames, with the literal window loops (the fuse pass recomputes the PCA normal): same merges, same new surfels, every
    attribute bit-identical except the positions (the shader multiplies pose * vec4 column by column: 2 ulps), and clean
    returns the identical surfel array.  Run at every BASELINE image width: the float-counter loops of data.vert:137-138 and
    copy_unstable.vert:106-108 overshoot depending on the texture size (the effect that bit row 10 at 640 pixels)."""
    from oracle import orc_pipeline as op
    cam = synth.default_camera(W, H)
    sc = synth.Scene("room")
    f = op.HRBFFusion(W, H, cam, icpWeight=100.0, so3=False)
    poses = synth.circle_trajectory(8, frames_per_rev=120)
    n_warm = 6 if W <= 640 else 5       # (1280x960: one frame less of the slow CPU pipeline; confidence still passes the thresholds)
    for i, p in enumerate(poses[:6]):
        depth, rgb = synth.render_depth(sc, p, W, H, cam, noise=True, seed=i)
        f.processFrame(rgb, depth)
    for t in (6, 7):                                            # both parities of the 1/4-pixel candidate pattern
        depth, rgb = synth.render_depth(sc, poses[t], W, H, cam, noise=True, seed=t)
        fr = orc.preprocess(f.pp, depth)
        conf = orc.vertexConfidence(f.pp, fr["gradient_mag"], 1.0)
        pose, tick, s0 = f.currPose.copy(), f.tick + (t - 6), f.surfels.copy()
        idx = orc.predictIndices(pose, s0, cam, W, H, f.maxDepthProcessed)
        a_s, a_u = orc.modelFuse(f.mp, pose, tick, rgb, fr, conf, idx, 0, s0)
        b_s, b_u = rg.modelFuse(f.mp, pose, tick, rgb, fr, conf, idx, 0, s0)
        assert a_u.shape == b_u.shape and (a_u[:, 7] == -1).sum() > 1000 and (a_u[:, 7] == -2).sum() > 0
        assert np.array_equal(a_u[:, 3:], b_u[:, 3:], equal_nan=True)           # confidence, colour, sub-map, times, normal, radius, curvatures
        assert np.abs(a_u[:, :3] - b_u[:, :3]).max() <= 1e-6
        assert (a_s != s0).any(1).sum() > 1000                                   # surfels were merged
        assert np.array_equal(a_s[:, 3:], b_s[:, 3:], equal_nan=True)
        assert np.abs(a_s[:, :3] - b_s[:, :3]).max() <= 1e-6
        idx2 = orc.predictIndices(pose, a_s, cam, W, H, f.maxDepthProcessed)
        a_c = orc.modelClean(f.mp, pose, tick, idx2, a_s, a_u)
        b_c = rg.modelClean(f.mp, pose, tick, idx2, a_s, a_u)
        assert len(a_c) < len(a_s) + len(a_u)                                    # something was dropped (the merged markers at least)
        assert a_c.shape == b_c.shape and np.array_equal(a_c, b_c, equal_nan=True)


@pytest.mark.parametrize("useConfEval", [0, 1])
def test_model_initialise_oracle_matches_reference_vertex_shader(orc, useConfEval):
    """GlobalModel::initialise = init_unstableTex.vert / .geom per pixel: same surfels in the same order, every attribute
    bit-identical except the confidence (an exp: fp32 round-off)."""
    W, H = 320, 240
    cam, depth, rgb = _frame(W, H, "room", seed=0)
    pp, mp = orc.prep_params(cam, W, H), orc.model_params(cam, W, H)
    fr = orc.preprocess(pp, depth)
    pose = synth.make_pose(0.02, -0.03, 0.01, (0.05, -0.02, 0.0))
    a = orc.modelInitialise(mp, pose, fr, rgb, useConfEval, 1000.0)
    b = rg.modelInitialise(mp, pose, fr, rgb, useConfEval, 1000.0)
    assert a.shape == b.shape and a.shape[0] > 0.5 * W * H
    cols = [c for c in range(20) if c != 3]
    assert np.array_equal(a[:, cols], b[:, cols], equal_nan=True)
    np.testing.assert_allclose(b[:, 3], a[:, 3], rtol=1e-6, atol=1e-9)


class _ShaderOrc:
    """oracle/orc_py with every GLSL pass replaced by the reference's own shader (the tracker, rows 1-5, stays the oracle's)"""

    def __init__(self, orc):
        self._orc, self._metric = orc, None

    def __getattr__(self, k):
        return getattr(self._orc, k)

    def preprocess(self, pp, depth):
        fr = rg.preprocess(pp, depth)
        self._metric = fr["metric"]
        return fr

    def vertexConfidence(self, pp, gm, weighting, useConfEval=0, epsilon=1000.0):
        return rg.vertexConfidence(pp, gm, self._metric, weighting, useConfEval, epsilon)

    modelInitialise = staticmethod(rg.modelInitialise)
    predictIndices = staticmethod(rg.predictIndices)
    modelFuse = staticmethod(rg.modelFuse)
    modelClean = staticmethod(rg.modelClean)
    predictHRBF = staticmethod(rg.predictHRBF)
    fillIn = staticmethod(rg.fillIn)


def _run_pipeline(orc, mode, n, W=320, H=240, **kw):
    from oracle import orc_pipeline as op
    cam = synth.default_camera(W, H)
    sc = synth.Scene("room")
    saved = op.orc
    try:
        if mode == "shader":
            op.orc = _ShaderOrc(orc)
        orc.lib().orc_set_float_loops(1 if mode == "literal" else 0)
        f = op.HRBFFusion(W, H, cam, **kw)
        poses = []
        for i, p in enumerate(synth.circle_trajectory(n, frames_per_rev=120)):
            depth, rgb = synth.render_depth(sc, p, W, H, cam, noise=True, seed=i)
            poses.append(f.processFrame(rgb, depth).copy())
        return poses, f.surfels.shape[0]
    finally:
        from tests.conftest import default_float_loops
        op.orc = saved
        orc.lib().orc_set_float_loops(default_float_loops())


def test_frame_loop_driven_by_reference_shaders_tracks_like_the_oracle(orc):
    """The whole per-frame loop (HRBFFusion::processFrame) with every GLSL pass executed by the reference's own shader text and only
    the tracker taken from the oracle, against the oracle pipeline, free-running over 6 frames (ICP-only: the configuration the
    north-star tolerance is asserted on).  Two comparisons:
      literal-window oracle vs shader loop: every pass is bit-identical or within fp32 round-off on equal inputs, what is left is
        how the tracker amplifies that round-off (5e-7 relative in the filtered depth -> ~5e-5 in the pose at 320x240);
      intended-window oracle (the default, what the CUDA kernels implement) vs literal: the stated deviation, in pose units."""
    from tests.util import pose_err
    kw = dict(icpWeight=100.0, so3=False)
    n = 6
    lit, n_lit = _run_pipeline(orc, "literal", n, **kw)
    sha, n_sha = _run_pipeline(orc, "shader", n, **kw)
    ide, n_ide = _run_pipeline(orc, "intended", n, **kw)
    assert abs(n_lit - n_sha) <= max(5, int(1e-3 * n_lit)) and abs(n_ide - n_lit) <= max(5, int(3e-3 * n_lit))
    worst_shader = worst_dev = 0.0
    for i in range(1, n):
        worst_shader = max(worst_shader, *pose_err(lit[i][:3, :3], lit[i][:3, 3], sha[i][:3, :3], sha[i][:3, 3]))
        worst_dev = max(worst_dev, *pose_err(ide[i][:3, :3], ide[i][:3, 3], lit[i][:3, :3], lit[i][:3, 3]))
    print(f"pose after {n} free-running frames: literal oracle vs shader loop {worst_shader:.2e}; intended vs literal windows {worst_dev:.2e}")
    assert worst_shader <= 3e-4          # measured 7e-5
    assert worst_dev <= 3e-3             # measured 9e-4: the price of the intended-window restatement


def test_dense_enough_oracle_matches_reference_resize_shader(orc):
    """denseEnough: resize.frag over a 32x24 viewport samples texel (20 i + 10, 20 j + 10) of the predicted vertex map; the decision
    flips exactly where the oracle's does when valid samples are removed one by one around the 75 % threshold."""
    W, H = 640, 480
    rng = np.random.default_rng(4)
    v = np.zeros((H, W, 4), np.float32)
    v[..., 2] = rng.uniform(0.5, 3.0, (H, W)).astype(np.float32)
    flag, sampled = rg.denseEnough(v)
    assert sampled.shape == (24, 32, 4) and np.array_equal(sampled, v[10::20, 10::20])
    ys, xs = np.meshgrid(np.arange(24), np.arange(32), indexing="ij")
    order = rng.permutation(24 * 32)
    for k in range(0, 260):                                  # 768 samples: the threshold sits at 576 valid ones
        j = order[k]
        v[20 * ys.ravel()[j] + 10, 20 * xs.ravel()[j] + 10, 2] = 0.0
        if k >= 180:
            assert rg.denseEnough(v)[0] == orc.denseEnough(v), k
    assert rg.denseEnough(v)[0] is False and orc.denseEnough(v) is False
    # a pixel that is not a sample does not matter
    v2 = v.copy(); v2[11::20, 11::20, 2] = 0.0
    assert np.array_equal(rg.denseEnough(v2)[1], rg.denseEnough(v)[1])


def test_update_model_oracle_matches_reference_vertex_shader(orc):
    """GlobalModel::updateModel = update_delta_trans.vert per surfel (rigid correction of its sub-map, read from the
    19200 x 1 DeltaTransformKF texture): bit-identical, everything but position and normal passes through."""
    rng = np.random.default_rng(9)
    n, k = 5000, 7
    s = rng.standard_normal((n, 20)).astype(np.float32)
    s[:, 5] = rng.integers(0, k, n).astype(np.float32)
    delta = np.stack([synth.make_pose(*rng.uniform(-0.05, 0.05, 3), tuple(rng.uniform(-0.1, 0.1, 3))) for _ in range(k)]).astype(np.float32)
    a, b = orc.modelUpdate(s, delta), rg.modelUpdate(s, delta)
    assert np.array_equal(a, b)
    assert not np.array_equal(a[:, :3], s[:, :3]) and np.array_equal(a[:, [3, 4, 5, 6, 7, 11] + list(range(12, 20))], s[:, [3, 4, 5, 6, 7, 11] + list(range(12, 20))])
