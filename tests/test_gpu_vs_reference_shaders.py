"""The CUDA frame pipeline against a frame loop whose every GLSL pass is the REFERENCE's own shader text
(Core/src/Shaders/*.{vert,frag,glsl} compiled for the CPU into oracle/_ref/libref_glsl.so by oracle/build_ref_glsl.py and driven
by oracle/refglsl_py.py + oracle/glsl_drivers/) -- preprocessing, initialise, splat, fuse, clean, HRBF prediction, fill-in,
vertex confidence -- with only the tracker (rows 1-5, pinned to the reference's CUDA kernels elsewhere) taken from the oracle.
This is the test that closes the chain  CUDA == oracle  and  oracle == reference shaders  in ONE comparison, at the BASELINE
resolution, free-running over several frames.  north_star bound: pose 1e-5 (rad, m) per frame, vertices 1e-4 m RMSE."""
import numpy as np
import pytest

from hrbffusion3d_b200 import synth
from oracle import refglsl_py as rg
from tests.util import pose_err

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not rg.available(), reason="oracle/_ref/libref_glsl.so not shipped")]


@pytest.mark.parametrize("kw,pose_tol", [
    (dict(icpWeight=100.0, so3=False), 1e-5),      # ICP + HRBF: the configuration the north-star tolerance is stated on
    ({}, 1e-5),                                    # reference defaults: joint RGB-D + SO3 pre-alignment
])
def test_cuda_pipeline_matches_reference_shader_loop_640x480(orc, cuda, kw, pose_tol):
    from hrbffusion3d_b200.fusion import HRBFFusion
    from oracle import orc_pipeline as op
    from tests.test_oracle_vs_reference_glsl import _ShaderOrc
    W, H, n = 640, 480, 5
    cam = synth.default_camera(W, H)
    sc = synth.Scene("room")
    frames = [synth.render_depth(sc, p, W, H, cam, noise=True, seed=i) for i, p in enumerate(synth.circle_trajectory(n, frames_per_rev=120))]
    gkw = dict(kw)
    if "so3" in gkw:
        gkw["so3"] = int(gkw["so3"])
    gpu = HRBFFusion(W, H, cam, capacity=1 << 20, **gkw)
    saved = op.orc
    op.orc = _ShaderOrc(orc)
    try:
        from tests.util import pipeline_tracker_inputs, tracker_noise_floor
        ref = op.HRBFFusion(W, H, cam, **kw)
        tol = 0.0
        for i, (depth, rgb) in enumerate(frames):
            # ICP + HRBF: the north-star bound as it stands.  With the photometric term the tracker amplifies what one unit in the last
            # place of its inputs does (the shaders' exp() in the bilateral filter is not the kernels' exp: 5e-7 relative in the filtered
            # depth) through the hard roundings of ~5 000 correspondences: the allowance per frame is then 8 x the oracle's own measured
            # sensitivity on this frame (tests/util.tracker_noise_floor), accumulated because the loop is free-running
            floor = 0.0 if (i == 0 or kw.get("icpWeight", 10.0) >= 100) else tracker_noise_floor(orc, W, H, cam, ref.currPose.copy(), pipeline_tracker_inputs(op.orc, ref, frames[i - 1][1], rgb, depth), ref.kw, n=3, seed=i)
            tol = pose_tol if kw.get("icpWeight", 10.0) >= 100 else tol + max(pose_tol, 8 * floor)
            To, Tg = ref.processFrame(rgb, depth), gpu.processFrame(rgb, depth)
            ang, dt = pose_err(To[:3, :3], To[:3, 3], Tg[:3, :3], Tg[:3, 3])
            cnt = gpu.globalModel.lastCount()
            print(f"frame {i}: CUDA vs reference-shader loop: pose diff ang {ang:.2e} t {dt:.2e} (allowed {tol:.1e}; oracle's own 1-ulp sensitivity {floor:.1e}); surfels {cnt} vs {ref.surfels.shape[0]}")
            assert ang <= tol and dt <= tol, (i, ang, dt, tol)
            assert abs(cnt - ref.surfels.shape[0]) <= max(5, int(1e-3 * ref.surfels.shape[0]))
            # the HRBF prediction the next frame is tracked against: vertices within 1e-4 m RMSE where both found a surface
            pv = gpu.indexMap.tex("vertexHRBF").cpu().numpy()
            fg, fo = pv[..., 2] > 0, ref.pred["vertex"][..., 2] > 0
            assert np.mean(fg != fo) < (2e-3 if kw.get("icpWeight", 10.0) >= 100 else 6e-2)      # see tests/test_gpu_fusion.py: confidence threshold of a young map
            both = fg & fo
            if both.sum() > 1000:
                d = (pv[..., :3].astype(np.float64) - ref.pred["vertex"][..., :3])[both]
                bad = np.linalg.norm(d, axis=-1) > 1e-3       # depth-discontinuity pixels where another crossing of f is picked (stated)
                icp_only = kw.get("icpWeight", 10.0) >= 100
                # photometric configuration: poses ~1e-4 apart -> fusion weights ~1e-2 apart -> in a young map the surfels on the prediction's
                # confidence threshold (3) enter the neighbourhoods of one side only (same effect as the found flags above)
                assert bad.mean() < (1e-3 if icp_only else 6e-2)
                assert np.sqrt((d[~bad] ** 2).sum(-1).mean()) <= (1e-4 if icp_only else 3e-4)
    finally:
        op.orc = saved
