"""Development: where do the CUDA and the oracle tracker part on a default-configuration (RGB-D + ICP + SO3) frame?
Teacher-forced per-iteration comparison (tests/gn_loop_py.py) + free-running poses, on the pipeline's own frame-1 inputs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import orc_py as orc, orc_pipeline as op
from hrbffusion3d_b200 import synth, odometry as od
from tests import gn_loop_py as gn
from tests.util import pose_err

W, H = 640, 480
cam = synth.default_camera(W, H); sc = synth.Scene("room")
frames = [synth.render_depth(sc, p, W, H, cam, noise=True, seed=i) for i, p in enumerate(synth.circle_trajectory(3, frames_per_rev=120))]
f = op.HRBFFusion(W, H, cam)
f.processFrame(frames[0][1], frames[0][0])
depth, rgb = frames[1]
fr = orc.preprocess(f.pp, depth)
fill = not orc.denseEnough(f.pred["vertex"]); src = f.fill if fill else f.pred
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()

def mk(which):
    o = orc.Odometry(W, H, cam[2], cam[3], cam[0], cam[1]) if which == "orc" else od.RGBDOdometry(W, H, cam[2], cam[3], cam[0], cam[1])
    g = (lambda a: a) if which == "orc" else dev
    o.initFirstRGB(g(f.rgba(frames[0][1])))
    o.initICPModel(g(src["vertex"]), g(src["normal"]), 20.0, f.currPose); o.initRGBModel(g(src["image"])); o.initCurvatureModel(g(src["curvk1"]), g(src["curvk2"]), f.currPose)
    o.initICP(g(fr["vertex_filtered"]), g(fr["normal"]), 20.0); o.initRGB(g(f.rgba(rgb))); o.initCurvature(g(fr["curv1"]), g(fr["curv2"])); o.initICPweight(g(src["icpw"]))
    return o

for name, kw in (("icp-only", dict(icpWeight=100.0, so3=False)), ("rgb+icp", dict(icpWeight=10.0, so3=False)), ("default", dict(icpWeight=10.0, so3=True))):
    to, Ro, so = mk("orc").getIncrementalTransformation(f.currPose[:3, 3], f.currPose[:3, :3], **kw)
    tg, Rg, sg = mk("gpu").getIncrementalTransformation(f.currPose[:3, 3], f.currPose[:3, :3], **kw)
    print(f"== {name}: free-running CUDA vs oracle on identical inputs: ang %.2e t %.2e; so3 cnt {so.lastSO3Count} / {sg.lastSO3Count}, rgb cnt {so.lastRGBCount} / {sg.lastRGBCount}, icp cnt {so.lastICPCount} / {sg.lastICPCount}" % pose_err(Ro, to, Rg, tg))
    oo, go = mk("orc"), mk("gpu")
    t, R, log = gn.run(gn.OracleBackend(orc, oo), cam, f.currPose[:3, 3], f.currPose[:3, :3], shadow=gn.CudaBackend(od, go, orc, torch), **kw)
    print("   python loop (oracle steps) vs oracle C loop: ang %.2e t %.2e" % pose_err(R, t, Ro, to))
    for r in log:
        if r["kind"] == "so3":
            d, s = r["d_sums"], r["s_sums"]
            print(f"   so3 it {r['it']}: count {r['d_res'][1]:.0f} / {r['s_res'][1]:.0f}  sums rel {np.abs(d - s).max() / np.abs(d).max():.1e}")
        else:
            line = f"   L{r['level']} it {r['it']}:"
            if "d_sigma" in r: line += f" rgb (sigma, n) {r['d_sigma']},{r['d_count']} / {r['s_sigma']},{r['s_count']}"
            if "d_icp_sums" in r: line += f"  icp n {r['d_icp_res'][1]:.0f} / {r['s_icp_res'][1]:.0f} sums rel {np.abs(r['d_icp_sums'] - r['s_icp_sums'])[:27].max() / np.abs(r['d_icp_sums'][:27]).max():.1e}"
            if "d_rgb_sums" in r: line += f"  rgb sums rel {np.abs(r['d_rgb_sums'] - r['s_rgb_sums'])[:27].max() / max(np.abs(r['d_rgb_sums'][:27]).max(), 1e-30):.1e}"
            print(line)
