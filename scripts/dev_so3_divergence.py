"""Development: free-running CUDA trackers (persistent kernel / kernel-per-reduction graph) vs the oracle on identical inputs,
for the option combinations around SO3 pre-alignment."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import orc_py as orc
from hrbffusion3d_b200 import odometry as od
from tests.test_gpu_odometry import _pipeline_frame1_inputs, _init_tracker
from tests.util import pose_err, pair

dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
for W, H in ((640, 480), (320, 240)):
    cam, pose, d = _pipeline_frame1_inputs(orc, W, H)
    for name, kw in (("icp", dict(icpWeight=100.0, so3=False)), ("icp+so3", dict(icpWeight=100.0, so3=True)), ("rgb+icp", dict(icpWeight=10.0, so3=False)),
                     ("default", dict()), ("default nopyr", dict(pyramid=False)), ("rgbOnly+so3", dict(rgbOnly=True))):
        to, Ro, so = _init_tracker(orc.Odometry(W, H, cam[2], cam[3], cam[0], cam[1]), lambda a: a, pose, d).getIncrementalTransformation(pose[:3, 3], pose[:3, :3], **kw)
        out = []
        for graph in (False, True):
            g = _init_tracker(od.RGBDOdometry(W, H, cam[2], cam[3], cam[0], cam[1]), dev, pose, d)
            g.setTracker(graph)
            tg, Rg, sg = g.getIncrementalTransformation(pose[:3, 3], pose[:3, :3], **kw)
            out.append("%s ang %.1e t %.1e (so3 err %.6g cnt %.0f | oracle %.6g %.0f)" % (("graph" if graph else "persistent",) + pose_err(Ro, to, Rg, tg) + (sg.lastSO3Error, sg.lastSO3Count, so.lastSO3Error, so.lastSO3Count)))
        print(f"{W}x{H} {name:14s}: " + " ; ".join(out))
