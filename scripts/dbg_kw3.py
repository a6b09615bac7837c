import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import orc_py as orc
from tests.test_gpu_odometry import build_both
from tests.util import pose_err
np.set_printoptions(precision=7, suppress=True, linewidth=200)
for kw in (dict(rgbOnly=True, so3=True), dict(rgbOnly=True, so3=False), dict(rgbOnly=True, so3=False, pyramid=False)):
    oo, go, (m0,pose0,m1,pose1,cam) = build_both(orc, torch, 320, 240)
    to,Ro,so = oo.getIncrementalTransformation(pose0[:3,3], pose0[:3,:3], **kw)
    tg,Rg,sg = go.getIncrementalTransformation(pose0[:3,3], pose0[:3,:3], **kw)
    print(kw, pose_err(Ro,to,Rg,tg))
    for s in (so, sg): print("  rgb", s.lastRGBError, s.lastRGBCount, "so3", s.lastSO3Error, s.lastSO3Count, "b", np.array(s.lastb[:]))
