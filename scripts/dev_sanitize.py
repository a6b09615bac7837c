"""Development: a short staged + unstaged pipeline run for compute-sanitizer (memcheck / racecheck): 6 frames at two sizes, default configuration."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hrbffusion3d_b200 import synth
from hrbffusion3d_b200.fusion import HRBFFusion
for W, H in ((320, 240), (640, 480)):
    cam = synth.default_camera(W, H)
    poses = synth.circle_trajectory(6, frames_per_rev=120)
    sc = synth.Scene("room")
    fr = [synth.render_depth(sc, p, W, H, cam, noise=True, seed=i) for i, p in enumerate(poses)]
    d = [torch.from_numpy(f[0].view(np.int16)).cuda() for f in fr]
    c = [torch.from_numpy(f[1]).cuda() for f in fr]
    F = HRBFFusion(W, H, cam, capacity=1 << 19)
    F.stageFrame(c[0], d[0])
    for i in range(6):
        if i + 1 < 6: F.stageFrame(c[i + 1], d[i + 1])
        F.processStaged(None)
    torch.cuda.synchronize()
    G = HRBFFusion(W, H, cam, capacity=1 << 19, trackerThreads=256)
    for i in range(4): G.processFrameDev(c[i], d[i])
    torch.cuda.synchronize()
    print(W, H, "surfels", F.globalModel.lastCount(), G.globalModel.lastCount(), flush=True)
print("done")
