"""frames/s of the staged pipeline vs plain processFrameDev (development)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from hrbffusion3d_b200.fusion import HRBFFusion
n = 60
depth, rgb, poses, cam = bench.make_sequence(0, bench.RING, only=n)
d = torch.from_numpy(depth.view(np.int16)).cuda(); c = torch.from_numpy(rgb).cuda()
def run(staged, steps=200, warm=20):
    F = HRBFFusion(bench.W, bench.H, cam, capacity=1 << 22)
    def step(i):
        if staged:
            F.stageFrame(c[(i + 1) % n], d[(i + 1) % n]); F.processStaged(None)
        else:
            F.processFrameDev(c[i % n], d[i % n])
    if staged: F.stageFrame(c[0], d[0])
    for i in range(warm): step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(warm, warm + steps): step(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e3
print("plain  %.1f us/frame" % run(False))
print("staged %.1f us/frame" % run(True))
