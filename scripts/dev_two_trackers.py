"""Do two persistent (cooperative) trackers of different odometry objects run concurrently on one GPU? (development)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hrbffusion3d_b200 import odometry as od
from tests.util import pair
W, H = 640, 480
m0, pose0, m1, pose1, cam = pair(W, H)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
d0 = {k: dev(v) for k, v in m0.items()}; d1 = {k: dev(v) for k, v in m1.items()}
def make(tt):
    go = od.RGBDOdometry(W, H, cam[2], cam[3], cam[0], cam[1]); go.setTrackerThreads(tt)
    go.initFirstRGB(d0["rgba"]); go.initICPModel(d0["vertex"], d0["normal"], 20.0, pose0); go.initRGBModel(d0["rgba"]); go.initCurvatureModel(d0["k1"], d0["k2"], pose0)
    go.initICP(d1["vertex"], d1["normal"], 20.0); go.initRGB(d1["rgba"]); go.initCurvature(d1["k1"], d1["k2"]); go.initICPweight(d0["icpw"])
    return go
pin = torch.from_numpy(np.concatenate([pose0[:3, :3].reshape(-1), pose0[:3, 3]]).astype(np.float32)).cuda()
for tt in (256, 512):
    for K in (1, 2, 3):
        gos = [make(tt) for _ in range(K)]; st = [torch.cuda.Stream() for _ in range(K)]; pouts = [torch.zeros(12, device="cuda") for _ in range(K)]
        def rnd(n):
            for _ in range(n):
                for k in range(K):
                    with torch.cuda.stream(st[k]): gos[k].trackAsync(pin, pouts[k], icpWeight=10.0, so3=False)
        rnd(5); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); n = 100; rnd(n)
        for k in range(K): torch.cuda.current_stream().wait_stream(st[k])
        e1.record(); torch.cuda.synchronize()
        print("TT=%d, %d tracker(s) on %d stream(s): %.1f us per round of %d tracking call(s)" % (tt, K, K, e0.elapsed_time(e1) / n * 1e3, K), flush=True)
