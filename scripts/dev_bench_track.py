"""Development micro-benchmark of the tracking path (not the contract bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hrbffusion3d_b200 import odometry as od
from tests.util import pair

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (640, 480)
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 50
m0, pose0, m1, pose1, cam = pair(W, H)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
d0 = {k: dev(v) for k, v in m0.items()}
d1 = {k: dev(v) for k, v in m1.items()}
go = od.RGBDOdometry(W, H, cam[2], cam[3], cam[0], cam[1])
go.initFirstRGB(d0["rgba"])

def prep():
    go.initICPModel(d0["vertex"], d0["normal"], 20.0, pose0)
    go.initRGBModel(d0["rgba"])
    go.initCurvatureModel(d0["k1"], d0["k2"], pose0)
    go.initICP(d1["vertex"], d1["normal"], 20.0)
    go.initRGB(d1["rgba"])
    go.initCurvature(d1["k1"], d1["k2"])
    go.initICPweight(d0["icpw"])

def timeit(fn, n):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

prep()
print(f"{W}x{H}")
print("prep (7 init calls): %.1f us" % timeit(prep, iters))
for name, kw in (("icp-only", dict(icpWeight=100.0, so3=False)), ("default rgb+icp+so3", dict(icpWeight=10.0, so3=True))):
    f = lambda: go.getIncrementalTransformation(pose0[:3, 3], pose0[:3, :3], **kw)
    us = timeit(f, iters)
    print("track %-22s: %.1f us  (%d kernels)" % (name, us, go.last_stats.kernel_launches))
pin = torch.from_numpy(np.concatenate([pose0[:3, :3].reshape(-1), pose0[:3, 3]]).astype(np.float32)).cuda()
pout = torch.zeros(12, device="cuda")
for name, kw in (("icp-only", dict(icpWeight=100.0, so3=False)), ("default", dict(icpWeight=10.0, so3=True))):
    f = lambda: go.trackAsync(pin, pout, **kw)
    print("trackAsync %-10s: %.1f us" % (name, timeit(f, iters)))
import ctypes as C
from hrbffusion3d_b200._lib import lib, check, stream_ptr
go.getIncrementalTransformation(pose0[:3, 3], pose0[:3, :3], icpWeight=10.0, so3=True)
print("isolated kernels, back-to-back (us): which level update -> us   [GB/s algorithmic for ICP = 68 B/px]")
for which, name in ((0, "icp"), (1, "rgbres"), (2, "rgbstep"), (3, "so3")):
    for lvl in (0, 1, 2):
        for upd in (0, 1):
            if which in (1, 3) and upd: continue
            if which == 3 and lvl != 2: continue
            us = C.c_float()
            check(lib().hrbf_odometry_time_kernel(go._h, which, lvl, upd, 200, C.byref(us), stream_ptr()))
            n = (W >> lvl) * (H >> lvl)
            extra = "  %.0f GB/s" % (68.0 * n / us.value * 1e-3) if which == 0 else ""
            print(f"  {name:8s} L{lvl} upd={upd}: {us.value:7.2f} us{extra}")
