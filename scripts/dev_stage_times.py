"""Per-stage CUDA-event timings of the frame pipeline (the reference's Stopwatch spans)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from hrbffusion3d_b200.fusion import HRBFFusion
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
bench.RING = n
depth, rgb, poses, cam = bench.make_sequence(0, n)
F = HRBFFusion(bench.W, bench.H, cam, capacity=1 << 22)
F.enableTimings(True)
acc = {}
for i in range(n):
    F.processFrame(rgb[i], depth[i])
    if i >= 10:
        for k, v in F.lastTimings().items(): acc.setdefault(k, []).append(v)
for k, v in acc.items(): print(f"{k:15s} {np.mean(v)*1e3:8.1f} us")
print("total %.1f us; surfels %d" % (sum(np.mean(v) for v in acc.values()) * 1e3, F.globalModel.lastCount()))

# ---- individual calls in steady state (back-to-back, CUDA events) ----
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
im, fr, gm, fi = F.indexMap, F.frame, F.globalModel, F.fillIn
pose = F.getPose()
surf, cnt = gm.model()
print("predictHRBF      %.1f us" % timeit(lambda: im.predictHRBF(0)))
print("predictIndices   %.1f us" % timeit(lambda: im.predictIndices(pose, 1, 1, (surf, cnt), 20.0)))
print("fillIn           %.1f us" % timeit(lambda: fi.run(im, fr)))
print("preprocess       %.1f us" % timeit(lambda: fr.preprocess()))
from hrbffusion3d_b200.odometry import RGBDOdometry
import ctypes as C
from hrbffusion3d_b200._lib import lib
od = RGBDOdometry.__new__(RGBDOdometry); od._h = C.c_void_p(lib().hrbf_fusion_odometry(F._h)); od.width, od.height = bench.W, bench.H; od.close = lambda: None
v, n, k1, k2, w, rgba = fr.tex("VERTEX_FILTERED"), fr.tex("NORMAL"), fr.tex("PRINCIPAL_CURV1"), fr.tex("PRINCIPAL_CURV2"), im.tex("icpweightHRBF"), fr.tex("RGBA")
def prep():
    od.initICPModel(im.tex("vertexHRBF"), im.tex("normalHRBF"), 20.0, pose); od.initRGBModel(im.tex("imageHRBF")); od.initCurvatureModel(im.tex("curvk1HRBF"), im.tex("curvk2HRBF"), pose)
    od.initICP(v, n, 20.0); od.initRGB(rgba); od.initCurvature(k1, k2); od.initICPweight(w)
print("7 init* calls    %.1f us" % timeit(prep))
pin = torch.from_numpy(np.concatenate([pose[:3, :3].reshape(-1), pose[:3, 3]]).astype(np.float32)).cuda(); pout = torch.zeros(12, device="cuda")
print("track (default)  %.1f us" % timeit(lambda: od.trackAsync(pin, pout)))
print("track (icp only) %.1f us" % timeit(lambda: od.trackAsync(pin, pout, icpWeight=100.0, so3=False)))
