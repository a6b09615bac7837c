import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
from hrbffusion3d_b200 import odometry as od
from hrbffusion3d_b200._lib import lib, check
from tests.util import pair
W,H=640,480
m0, pose0, m1, pose1, cam = pair(W, H)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
d0 = {k: dev(v) for k, v in m0.items()}; d1 = {k: dev(v) for k, v in m1.items()}
go = od.RGBDOdometry(W, H, cam[2], cam[3], cam[0], cam[1])
go.initFirstRGB(d0["rgba"])
go.initICPModel(d0["vertex"], d0["normal"], 20.0, pose0); go.initRGBModel(d0["rgba"]); go.initCurvatureModel(d0["k1"], d0["k2"], pose0)
go.initICP(d1["vertex"], d1["normal"], 20.0); go.initRGB(d1["rgba"]); go.initCurvature(d1["k1"], d1["k2"]); go.initICPweight(d0["icpw"])
buf = (C.c_longlong * 512)()

check(lib().hrbf_odometry_debug_stamps(go._h, buf, 512))
for kw in (dict(icpWeight=100.0, so3=False), dict(icpWeight=10.0, so3=True)):
    for _ in range(3): go.getIncrementalTransformation(pose0[:3, 3], pose0[:3, :3], **kw)
    check(lib().hrbf_odometry_debug_stamps(go._h, buf, 512))
    a = np.array(buf[:], dtype=np.int64); a = a[a != 0]
    slot = (a >> 56).astype(int); t = (a & ((1 << 56) - 1)).astype(np.int64)
    print(kw, "stamps", len(a), "total us", (t[-1] - t[0]) / 1e3)
    names = {2: "icp pass", 3: "block reduce", 4: "block reduce (rgb) / none", 5: "all-reduce", 6: "solve", 1: "loop overhead", 7: "residual pass", 8: "cta int reduce + publish", 9: "int all-reduce (wait for all CTAs)", 10: "sigma", 11: "step pass", 12: "gn: normal equations (lanes)", 13: "gn: LDLT", 14: "gn: rodrigues", 15: "gn: compose"}
    agg = {}
    for i in range(1, len(a)):
        agg.setdefault((slot[i - 1], slot[i]), []).append((t[i] - t[i - 1]) / 1e3)
    for k, v in sorted(agg.items()):
        print("   %d->%d %-28s n=%3d mean %.2f us  first10 %s" % (k[0], k[1], names.get(k[1], ""), len(v), np.mean(v), np.round(v[:10], 2)))
