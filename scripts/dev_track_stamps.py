"""Development: %globaltimer stamps of the persistent tracker (one CTA's thread 0), averaged per pyramid level and transition.
Stamp ids (track_persistent.cuh): 1 iteration start, 7 after the RGB residual pass, 8 its integer pair published, 2 after the ICP pass,
3 ICP partial published, 9 integer all-reduce done, 10 sigma known, 11 after the RGB step pass, 4 RGB partial published,
5 float all-reduce done, 12 normal equations built, 13 solved, 14 rodrigues, 15 pose composed, 6 iteration end."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
from hrbffusion3d_b200 import odometry as od
from hrbffusion3d_b200._lib import lib, check
from tests.util import pair

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (640, 480)
m0, pose0, m1, pose1, cam = pair(W, H)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
d0 = {k: dev(v) for k, v in m0.items()}
d1 = {k: dev(v) for k, v in m1.items()}
go = od.RGBDOdometry(W, H, cam[2], cam[3], cam[0], cam[1])
go.initFirstRGB(d0["rgba"])
go.initICPModel(d0["vertex"], d0["normal"], 20.0, pose0); go.initRGBModel(d0["rgba"]); go.initCurvatureModel(d0["k1"], d0["k2"], pose0)
go.initICP(d1["vertex"], d1["normal"], 20.0); go.initRGB(d1["rgba"]); go.initCurvature(d1["k1"], d1["k2"]); go.initICPweight(d0["icpw"])
NAMES = {1: "start", 7: "res pass", 8: "int pub", 2: "icp pass", 3: "icp pub", 9: "int allred", 10: "sigma", 11: "rgb step", 4: "rgb pub", 5: "f allred",
         12: "build Ab", 13: "solve", 14: "rodrigues", 15: "compose", 6: "end"}
for name, kw in (("icp-only", dict(icpWeight=100.0, so3=False)), ("default", dict(icpWeight=10.0, so3=True))):
    for _ in range(3):
        go.getIncrementalTransformation(pose0[:3, 3], pose0[:3, :3], **kw); go.initRGB(d1["rgba"])
    buf = (C.c_longlong * 512)()
    check(lib().hrbf_odometry_debug_stamps(go._h, buf, 512))        # arms
    go.getIncrementalTransformation(pose0[:3, 3], pose0[:3, :3], **kw); go.initRGB(d1["rgba"])
    torch.cuda.synchronize()
    check(lib().hrbf_odometry_debug_stamps(go._h, buf, 512))
    st = [(v >> 56, v & ((1 << 56) - 1)) for v in buf if v]
    print(f"== {name}: {len(st)} stamps, first -> last {(st[-1][1] - st[0][1]) / 1e3:.1f} us")
    # split into iterations at stamp 1; level from the iteration index (4 / 5 / 10)
    its, cur = [], []
    for k, t in st:
        if k == 1 and cur: its.append(cur); cur = []
        cur.append((k, t))
    its.append(cur)
    lv = [2] * 4 + [1] * 5 + [0] * 10
    for L in (2, 1, 0):
        sel = [it for it, l in zip(its, lv) if l == L]
        if not sel: continue
        ks = [k for k, _ in sel[0]]
        tot = np.mean([(it[-1][1] - it[0][1]) / 1e3 for it in sel])
        parts = []
        for j in range(1, len(ks)):
            d = np.mean([(it[j][1] - it[j - 1][1]) / 1e3 for it in sel if len(it) == len(ks)])
            parts.append(f"{NAMES.get(ks[j], ks[j])} {d:.2f}")
        print(f"  level {L}: {len(sel)} iterations, {tot:.2f} us each: " + " | ".join(parts))
    gaps = [(its[i + 1][0][1] - its[i][-1][1]) / 1e3 for i in range(len(its) - 1)]
    print("  between iterations (us):", np.round(gaps, 2))
