"""Cases that pin row 4 (the Gauss-Newton loop of RGBDOdometry::getIncrementalTransformation) to the REFERENCE's own
tracking loop (oracle/_ref/libref_odometry*.so: Core/src/Utils/RGBDOdometry.cpp compiled verbatim on the reference's own kernels,
oracle/build_ref_odometry.py).  Shared by
  * oracle/gen_ref4_golden.py (runs the reference on a GPU box -> tests/golden/ref_odometry.npz),
  * tests/test_oracle_vs_reference_row4.py (CPU: oracle vs golden; GPU: oracle vs CUDA library vs reference, live, with timings)."""
import os

import numpy as np

from tests.util import init_tracker, pair


def _pair_inputs(W, H):
    m0, pose0, m1, pose1, cam = pair(W, H)
    d = dict(first=m0["rgba"], rgba=m1["rgba"], src=dict(vertex=m0["vertex"], normal=m0["normal"], image=m0["rgba"], curvk1=m0["k1"], curvk2=m0["k2"], icpw=m0["icpw"]),
             fr=dict(vertex_filtered=m1["vertex"], normal=m1["normal"], curv1=m1["k1"], curv2=m1["k2"]))
    return cam, pose0, d


def _gputest_inputs():
    """the reference's GPUTest fixture pair (GPUTest/{1c,1d,2c,2d}.png) under its protocol (tests/gputest_pair.py), with the maps GPUTest
    leaves uninitialised set to neutral values through the regular init calls: curvature 0, icp weight 1"""
    from tests import gputest_pair as gp
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "gputest_pair.npz"))
    V1, N1 = gp.load_vertices(g["d1"])
    V2, N2 = gp.load_vertices(g["d2"])
    zero = np.zeros((480, 640, 4), np.float32)
    d = dict(first=gp.rgba(g["c1"]), rgba=gp.rgba(g["c2"]), src=dict(vertex=V1, normal=N1, image=gp.rgba(g["c1"]), curvk1=zero, curvk2=zero, icpw=np.ones((480, 640), np.float32)),
             fr=dict(vertex_filtered=V2, normal=N2, curv1=zero, curv2=zero))
    return (gp.K[0], gp.K[1], gp.K[2], gp.K[3]), np.eye(4, dtype=np.float32), d


def cases(orc):
    """[(name, W, H, cam, pose, inputs, kwargs)]"""
    from tests.test_gpu_odometry import _pipeline_frame1_inputs
    out = []
    cam, pose, d = _pair_inputs(640, 480)
    for name, kw in (("pair640_icp", dict(icpWeight=100.0, so3=False)), ("pair640_default", dict()), ("pair640_rgbicp", dict(icpWeight=10.0, so3=False)),
                     ("pair640_nopyr_fast", dict(icpWeight=100.0, so3=False, pyramid=False, fastOdom=True)), ("pair640_rgbonly", dict(rgbOnly=True, so3=False))):
        out.append((name, 640, 480, cam, pose, d, kw))
    cam, pose, d = _pair_inputs(320, 240)
    out.append(("pair320_default", 320, 240, cam, pose, d, dict()))
    cam, pose, d = _pipeline_frame1_inputs(orc, 640, 480)
    out.append(("pipeline640_icp", 640, 480, cam, pose, d, dict(icpWeight=100.0, so3=False)))
    out.append(("pipeline640_default", 640, 480, cam, pose, d, dict()))
    cam, pose, d = _gputest_inputs()
    out.append(("gputest_icp", 640, 480, cam, pose, d, dict(icpWeight=100.0, so3=False, pyramid=False, if_curvature_info=False)))
    out.append(("gputest_faithful", 640, 480, cam, pose, d, dict(icpWeight=10.0, so3=True, pyramid=False, if_curvature_info=False)))
    return out


def run(make, up, case):
    """-> dict(trans, rot, counts[3] = ICP / RGB / SO3 counts of the last iteration)"""
    name, W, H, cam, pose, d, kw = case
    o = init_tracker(make(W, H, cam), up, pose, d)
    t, R, st = o.getIncrementalTransformation(pose[:3, 3], pose[:3, :3], **kw)
    g = (lambda k: st[k]) if isinstance(st, dict) else (lambda k: getattr(st, k))
    return dict(trans=np.asarray(t, np.float32), rot=np.asarray(R, np.float32), counts=np.array([g("lastICPCount"), g("lastRGBCount"), g("lastSO3Count")], np.float64), odom=o, stats=st)
