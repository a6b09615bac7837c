"""The reference's GPUTest protocol (GPUTest/src/GPUTest.cpp:42-129, 150-152, 247-278) on its own
fixture pair, restated for both the oracle and the CUDA path.

  K = (528, 528, 320, 240); model vertex/normal maps from 1d.png by forward differences with w = 1
  (loadVertices), current frame from 2d.png through initICP(depth) with the /5 -> mm conversion
  (loadDepth), identity start pose, depth cut-off 20 m.  Curvature / weight maps are never
  initialised by GPUTest; SURVEY 8c fixes them to "neutral" (curvature 0, use_weight off).
  Two variants: ICP-only (icpWeight 100, so3 off) and GPUTest-faithful (icpWeight 10, so3 on)."""
import numpy as np

K = (528.0, 528.0, 320.0, 240.0)  # fx, fy, cx, cy


def load_vertices(depth_u16):
    d = depth_u16
    H, W = d.shape
    fx, fy, cx, cy = [np.float32(k) for k in K]
    z = d.astype(np.float32) / np.float32(5000.0)
    col, row = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))
    V = np.zeros((H, W, 4), np.float32)
    V[..., 0] = (col - cx) * z * (np.float32(1.0) / fx)
    V[..., 1] = (row - cy) * z * (np.float32(1.0) / fy)
    V[..., 2] = z
    ok = np.zeros((H, W), bool)
    ok[1:-1, 1:-1] = (d[1:-1, 1:-1] > 0) & (d[2:, 1:-1] > 0) & (d[1:-1, 2:] > 0) & (d[:-2, 1:-1] > 0) & (d[1:-1, :-2] > 0)
    dx = np.zeros((H, W, 3), np.float32)
    dy = np.zeros((H, W, 3), np.float32)
    dx[:, :-1] = V[:, 1:, :3] - V[:, :-1, :3]
    dy[:-1] = V[1:, :, :3] - V[:-1, :, :3]
    n = np.cross(dx, dy).astype(np.float32)
    nn = np.sqrt(np.sum(n * n, -1, keepdims=True, dtype=np.float32))
    n = np.where(nn > 0, n / np.maximum(nn, np.float32(1e-30)), 0).astype(np.float32)
    N = np.zeros((H, W, 4), np.float32)
    N[..., :3] = n
    V[..., 3] = 1.0
    N[..., 3] = 1.0
    V[~ok] = (0, 0, 0, 1)
    N[~ok] = (0, 0, 0, 1)
    # border rows/cols are never written by loadVertices: defined as empty
    for a in (V, N):
        a[0] = a[-1] = 0
        a[:, 0] = a[:, -1] = 0
    return V, N


def load_depth_mm(depth_u16):
    return (depth_u16 // 5).astype(np.float32)


def rgba(rgb):
    return np.ascontiguousarray(np.concatenate([rgb, np.full(rgb.shape[:2] + (1,), 255, np.uint8)], -1))


def _run(make_odom, to_dev, g, first_step):
    V1, N1 = load_vertices(g["d1"])
    depth2 = load_depth_mm(g["d2"])
    I = np.eye(4, dtype=np.float32)
    out = {}
    for name, kw in (("icp_only", dict(icpWeight=100.0, so3=False)), ("faithful", dict(icpWeight=10.0, so3=True))):
        o = make_odom()
        o.initFirstRGB(to_dev(rgba(g["c1"])))
        o.initICPModel(to_dev(V1), to_dev(N1), 20.0, I)
        o.initRGBModel(to_dev(rgba(g["c1"])))
        # GPUTest calls initICP(depth); initRGB needs the vertex texture of the current frame
        V2, N2 = load_vertices(g["d2"])
        o.initICP(to_dev(V2), to_dev(N2), 20.0)
        o.initRGB(to_dev(rgba(g["c2"])))
        o.initICP_depth(to_dev(depth2), 20.0, 0.001)
        o.fillNeutralCurvature()
        if name == "icp_only":
            out["A0"], out["b0"], out["res0"] = first_step(o)
        t, R, st = o.getIncrementalTransformation(I[:3, 3], I[:3, :3], rgbOnly=False, pyramid=False, fastOdom=False,
                                                  if_curvature_info=False, **kw)
        out[name + "_trans"], out[name + "_rot"] = np.asarray(t), np.asarray(R)
        out[name + "_count"] = st.lastICPCount
    return out


def run_oracle(orc, g):
    def first_step(o):
        I3 = np.eye(3, dtype=np.float32)
        z3 = np.zeros(3, np.float32)
        A, b, res, _, _ = orc.icpStep(I3, z3, o.map(4, 0), o.map(5, 0), o.map(6, 0), o.map(7, 0), I3, z3, K,
                                      o.map(0, 0), o.map(1, 0), o.map(2, 0), o.map(3, 0), o.map(8, 0), use_weight=0)
        return A, b, res
    return _run(lambda: orc.Odometry(640, 480, K[2], K[3], K[0], K[1]), lambda a: a, g, first_step)


def run_cuda(g):
    import torch
    from hrbffusion3d_b200 import odometry as od

    def first_step(o):
        I3 = np.eye(3, dtype=np.float32)
        z3 = np.zeros(3, np.float32)
        A, b, res, _, _ = od.icpStep(I3, z3, o.map(4, 0), o.map(5, 0), o.map(6, 0), o.map(7, 0), I3, z3, K,
                                     o.map(0, 0), o.map(1, 0), o.map(2, 0), o.map(3, 0), o.map(8, 0), use_weight=False)
        return A, b, res
    return _run(lambda: od.RGBDOdometry(640, 480, K[2], K[3], K[0], K[1]), lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda(), g, first_step)
