"""Development: the actual CUDA-vs-oracle differences behind the loosest unit bounds (search-window ICP sums, k1 / k2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import util
from tests.test_gpu_odometry import build_both
from oracle import orc_py as orc
orc.lib().orc_set_float_loops(1)
from hrbffusion3d_b200 import odometry as od
oo, go, (m0, pose0, m1, pose1, cam) = build_both(orc, torch, 160, 120)
Rp, tp = pose0[:3, :3], pose0[:3, 3]
Rpi = np.linalg.inv(Rp).astype(np.float32)
names = ("vmap_curr", "nmap_curr", "ck1_curr", "ck2_curr")
gnames = ("vmap_g_prev", "nmap_g_prev", "ck1_g_prev", "ck2_g_prev", "icpWeight")
A, b, res, sums, cor = od.icpStep(Rp, tp, *[go.map(k, 0) for k in names], Rpi, tp, cam, *[go.map(k, 0) for k in gnames], use_search=True, search_radius=2)
Ao, bo, reso, sumso, coro = orc.icpStep(Rp, tp, *[oo.map(k, 0) for k in names], Rpi, tp, cam, *[oo.map(k, 0) for k in gnames], use_search=1, radius=2)
scale = np.abs(sumso[:27]).max()
print("search window: counts", res[1], reso[1], " max |d| / scale", np.abs(sums[:27] - sumso[:27]).max() / scale, " max rel", np.max(np.abs(sums[:27] - sumso[:27]) / np.maximum(np.abs(sumso[:27]), 1e-30)))
if cor is not None and coro is not None:
    cg = cor.cpu().numpy() if hasattr(cor, "cpu") else np.asarray(cor)
    print("  correspondences differing:", int((cg.reshape(-1, 2) != np.asarray(coro).reshape(-1, 2)).any(axis=1).sum()), "of", cg.size // 2)

from tests.test_gpu_fusion import _frames
from hrbffusion3d_b200.fusion import Frame, frame_params
for W, H in ((160, 120), (640, 480)):
    cam, poses, fr = _frames(W, H, 1)
    depth, rgb = fr[0]
    pp = orc.prep_params(cam, W, H)
    ref = orc.preprocess(pp, depth)
    f = Frame(frame_params(W, H, cam))
    f.upload(rgb, depth); f.preprocess()
    g = lambda n: f.tex(n).cpu().numpy()
    gk1, rk1 = g("PRINCIPAL_CURV1"), ref["curv1"]
    both = (np.abs(gk1[..., 3]) < 300) & (np.abs(rk1[..., 3]) < 300)
    for name, a, r in (("k1", gk1[..., 3], rk1[..., 3]), ("k2", g("PRINCIPAL_CURV2")[..., 3], ref["curv2"][..., 3]), ("gradient_mag", g("GRADIENT_MAG"), ref["gradient_mag"]),
                       ("normal_opt", g("NORMAL")[..., :3], ref["normal"][..., :3]), ("normal_pca", g("NORMAL_PCA")[..., :3], ref["normal_pca"][..., :3])):
        m = both if a.ndim == 2 else both
        d = np.abs(a[m] - r[m]).ravel(); s = np.abs(r[m]).ravel()
        rel = d / np.maximum(s, 1e-3)
        print(f"{W}x{H} {name:12s}: |d| max {d.max():.3e} p99.9 {np.quantile(d, 0.999):.3e} p99 {np.quantile(d, 0.99):.3e} | rel(max(|ref|,1e-3)) max {rel.max():.3e} p99.9 {np.quantile(rel, 0.999):.3e} p99 {np.quantile(rel, 0.99):.3e}")
