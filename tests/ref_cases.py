"""Deterministic input cases for the rows 1-3 step functions, shared by
  * oracle/gen_ref_golden.py  (runs the REFERENCE's own kernels on the GPU box -> tests/golden/ref_reduce.npz)
  * tests/test_oracle_vs_reference.py (CPU: oracle vs those golden vectors; GPU: oracle and CUDA path vs the
    reference kernels live).
Inputs are built with the CPU oracle's pyramid prep from the seeded synthetic pair and from the reference's
GPUTest fixture pair (tests/golden/gputest_pair.npz)."""
import os

import numpy as np

from hrbffusion3d_b200 import synth
from tests import gputest_pair as gp
from tests.util import pair

GOLD = os.path.join(os.path.dirname(__file__), "golden")
NAMES_C = ("vmap_curr", "nmap_curr", "ck1_curr", "ck2_curr")
NAMES_G = ("vmap_g_prev", "nmap_g_prev", "ck1_g_prev", "ck2_g_prev", "icpWeight")


def _synth_odom(orc, W, H):
    m0, pose0, m1, pose1, cam = pair(W, H)
    oo = orc.Odometry(W, H, cam[2], cam[3], cam[0], cam[1])
    oo.initFirstRGB(m0["rgba"])
    oo.initICPModel(m0["vertex"], m0["normal"], 20.0, pose0)
    oo.initRGBModel(m0["rgba"])
    oo.initCurvatureModel(m0["k1"], m0["k2"], pose0)
    oo.initICP(m1["vertex"], m1["normal"], 20.0)
    oo.initRGB(m1["rgba"])
    oo.initCurvature(m1["k1"], m1["k2"])
    oo.initICPweight(m0["icpw"])
    return oo, pose0, cam


def _gputest_odom(orc):
    g = np.load(os.path.join(GOLD, "gputest_pair.npz"))
    V1, N1 = gp.load_vertices(g["d1"])
    V2, N2 = gp.load_vertices(g["d2"])
    I = np.eye(4, dtype=np.float32)
    oo = orc.Odometry(640, 480, gp.K[2], gp.K[3], gp.K[0], gp.K[1])
    oo.initFirstRGB(gp.rgba(g["c1"]))
    oo.initICPModel(V1, N1, 20.0, I)
    oo.initRGBModel(gp.rgba(g["c1"]))
    oo.initICP(V2, N2, 20.0)
    oo.initRGB(gp.rgba(g["c2"]))
    oo.initICP_depth(gp.load_depth_mm(g["d2"]), 20.0, 0.001)
    oo.fillNeutralCurvature()
    return oo, I, gp.K


def cases(orc):
    """yields (name, dict) -- every array a host numpy array"""
    for tag, (oo, pose0, cam), levels, use_weight in (
            ("synth160", _synth_odom(orc, 160, 120), (0,), 1),
            ("synth320", _synth_odom(orc, 320, 240), (0, 1, 2), 1),
            ("gputest", _gputest_odom(orc), (0, 2), 0)):
        Rp, tp = np.ascontiguousarray(pose0[:3, :3]), np.ascontiguousarray(pose0[:3, 3])
        Rpi = np.linalg.inv(Rp).astype(np.float32)
        for lvl in levels:
            camL = tuple(np.float32(c) / np.float32(1 << lvl) for c in cam)
            c = dict(lvl=lvl, cam=camL, Rp=Rp, tp=tp, Rpi=Rpi, use_weight=use_weight)
            c["curr"] = [np.array(oo.map(k, lvl)) for k in NAMES_C]
            c["model"] = [np.array(oo.map(k, lvl)) for k in NAMES_G]
            c["nextI"], c["lastI"], c["lastNextI"] = np.array(oo.image(1, lvl)), np.array(oo.image(0, lvl)), np.array(oo.image(2, lvl))
            c["lastD"], c["nextD"] = np.array(oo.depth(0, lvl)), np.array(oo.depth(1, lvl))
            c["dx"], c["dy"] = orc.sobel(c["nextI"])
            Kl = np.array([[camL[0], 0, camL[2]], [0, camL[1], camL[3]], [0, 0, 1]], np.float64)
            c["krkinv"] = (Kl @ np.linalg.inv(Kl)).astype(np.float32)
            c["kt"] = (Kl @ np.array([0.002, -0.001, 0.003])).astype(np.float32)
            c["minScale"] = float([5, 3, 1][lvl] ** 2 / 0.125 ** 2)
            c["cloud"] = orc.projectToPointCloud(c["lastD"], camL)
            Rr = synth.rot_xyz(0.002, -0.003, 0.001)
            c["so3_B"] = (Kl @ Rr @ np.linalg.inv(Kl)).astype(np.float32)
            c["so3_kinv"] = np.linalg.inv(Kl).astype(np.float32)
            c["so3_krlr"] = (Kl @ Rr).astype(np.float32)
            yield f"{tag}_l{lvl}", c


def run_steps(impl, c, dataterm_dtype):
    """impl = oracle.orc_py or oracle.ref_py.  -> dict of flat outputs"""
    out = {}
    r = impl.icpStep(c["Rp"], c["tp"], *c["curr"], c["Rpi"], c["tp"], c["cam"], *c["model"], use_weight=c["use_weight"], want_corres=True)
    out["icp_A"], out["icp_b"], out["icp_res"], out["icp_corres"] = r[0], r[1], r[2], r[-1]
    r = impl.icpStep(c["Rp"], c["tp"], *c["curr"], c["Rpi"], c["tp"], c["cam"], *c["model"], use_search=1, radius=2, use_weight=c["use_weight"])
    out["icps_A"], out["icps_b"], out["icps_res"] = r[0], r[1], r[2]
    corr, sig, cnt = impl.computeRgbResidual(c["minScale"], c["dx"], c["dy"], c["lastD"], c["nextD"], c["lastI"], c["nextI"], 0.07, c["kt"], c["krkinv"])
    corr = np.ascontiguousarray(corr).view(np.uint8).reshape(corr.shape[0], corr.shape[1], 16).copy()
    corr[..., 13:] = 0                       # struct padding is undefined in the reference
    corr[corr[..., 12] == 0] = 0             # ... and so is every field of a DataTerm with valid == false (reduce.cu:994-996)
    out["res_sigma_count"] = np.array([sig, cnt], np.int64)
    out["res_corr"] = corr
    corr_dt = corr.view(dataterm_dtype).reshape(corr.shape[0], corr.shape[1])
    sigma = float(np.sqrt(max(cnt, 1)))
    r = impl.rgbStep(corr_dt, sigma, c["cloud"], c["cam"][0], c["cam"][1], c["dx"], c["dy"], 0, 0.125)
    out["rgb_A"], out["rgb_b"] = r[0], r[1]
    r = impl.so3Step(c["lastNextI"], c["nextI"], c["so3_B"], c["so3_kinv"], c["so3_krlr"])
    out["so3_A"], out["so3_b"], out["so3_res"] = r[0], r[1], r[2]
    return out
