import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hrbffusion3d_b200 import synth
from hrbffusion3d_b200.fusion import Frame, frame_params
from oracle import orc_py as orc
for W, H in ((320, 240), (640, 480)):
    cam = synth.default_camera(W, H)
    sc = synth.Scene("room")
    depth, rgb = synth.render_depth(sc, synth.circle_trajectory(2, frames_per_rev=120)[1], W, H, cam, noise=True, seed=1)
    pp = orc.prep_params(cam, W, H)
    ref = orc.preprocess(pp, depth)
    f = Frame(frame_params(W, H, cam)); f.upload(rgb, depth); f.preprocess()
    for name, key in (("DEPTH_FILTERED", "filtered"), ("DEPTH_METRIC_FILTERED", "metric_filtered"), ("NORMAL_PCA", "normal_pca"), ("VERTEX_FILTERED", "vertex_filtered")):
        g = f.tex(name).cpu().numpy(); r = ref[key]
        neq = (g != r) & ~(np.isnan(g) & np.isnan(r))
        print(W, name, "unequal fraction %.3e" % neq.mean(), "max abs diff %.3e" % (np.abs(g - r)[neq].max() if neq.any() else 0))
