import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import orc_py as orc
from hrbffusion3d_b200 import synth
from hrbffusion3d_b200.fusion import Frame, frame_params
W,H=160,120
cam=synth.default_camera(W,H); sc=synth.Scene("room")
depth,rgb=synth.render_depth(sc,synth.make_pose(),W,H,cam,noise=True,seed=0)
for bil in (0,1):
    pp=orc.prep_params(cam,W,H,bilateral=bil)
    ref=orc.preprocess(pp,depth)
    f=Frame(frame_params(W,H,cam,bilateral=bil)); f.upload(rgb,depth); f.preprocess()
    g=lambda n: f.tex(n).cpu().numpy()
    mf=g("DEPTH_METRIC_FILTERED"); print("bil",bil,"metric_filtered max diff",np.abs(mf-ref["metric_filtered"]).max(), "nonequal frac",(mf!=ref["metric_filtered"]).mean())
    n=g("NORMAL_PCA"); d=np.abs(n[...,:3]-ref["normal_pca"][...,:3]).max(-1)
    print("  normal max diff",d.max(),"frac>1e-6",(d>1e-6).mean(),"frac>1e-4",(d>1e-4).mean(),"frac>1e-3",(d>1e-3).mean())
    y,x=np.unravel_index(np.argmax(d),d.shape); print("  worst at",y,x,n[y,x],ref["normal_pca"][y,x])
