import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from tests.test_gpu_odometry import build_both
from oracle import orc_py as orc
W, H = int(sys.argv[1]), int(sys.argv[2])
oo, go, (m0, pose0, m1, pose1, cam) = build_both(orc, torch, W, H)
Rp, tp = pose0[:3, :3], pose0[:3, 3]
Rpi = np.linalg.inv(Rp).astype(np.float32)
for lvl in (0, 1, 2):
    print("level", lvl, flush=True)
    A, b, res, sums = go.icpStepLevel(lvl, Rp, tp, Rpi, tp, use_weight=True, tiled=True)
    print(res, flush=True)
