"""Development: the ICP reduction launch, per-pixel-gather form vs TMA-staged tile form, L2-hot (back-to-back launches in one graph)
and cold (L2 flushed before every launch, each launch timed alone), at the three pyramid levels of 640x480 and 1280x960."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_gpu_odometry import build_both
from oracle import orc_py as orc
for W, H in ((640, 480), (1280, 960)):
    oo, go, d = build_both(orc, torch, W, H)
    go.getIncrementalTransformation(d[1][:3, 3], d[1][:3, :3], icpWeight=100.0, so3=False)
    for lvl in (0, 1, 2):
        n = (W >> lvl) * (H >> lvl)
        r = [go.timeKernel(w, lvl, 0, 200 if w in (0, 5) else 30) for w in (0, 5, 7, 6)]
        f = lambda us: f"{us:6.2f} us {68 * n / us * 1e-3:5.0f} GB/s"
        print(f"{W}x{H} L{lvl}: gather hot {f(r[0])} | tile hot {f(r[1])} | gather cold {f(r[2])} | tile cold {f(r[3])}")
