"""Run n frames of the bench workload through the pipeline (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from hrbffusion3d_b200.fusion import HRBFFusion
n = int(sys.argv[1]) if len(sys.argv) > 1 else 14
bench.RING = n
depth, rgb, poses, cam = bench.make_sequence(0, n)
F = HRBFFusion(bench.W, bench.H, cam, capacity=1 << 22)
for i in range(n):
    F.processFrame(rgb[i], depth[i])
print("surfels", F.globalModel.lastCount())
