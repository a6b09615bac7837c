"""Development: where a single replayed sequence's frame time goes -- the same frames (a) fully serial on one stream (processFrameDev),
(b) staged (stageFrame one frame ahead + processStaged), per tracker shape; and the reference's four stage spans of the serial form."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hrbffusion3d_b200 import synth
from hrbffusion3d_b200.fusion import HRBFFusion

W, H, RING, N = 640, 480, 24, 120
cam = synth.default_camera(W, H)
poses = synth.circle_trajectory(RING, frames_per_rev=RING)
fr = synth.render_sequence("plane", poses, W, H, cam, seed0=100)
depth = torch.from_numpy(np.stack([f[0] for f in fr]).view(np.int16)).cuda()
rgb = torch.from_numpy(np.stack([f[1] for f in fr])).cuda()

def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)

torch.cuda.set_stream(torch.cuda.Stream(priority=-1))      # above the library's staging streams
for tt in (512, 384, 256):
    F = HRBFFusion(W, H, cam, capacity=1 << 22, trackerThreads=tt)
    for i in range(8): F.processFrameDev(rgb[i % RING], depth[i % RING])
    ms = timed(lambda: [F.processFrameDev(rgb[i % RING], depth[i % RING]) for i in range(8, 8 + N)])
    print(f"TT={tt} serial : {ms / N * 1e3:7.1f} us/frame  {N / ms * 1e3:7.1f} fps")
    F.enableTimings(True)
    acc = np.zeros(4)
    F.stageFrame(rgb[(8 + N) % RING], depth[(8 + N) % RING])
    for i in range(8 + N, 8 + N + 20):          # pipelined as in the timed loop (this frame, then the next frame's staging beside it), read back after a sync
        F.processStaged(None); F.stageFrame(rgb[(i + 1) % RING], depth[(i + 1) % RING]); torch.cuda.synchronize()
        acc += np.array(list(F.lastTimings().values()))
    F.processStaged(None); torch.cuda.synchronize()
    print("         stage spans of the dependent chain, next frame's staging beside it (us):", np.round(acc / 20 * 1e3, 1), "(Initialization (staged: empty), Registration, Integration, Prediction)")
    del F
    F = HRBFFusion(W, H, cam, capacity=1 << 22, trackerThreads=tt)
    F.stageFrame(rgb[0], depth[0])
    for i in range(8):
        F.stageFrame(rgb[(i + 1) % RING], depth[(i + 1) % RING]); F.processStaged(None)
    def go():
        for i in range(8, 8 + N):
            F.stageFrame(rgb[(i + 1) % RING], depth[(i + 1) % RING]); F.processStaged(None)
    ms = timed(go)
    print(f"TT={tt} staged : {ms / N * 1e3:7.1f} us/frame  {N / ms * 1e3:7.1f} fps")
    F.processStaged(None); torch.cuda.synchronize()
    del F
