"""Aggregate frames/s of S independent sequences on ONE GPU, each on its own stream (development: is the GPU under-used by a
single latency-bound sequence?).  Usage: python scripts/dev_multiseq.py S [S ...]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from hrbffusion3d_b200.fusion import HRBFFusion
n = 48
STAGES = int(os.environ.get('STAGES', '0'))
TT = int(os.environ.get('TT', '256'))      # tracker threads per CTA
depth, rgb, poses, cam = bench.make_sequence(0, bench.RING, only=n)
d = torch.from_numpy(depth.view(np.int16)).cuda(); c = torch.from_numpy(rgb).cuda()


def run(S, steps=150, warm=20):
    Fs = [HRBFFusion(bench.W, bench.H, cam, capacity=1 << 21, trackerThreads=TT) for _ in range(S)]
    st = [torch.cuda.Stream() for _ in range(S)]
    off = [(s * n) // S for s in range(S)]

    stages = []
    pose = np.zeros(16, np.float32)
    if STAGES: Fs[0].enableTimings(True)

    def step(i):
        for s in range(S):
            with torch.cuda.stream(st[s]):
                Fs[s].stageFrame(c[(i + 1 + off[s]) % n], d[(i + 1 + off[s]) % n])
                if STAGES and s == 0:      # sequence 0 is read back every frame: its per-stage times under the others' load
                    Fs[0].processStaged(pose); stages.append(list(Fs[0].lastTimings().values()))
                else:
                    Fs[s].processStaged(None)
    for s in range(S):
        with torch.cuda.stream(st[s]):
            Fs[s].stageFrame(c[off[s]], d[off[s]])
    for i in range(warm): step(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(warm, warm + steps): step(i)
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    t = time.perf_counter() - t0
    cnt = [F.globalModel.lastCount() for F in Fs]
    if STAGES:
        m = np.median(np.array(stages[warm:]), axis=0) * 1e3
        print("   sequence 0 stage medians (us): preprocess-wait %.0f  registration %.0f  integration %.0f  prediction %.0f  (sum %.0f)" % (*m, m.sum()))
    return S * steps / t, t_host / (S * steps) * 1e6, cnt


for S in [int(a) for a in sys.argv[1:]] or [1, 2]:
    fps, host_us, cnt = run(S)
    print("lib=%s TT=%d S=%d: %.0f frames/s aggregate (%.1f us per frame; host enqueue %.1f us per frame) surfels %s" % (
        os.path.basename(os.environ.get("HRBF_B200_LIB", "default")), TT, S, fps, 1e6 / fps, host_us, cnt), flush=True)
