#!/usr/bin/env python
"""BASELINE configs 3 and 4 through the full pipeline (extra keys of the bench line; bench.py runs this in a subprocess):

  config3 : synthetic 640x480 ROOM sequence with a loop, the surfel map growing around the whole room:
            frames/s, surfel count and per-stage milliseconds (the reference's four Stopwatch spans) as the map grows
  config4 : 1280x960 stream, HRBF K = 16 neighbours, window 3 (bandwidth stress): frames/s of the full pipeline

    python scripts/bench_extra.py [--frames3 1000] [--frames4 120] [--only 3|4]      -> one JSON line per configuration
Frames are rendered on the host (process pool) BEFORE CUDA is touched; inputs are resident in HBM; one live sequence (384-thread
tracker), reference defaults; device time with CUDA events."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hrbffusion3d_b200 import synth  # noqa: E402


def ate(traj12, poses):
    P0inv = np.linalg.inv(np.asarray(poses[0], np.float64))
    gt = np.stack([(P0inv @ np.asarray(p, np.float64))[:3, 3] for p in poses[:traj12.shape[0]]])
    return float(np.sqrt(np.mean(np.sum((traj12[:, 9:12].astype(np.float64) - gt) ** 2, axis=1))))


def run(W, H, kind, poses, n_frames, name, capacity, **kw):
    t0 = time.perf_counter()
    fr = synth.render_sequence(kind, poses, W, H, synth.default_camera(W, H), seed0=3000 if W == 640 else 4000)
    t_render = time.perf_counter() - t0
    import torch
    from hrbffusion3d_b200.fusion import HRBFFusion
    cam = synth.default_camera(W, H)
    depth = torch.from_numpy(np.stack([f[0] for f in fr]).view(np.int16)).cuda()
    rgb = torch.from_numpy(np.stack([f[1] for f in fr])).cuda()
    n_in = len(fr)
    del fr
    F = HRBFFusion(W, H, cam, capacity=capacity, **kw)
    torch.cuda.set_stream(torch.cuda.Stream(priority=-1))   # above the library's staging streams (lowest priority), as in bench.py
    F.stageFrame(rgb[0], depth[0])
    F.processStaged(None)                                   # frame 1 initialises the map
    F.stageFrame(rgb[1 % n_in], depth[1 % n_in])
    torch.cuda.synchronize()
    checkpoints, stages = [], []
    mark = sorted(set([max(2, n_frames // 10), n_frames // 4, n_frames // 2, 3 * n_frames // 4, n_frames - 1]))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t_wall = time.perf_counter()
    for i in range(1, n_frames):
        F.processStaged(None)
        F.stageFrame(rgb[(i + 1) % n_in], depth[(i + 1) % n_in])
        if i in mark:                                        # (a count read-back synchronises: part of the measured time, 5 times per run)
            checkpoints.append({"frame": i + 1, "surfels": int(F.globalModel.lastCount())})
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t_wall
    ms = e0.elapsed_time(e1)
    tr = F.trajectory().cpu().numpy()
    # per-stage times of a few more frames at the final map size (timed frames synchronise, so they are outside the fps measurement)
    F.enableTimings(True)
    acc = np.zeros(4)
    k = 8
    pose_out = np.zeros(16, np.float32)
    for j in range(k):
        i = n_frames + j
        F.processStaged(None)                                # as in the timed loop: this frame, then the next frame's staging beside it ...
        F.stageFrame(rgb[(i + 1) % n_in], depth[(i + 1) % n_in])
        torch.cuda.synchronize()                             # ... and only then the read-back of the four stage spans
        acc += np.asarray(list(F.lastTimings().values()), np.float64)
    F.enableTimings(False)
    out = {"config": name, "width": W, "height": H, "frames": n_frames - 1, "frames_per_s": (n_frames - 1) / (ms * 1e-3), "ms_per_frame": ms / (n_frames - 1),
           "frames_per_s_host_wall_clock": (n_frames - 1) / wall, "surfels": checkpoints, "surfels_at_end": int(F.globalModel.lastCount()),
           "stage_ms_at_final_map": dict(zip(("Initialization (preprocessing runs ahead on the staging stream: not in this span)", "Registration", "Integration", "Prediction"), [float(x) / k for x in acc])),
           "trajectory_ate_rmse_m": ate(tr, [poses[i % len(poses)] for i in range(tr.shape[0])]) if n_frames <= len(poses) else None,
           "overflowed": bool(F.globalModel.overflowed()), "host_render_s": t_render, "params": kw,
           "timing": "CUDA events around frames 2..%d, inputs resident in HBM, one sequence, 384-thread tracker" % n_frames}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames3", type=int, default=1000)
    ap.add_argument("--frames4", type=int, default=120)
    ap.add_argument("--only", type=int, default=0)
    a = ap.parse_args()
    if a.only in (0, 3):
        run(640, 480, "room", synth.room_loop_trajectory(a.frames3), a.frames3, "BASELINE configs[2]: synthetic 640x480 room sequence with loop, full surfel map", 1 << 23)
    if a.only in (0, 4):
        n_loop = min(a.frames4, 60)
        run(1280, 960, "room", synth.circle_trajectory(n_loop, frames_per_rev=n_loop), a.frames4, "BASELINE configs[3]: 1280x960 stream, HRBF k=16, window 3", 1 << 23,
            predMaxNeighbors=16)


if __name__ == "__main__":
    main()
