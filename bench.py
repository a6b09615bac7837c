#!/usr/bin/env python
"""bench.py -- frames/sec of the HRBFFusion per-frame hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one frame through the hot path on the workload named in `config.workload`.
  value    : frames/s with the frame's inputs already resident in HBM (device timed, CUDA events)
  e2e      : frames/s through the reference-facing C ABI with HOST (pinned) input buffers: the H2D copy of
             the frame's inputs and the D2H read of the estimated pose are inside the timed region
  roofline : the ICP JTJ/JTr reduction kernel (level 0), algorithmic 68 B per pixel-iteration,
             timed live with CUDA events on its own stream (hrbf_odometry_time_kernel)
  cpu_baseline : the CPU oracle (oracle/, a restatement of the reference; kind "port") on the host cores,
             on a bounded sample of the same frames
N > 1 : one process per GPU (torchrun), one independent sequence per rank (weak scaling), NCCL only to
scatter the inputs' seeds and gather the trajectories; no collective inside the frame loop.
`--impl reference` times the oracle's CPU path (the reference itself needs OpenGL + Pangolin + Eigen and
cannot run headless) with all host threads on the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from hrbffusion3d_b200 import synth  # noqa: E402

W, H = 640, 480
RING = 8                     # distinct synthetic frames cycled through (inputs > L2, see config)
ICP_BYTES_PER_PIXEL_ITER = 68.0
TRACK_KW = dict(rgbOnly=False, icpWeight=10.0, pyramid=True, fastOdom=False, so3=True, if_curvature_info=True)


def make_frames(seed, n=RING):
    """n consecutive views of the planar scene: per frame the textures the
    tracking path consumes (the reference's RGBA32F vertex / normal / curvature maps, icp weight, RGBA8)."""
    sc = synth.Scene("plane")
    cam = synth.default_camera(W, H)
    poses = synth.circle_trajectory(n + 1)
    frames = [synth.ideal_maps(sc, p, W, H, cam, seed=seed * 1000 + i) for i, p in enumerate(poses)]
    return frames, poses, cam


class ClockSampler:
    FIELDS = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------- CPU arm
def run_oracle_frames(frames, poses, cam, n_frames, threads):
    """The oracle's tracking path (prep + getIncrementalTransformation) on n_frames frames -> seconds"""
    os.environ["OMP_NUM_THREADS"] = str(threads)
    from oracle import orc_py as orc
    oo = orc.Odometry(W, H, cam[2], cam[3], cam[0], cam[1])
    oo.initFirstRGB(frames[0]["rgba"])
    t0 = time.perf_counter()
    for i in range(n_frames):
        m0, m1, pose0 = frames[i % RING], frames[i % RING + 1], poses[i % RING]
        oo.initICPModel(m0["vertex"], m0["normal"], 20.0, pose0)
        oo.initRGBModel(m0["rgba"])
        oo.initCurvatureModel(m0["k1"], m0["k2"], pose0)
        oo.initICP(m1["vertex"], m1["normal"], 20.0)
        oo.initRGB(m1["rgba"])
        oo.initCurvature(m1["k1"], m1["k2"])
        oo.initICPweight(m0["icpw"])
        oo.getIncrementalTransformation(pose0[:3, 3], pose0[:3, :3], **TRACK_KW)
    return time.perf_counter() - t0


def config_dict(n_gpus):
    return {"workload": "synthetic 640x480 planar scene (SURVEY 8d config 2), tracking stage: pyramid prep (7 init* calls) + "
                        "getIncrementalTransformation, reference defaults (RGB+ICP weight 10, SO3 pre-align, iterations 10/5/4)",
            "width": W, "height": H, "frames_in_ring": RING,
            "l2": "inputs cycle through a ring of %d frames x 22 MB = %d MB > 126 MB L2" % (RING, RING * 22),
            "sequences": n_gpus, "parallelism": "one independent sequence per GPU"}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    frames, poses, cam = make_frames(0)
    run_oracle_frames(frames, poses, cam, max(1, args.warmup // 3 or 1), cores)
    per_step = []
    for _ in range(args.steps):
        per_step.append(run_oracle_frames(frames, poses, cam, 1, cores))
    t = float(np.sum(per_step))
    v = args.steps / t
    print(json.dumps({"impl": "reference", "metric": "frames/sec HRBF+ICP 640x480", "value": v, "unit": "frames/s", "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
                      "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(args.gpus),
                      "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": "port",
                                       "sample": "%d frames, one per step, OpenMP over %d threads" % (args.steps, cores)},
                      "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# --------------------------------------------------------------------------- GPU arm
def ours_arm(args):
    import torch
    import torch.distributed as dist
    from hrbffusion3d_b200 import odometry as od
    from hrbffusion3d_b200._lib import check, lib, stream_ptr

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # rank 0 scatters the per-sequence seeds (stand-in for the .klg byte ranges); trajectories are gathered at the end
    seed_t = torch.zeros(1, dtype=torch.int64, device="cuda")
    if world > 1:
        seeds = [torch.tensor([r], dtype=torch.int64, device="cuda") for r in range(world)] if rank == 0 else None
        dist.scatter(seed_t, seeds, src=0)
    frames, poses, cam = make_frames(int(seed_t.item()))

    KEYS = ("vertex", "normal", "k1", "k2", "icpw", "rgba")
    pinned = [{k: torch.from_numpy(np.ascontiguousarray(f[k])).pin_memory() for k in KEYS} for f in frames]
    dev = [{k: v.cuda() for k, v in f.items()} for f in pinned]
    stage = [{k: torch.empty_like(v, device="cuda") for k, v in pinned[0].items()} for _ in range(2)]
    h2d_bytes = sum(v.numel() * v.element_size() for v in pinned[0].values()) * 2      # model + current frame textures
    go = od.RGBDOdometry(W, H, cam[2], cam[3], cam[0], cam[1])
    go.initFirstRGB(dev[0]["rgba"])
    traj = torch.zeros((args.steps, 12), dtype=torch.float32)
    launches0 = lib().hrbf_launch_count()

    def step(i, host_inputs):
        a, b = i % RING, i % RING + 1
        if host_inputs:
            for k in KEYS:
                stage[0][k].copy_(pinned[a][k], non_blocking=True)
                stage[1][k].copy_(pinned[b][k], non_blocking=True)
            m0, m1 = stage
        else:
            m0, m1 = dev[a], dev[b]
        pose0 = poses[a]
        go.initICPModel(m0["vertex"], m0["normal"], 20.0, pose0)
        go.initRGBModel(m0["rgba"])
        go.initCurvatureModel(m0["k1"], m0["k2"], pose0)
        go.initICP(m1["vertex"], m1["normal"], 20.0)
        go.initRGB(m1["rgba"])
        go.initCurvature(m1["k1"], m1["k2"])
        go.initICPweight(m0["icpw"])
        t, R, _ = go.getIncrementalTransformation(pose0[:3, 3], pose0[:3, :3], **TRACK_KW)   # syncs; pose lands on the host
        return t, R

    def timed(host_inputs, record):
        for i in range(args.warmup):
            step(i, host_inputs)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            t, R = step(i, host_inputs)
            if record:
                traj[i, :9] = torch.from_numpy(np.asarray(R).reshape(9))
                traj[i, 9:] = torch.from_numpy(np.asarray(t))
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l_before = lib().hrbf_launch_count()
    ms_dev = timed(False, True)
    launches = lib().hrbf_launch_count() - l_before
    ms_e2e = timed(True, False)
    clocks = sampler.stop() if rank == 0 else None

    # roofline of the dominant kernel: level-0 ICP reduction, timed live with CUDA events
    us = C.c_float()
    check(lib().hrbf_odometry_time_kernel(go._h, 0, 0, 1, 200, C.byref(us), stream_ptr()))
    torch.cuda.synchronize()
    if world > 1:
        gathered = [torch.zeros_like(traj).cuda() for _ in range(world)] if rank == 0 else None
        dist.gather(traj.cuda(), gathered, dst=0)       # trajectories back to rank 0
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak_hbm()
    alg_bytes = ICP_BYTES_PER_PIXEL_ITER * W * H
    achieved = alg_bytes / (us.value * 1e-6) / 1e9
    cores = os.cpu_count() or 1
    n_cpu = 3
    run_oracle_frames(frames, poses, cam, 1, cores)
    cpu_s = run_oracle_frames(frames, poses, cam, n_cpu, cores)
    total_frames = args.steps * world
    out = {"metric": "frames/sec HRBF+ICP 640x480", "value": total_frames / (ms_dev * 1e-3), "unit": "frames/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(world),
           "e2e": {"value": total_frames / (ms_e2e * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 48 + 8 * 48},
           "gpu_launches": int(launches), "clocks": clocks,
           "roofline": {"bound": "hbm", "kernel": "icp_reduce_kernel<false> level 0 (640x480), incl. in-kernel Gauss-Newton solve",
                        "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                        "algorithmic_bytes_per_launch": alg_bytes, "us_per_launch": us.value, "traffic": None},
           "cpu_baseline": {"value": n_cpu / cpu_s, "unit": "frames/s", "cores": cores, "kind": "port",
                            "sample": "%d frames of the same ring (oracle, OpenMP over %d threads)" % (n_cpu, cores)}}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours_arm(args)


if __name__ == "__main__":
    main()
