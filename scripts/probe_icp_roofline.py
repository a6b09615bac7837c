"""ICP JTJ/JTr reduction launch at another image size (default 1280x960, BASELINE config 4): the same probe bench.py runs at 640x480
(hrbf_odometry_time_kernel: 200 back-to-back launches in one graph / 200 iterations inside the persistent tracker), on a synthetic
pair of that size.  Prints one JSON object.  bench.py runs it in a subprocess so that nothing here can cost the measured line."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def main(W, H):
    import torch
    from hrbffusion3d_b200 import odometry as od, synth
    from hrbffusion3d_b200._lib import check, lib, stream_ptr
    cam = synth.default_camera(W, H)
    sc = synth.Scene("room")
    pose0 = synth.make_pose(0.02, -0.03, 0.01, (0.05, -0.02, 0.0))
    delta = synth.make_pose(0.004, -0.006, 0.003, (0.008, -0.005, 0.006))
    pose1 = (pose0.astype(np.float64) @ delta.astype(np.float64)).astype(np.float32)
    m0 = synth.ideal_maps(sc, pose0, W, H, cam, seed=0)
    m1 = synth.ideal_maps(sc, pose1, W, H, cam, seed=1)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    d0 = {k: dev(v) for k, v in m0.items()}
    d1 = {k: dev(v) for k, v in m1.items()}
    go = od.RGBDOdometry(W, H, cam[2], cam[3], cam[0], cam[1])
    go.initFirstRGB(d0["rgba"])
    go.initICPModel(d0["vertex"], d0["normal"], 20.0, pose0)
    go.initRGBModel(d0["rgba"])
    go.initCurvatureModel(d0["k1"], d0["k2"], pose0)
    go.initICP(d1["vertex"], d1["normal"], 20.0)
    go.initRGB(d1["rgba"])
    go.initCurvature(d1["k1"], d1["k2"])
    go.initICPweight(d0["icpw"])
    go.getIncrementalTransformation(pose0[:3, 3], pose0[:3, :3], icpWeight=100.0, so3=False)      # sets the state the probes start from
    us, us_iter, us_cold = C.c_float(), C.c_float(), C.c_float()
    check(lib().hrbf_odometry_time_kernel(go._h, 0, 0, 0, 200, C.byref(us), stream_ptr()))
    check(lib().hrbf_odometry_time_kernel(go._h, 4, 0, 0, 200, C.byref(us_iter), stream_ptr()))
    check(lib().hrbf_odometry_time_kernel(go._h, 7, 0, 0, 30, C.byref(us_cold), stream_ptr()))
    torch.cuda.synchronize()
    alg = 68.0 * W * H
    print(json.dumps({"width": W, "height": H, "algorithmic_bytes_per_launch": alg, "us_per_launch": us.value,
                      "achieved": alg / (us.value * 1e-6) / 1e9, "us_per_launch_cold": us_cold.value, "achieved_cold": alg / (us_cold.value * 1e-6) / 1e9,
                      "us_per_iteration_in_tracker": us_iter.value,
                      "achieved_in_tracker": alg / (us_iter.value * 1e-6) / 1e9, "unit": "GB/s"}))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 2 else 1280, int(sys.argv[2]) if len(sys.argv) > 2 else 960)
