"""Runs the REFERENCE's own tracking loop (oracle/_ref/libref_odometry.so and its IEEE-flag twin, oracle/build_ref_odometry.py) on the
cases of tests/ref4_cases.py and writes the poses as golden vectors; prints how the CPU oracle compares.  Needs a GPU:

    gpurun -- 'python oracle/gen_ref4_golden.py gpurun_out/ref_odometry.npz'  ;  cp gpurun_out/ref_odometry.npz tests/golden/"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import orc_py, refodom_py  # noqa: E402
from tests import ref4_cases  # noqa: E402
from tests.util import pose_err  # noqa: E402


def main(path):
    res = {}
    mk_orc = lambda W, H, cam: orc_py.Odometry(W, H, cam[2], cam[3], cam[0], cam[1])
    for case in ref4_cases.cases(orc_py):
        name = case[0]
        o = ref4_cases.run(mk_orc, lambda a: a, case)
        for build, ieee in (("asbuilt", False), ("ieee", True)):
            mk_ref = lambda W, H, cam: refodom_py.Odometry(W, H, cam[2], cam[3], cam[0], cam[1], ieee=ieee)
            r = ref4_cases.run(mk_ref, lambda a: a, case)
            ang, dt = pose_err(o["rot"], o["trans"], r["rot"], r["trans"])
            print(f"{name:22s} {build:8s}: oracle vs reference  ang {ang:.2e}  t {dt:.2e}   counts oracle {o['counts']} reference {r['counts']}   reference call {r['stats']['wall_us']:.0f} us")
            res[f"{build}/{name}/trans"], res[f"{build}/{name}/rot"], res[f"{build}/{name}/counts"] = r["trans"], r["rot"], r["counts"]
            A, b = r["odom"].lastSystem()
            res[f"{build}/{name}/lastA"], res[f"{build}/{name}/lastb"] = A, b
    np.savez_compressed(path, **res)
    print("wrote", path, len(res), "arrays")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "ref_odometry.npz"))
