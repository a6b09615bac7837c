"""Runs the REFERENCE's own map / pyramid kernels (oracle/_ref/libref_cudafuncs.so, built by oracle/build_ref.sh from
/root/reference/Core/src/Cuda/cudafuncs.cu) on the cases of tests/ref5_cases.py and writes their outputs as golden vectors.
Needs a GPU:

    gpurun -- 'python oracle/gen_ref5_golden.py gpurun_out/ref_cudafuncs.npz'
    cp gpurun_out/ref_cudafuncs.npz tests/golden/ref_cudafuncs.npz

The 96x72 case is stored in full; of the 640x480 case every 5th pixel of every 5th row (+ the count of NaNs of the whole array).
It also prints how the CPU oracle compares, so a disagreement shows up before anything is committed."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import orc_py, ref5_py  # noqa: E402
from tests import ref5_cases  # noqa: E402


def sample(name, a, full):
    """what is stored of one output"""
    if full:
        return {name: a}
    planes = 4 if (a.ndim == 2 and name not in ("copy_w", "resize_w", "v2d", "v2d_cut", "pyr_gauss_f", "pyr_depth") and a.dtype.kind == "f") else 1
    rows = a.shape[0] // planes
    s = a.reshape((planes, rows) + a.shape[1:])[:, ::5, ::5]
    return {name: np.ascontiguousarray(s), name + "/nan": np.array([int(np.isnan(a).sum()) if a.dtype.kind == "f" else 0], np.int64)}


def main(path):
    res = {}
    for W, H in ref5_cases.SIZES:
        ref = ref5_cases.run_all(ref5_py, orc_py, W, H)
        orc = ref5_cases.run_all(orc_py, orc_py, W, H)
        for k, v in ref.items():
            try:
                ref5_cases.compare(k, orc[k], v)
                verdict = "oracle agrees"
            except AssertionError as e:
                verdict = "ORACLE DIFFERS: " + str(e).strip().split("\n")[0][:160]
            print(f"{W}x{H} {k:14s} {verdict}")
            for kk, vv in sample(k, v, full=(W, H) == ref5_cases.SIZES[0]).items():
                res[f"{W}x{H}/{kk}"] = vv
    np.savez_compressed(path, **res)
    print("wrote", path, len(res), "arrays")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "ref_cudafuncs.npz"))
