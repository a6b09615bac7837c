"""Runs the REFERENCE's own CUDA reduction kernels (oracle/_ref/libref_reduce.so, built by
oracle/build_ref.sh from /root/reference/Core/src/Cuda/reduce.cu) on the cases of tests/ref_cases.py
and writes their outputs as golden vectors.  Needs a GPU:

    gpurun -- 'python oracle/gen_ref_golden.py gpurun_out/ref_reduce.npz'
    cp gpurun_out/ref_reduce.npz tests/golden/ref_reduce.npz

The big per-pixel outputs (correspondences, DataTerm image) are stored as checksums + a strided sample."""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import orc_py, ref_py  # noqa: E402
from tests import ref_cases  # noqa: E402


def compact(out):
    """full-image outputs -> crc32 + counts (kept small enough to commit)"""
    o = {}
    for k, v in out.items():
        if k in ("icp_corres", "res_corr"):
            v = np.ascontiguousarray(v)
            o[k + "_crc"] = np.array([zlib.crc32(v.tobytes())], np.int64)
            if k == "icp_corres":
                o[k + "_nfound"] = np.array([(v[..., 0] >= 0).sum()], np.int64)
        else:
            o[k] = np.asarray(v)
    return o


def main(path):
    res = {}
    for name, c in ref_cases.cases(orc_py):
        out = compact(ref_cases.run_steps(ref_py, c, orc_py.DATATERM))
        for k, v in out.items():
            res[f"{name}/{k}"] = v
        print(name, "icp inliers", out["icp_res"][1], "rgb", out["res_sigma_count"], "so3", out["so3_res"])
    np.savez_compressed(path, **res)
    print("wrote", path, len(res), "arrays")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "ref_reduce.npz"))
