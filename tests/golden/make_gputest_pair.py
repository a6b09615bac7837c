"""Generates tests/golden/gputest_pair.npz from the reference's own fixtures
(/root/reference/GPUTest/{1c,1d,2c,2d}.png -- image DATA, not source code) and records the CPU
oracle's outputs on them as the golden vector (the reference records no expected pose).
Run in the build container (needs /root/reference and cv2):  python tests/golden/make_gputest_pair.py"""
import hashlib
import os
import subprocess
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import orc_py  # noqa: E402
from tests.gputest_pair import run_oracle  # noqa: E402

REF = "/root/reference/GPUTest"
g = {}
sha = {}
for n in ("1c", "1d", "2c", "2d"):
    p = os.path.join(REF, n + ".png")
    sha[n] = hashlib.sha256(open(p, "rb").read()).hexdigest()[:8]
    im = cv2.imread(p, cv2.IMREAD_UNCHANGED)
    g[n[1] + n[0]] = im if im.ndim == 2 else np.ascontiguousarray(im[..., ::-1])  # BGR -> RGB
print("fixture sha256 prefixes:", sha)   # 0830aec0 a9097ea8 d5272c81 6b36a129
out = run_oracle(orc_py, g)
try:
    rev = subprocess.check_output(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"]).decode().strip()
except Exception:
    rev = "unknown"
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "gputest_pair.npz"), oracle_git=rev,
                    sha=np.array([sha[k] for k in ("1c", "1d", "2c", "2d")]), **g, **out)
for k, v in out.items():
    print(k, np.round(np.asarray(v), 6).tolist() if np.size(v) < 12 else "...")
