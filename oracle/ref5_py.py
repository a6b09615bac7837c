"""ctypes binding of oracle/_ref/libref_cudafuncs.so: the REFERENCE's own map / pyramid kernels (Core/src/Cuda/cudafuncs.cu,
compiled unmodified by oracle/build_ref.sh with oracle/ref_texshim.h force-included) behind oracle/ref_shim_cudafuncs.cu.

TEST INFRASTRUCTURE ONLY (needs a GPU): pins row 5 of the CPU oracle to the real reference kernels.  Every function has the
name, arguments and return value of the oracle function it is compared with (oracle/orc_py.py)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(_HERE, "_ref", "libref_cudafuncs.so")
_LIB = None


def available():
    return os.path.exists(PATH)


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(PATH)
    return _LIB


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t=C.c_float):
    return a.ctypes.data_as(C.POINTER(t))


def _ok(rc):
    if rc != 0:
        raise RuntimeError(f"reference wrapper failed ({rc})")


def copyMaps(v_aos, n_aos):
    rows, cols = v_aos.shape[:2]
    v, n = np.empty((4 * rows, cols), np.float32), np.empty((4 * rows, cols), np.float32)
    _ok(lib().ref5_copyMaps(rows, cols, _p(_f(v_aos)), _p(_f(n_aos)), _p(v), _p(n)))
    return v, n


def copyCurvatureMap(c_aos, thr):
    rows, cols = c_aos.shape[:2]
    c = np.empty((4 * rows, cols), np.float32)
    _ok(lib().ref5_copyCurvatureMap(rows, cols, _p(_f(c_aos)), _p(c), C.c_float(thr)))
    return c


def copyicpWeightMap(w):
    rows, cols = w.shape
    o = np.empty((rows, cols), np.float32)
    _ok(lib().ref5_copyicpWeightMap(rows, cols, _p(_f(w)), _p(o)))
    return o


def resizeMap(src, normalize, init=None):
    rows, cols = src.shape[0] // 4, src.shape[1]
    dst = np.full((4 * (rows // 2), cols // 2), np.nan, np.float32) if init is None else init.copy()
    _ok(lib().ref5_resizeMap(rows // 2, cols // 2, _p(_f(src)), _p(dst), int(normalize)))
    return dst


def resizeCMap(src):
    rows, cols = src.shape[0] // 4, src.shape[1]
    dst = np.full((4 * (rows // 2), cols // 2), np.nan, np.float32)
    _ok(lib().ref5_resizeCMap(rows // 2, cols // 2, _p(_f(src)), _p(dst)))
    return dst


def resizeicpWeightMap(src):
    rows, cols = src.shape
    dst = np.zeros((rows // 2, cols // 2), np.float32)
    _ok(lib().ref5_resizeicpWeightMap(rows // 2, cols // 2, _p(_f(src)), _p(dst)))
    return dst


def tranformMaps(v, n, R, t):
    rows, cols = v.shape[0] // 4, v.shape[1]
    vd, nd = v.copy(), n.copy()
    _ok(lib().ref5_tranformMaps(rows, cols, _p(_f(v)), _p(_f(n)), _p(_f(R)), _p(_f(t)), _p(vd), _p(nd)))
    return vd, nd


def transformCurvMaps(k1, k2, R, t):
    rows, cols = k1.shape[0] // 4, k1.shape[1]
    a, b = k1.copy(), k2.copy()
    _ok(lib().ref5_transformCurvMaps(rows, cols, _p(_f(k1)), _p(_f(k2)), _p(_f(R)), _p(_f(t)), _p(a), _p(b)))
    return a, b


def pyrDownDepth(src):
    rows, cols = src.shape
    dst = np.zeros((rows // 2, cols // 2), np.float32)
    _ok(lib().ref5_pyrDownDepth(rows, cols, _p(_f(src)), _p(dst)))
    return dst


def createVMap(cam, depth, cutoff, factor):
    rows, cols = depth.shape
    v = np.zeros((4 * rows, cols), np.float32)
    _ok(lib().ref5_createVMap(C.c_float(cam[0]), C.c_float(cam[1]), C.c_float(cam[2]), C.c_float(cam[3]), rows, cols, _p(_f(depth)), _p(v),
                              C.c_float(cutoff), C.c_float(factor)))
    return v


def createNMap(vmap):
    rows, cols = vmap.shape[0] // 4, vmap.shape[1]
    n = np.zeros((4 * rows, cols), np.float32)
    _ok(lib().ref5_createNMap(rows, cols, _p(_f(vmap)), _p(n)))
    return n


def verticesToDepth(v_aos, cutoff):
    rows, cols = v_aos.shape[:2]
    d = np.zeros((rows, cols), np.float32)
    _ok(lib().ref5_verticesToDepth(rows, cols, _p(_f(v_aos)), _p(d), C.c_float(cutoff)))
    return d


def pyrDownGaussF(src):
    rows, cols = src.shape
    dst = np.zeros((rows // 2, cols // 2), np.float32)
    _ok(lib().ref5_pyrDownGaussF(rows, cols, _p(_f(src)), _p(dst)))
    return dst


def pyrDownUcharGauss(src):
    rows, cols = src.shape
    dst = np.zeros((rows // 2, cols // 2), np.uint8)
    _ok(lib().ref5_pyrDownUcharGauss(rows, cols, _p(np.ascontiguousarray(src, np.uint8), C.c_ubyte), _p(dst, C.c_ubyte)))
    return dst


def rgbaToIntensity(rgba):
    rows, cols = rgba.shape[:2]
    dst = np.zeros((rows, cols), np.uint8)
    _ok(lib().ref5_rgbaToIntensity(rows, cols, _p(np.ascontiguousarray(rgba, np.uint8), C.c_ubyte), _p(dst, C.c_ubyte)))
    return dst


def sobel(img):
    rows, cols = img.shape
    dx, dy = np.zeros((rows, cols), np.int16), np.zeros((rows, cols), np.int16)
    _ok(lib().ref5_sobel(rows, cols, _p(np.ascontiguousarray(img, np.uint8), C.c_ubyte), _p(dx, C.c_short), _p(dy, C.c_short)))
    return dx, dy


def projectToPointCloud(depth, cam_level):
    rows, cols = depth.shape
    cl = np.zeros((rows, cols, 3), np.float32)
    _ok(lib().ref5_projectToPointCloud(rows, cols, _p(_f(depth)), _p(cl), C.c_float(cam_level[0]), C.c_float(cam_level[1]),
                                       C.c_float(cam_level[2]), C.c_float(cam_level[3]), 0))
    return cl
