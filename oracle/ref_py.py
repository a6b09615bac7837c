"""ctypes binding of oracle/_ref/libref_reduce.so: the REFERENCE's own CUDA reduction kernels
(Core/src/Cuda/reduce.cu, compiled unmodified by oracle/build_ref.sh) behind oracle/ref_shim.cu.

TEST INFRASTRUCTURE ONLY (needs a GPU): used by the `-m gpu` tests to pin the CPU oracle to the
real reference kernels and by oracle/gen_ref_golden.py to produce tests/golden/ref_reduce.npz.
All arrays are host numpy arrays; maps are dense SoA float32 [4*rows, cols]."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(_HERE, "_ref", "libref_reduce.so")
_LIB = None

# launch shapes the reference ships as defaults (Core/src/Utils/GPUConfig.h:53-60)
ICP_SHAPE = (128, 112)
RGB_SHAPE = (128, 112)
RGBRES_SHAPE = (256, 336)
SO3_SHAPE = (160, 64)


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


def icpStep(Rcurr, tcurr, vc, nc, k1c, k2c, Rprev_inv, tprev, cam, vg, ng, k1g, k2g, w,
            use_search=0, radius=2, use_weight=1, dist_thres=0.1, angle_thres=float(np.sin(np.deg2rad(20.0))),
            shape=ICP_SHAPE, want_corres=False):
    rows, cols = vc.shape[0] // 4, vc.shape[1]
    A, b, res = np.zeros(36, np.float32), np.zeros(6, np.float32), np.zeros(2, np.float32)
    corres = np.zeros((rows, cols, 2), np.int32) if want_corres else None
    lib().ref_icpStep(rows, cols, _p(_f(Rcurr)), _p(_f(tcurr)), _p(_f(vc)), _p(_f(nc)), _p(_f(k1c)), _p(_f(k2c)),
                      _p(_f(Rprev_inv)), _p(_f(tprev)), C.c_float(cam[0]), C.c_float(cam[1]), C.c_float(cam[2]), C.c_float(cam[3]),
                      _p(_f(vg)), _p(_f(ng)), _p(_f(k1g)), _p(_f(k2g)), _p(_f(w)), C.c_float(dist_thres), C.c_float(angle_thres),
                      int(use_search), int(radius), int(use_weight), shape[0], shape[1], _p(A), _p(b), _p(res),
                      _p(corres, C.c_int) if want_corres else None)
    return A.reshape(6, 6), b, res, corres


def computeRgbResidual(minScale, dIdx, dIdy, lastDepth, nextDepth, lastImage, nextImage, maxDepthDelta, kt, krkinv,
                       shape=RGBRES_SHAPE, dataterm_dtype=None):
    rows, cols = nextImage.shape
    corr = np.zeros((rows, cols, 16), np.uint8)
    sig, cnt = C.c_int(0), C.c_int(0)
    lib().ref_computeRgbResidual(rows, cols, C.c_float(minScale), _p(np.ascontiguousarray(dIdx), C.c_short), _p(np.ascontiguousarray(dIdy), C.c_short),
                                 _p(_f(lastDepth)), _p(_f(nextDepth)), _p(np.ascontiguousarray(lastImage), C.c_ubyte),
                                 _p(np.ascontiguousarray(nextImage), C.c_ubyte), corr.ctypes.data_as(C.c_void_p), C.c_float(maxDepthDelta),
                                 _p(_f(kt)), _p(_f(krkinv)), shape[0], shape[1], C.byref(sig), C.byref(cnt))
    if dataterm_dtype is not None:
        corr = corr.view(dataterm_dtype).reshape(rows, cols)
    return corr, sig.value, cnt.value


def rgbStep(corr, sigma, cloud3, fx, fy, dIdx, dIdy, use_grad_weight, sobelScale, shape=RGB_SHAPE):
    rows, cols = corr.shape[:2]
    A, b = np.zeros(36, np.float32), np.zeros(6, np.float32)
    corr = np.ascontiguousarray(corr)
    lib().ref_rgbStep(rows, cols, corr.ctypes.data_as(C.c_void_p), C.c_float(sigma), _p(_f(cloud3)), C.c_float(fx), C.c_float(fy),
                      _p(np.ascontiguousarray(dIdx), C.c_short), _p(np.ascontiguousarray(dIdy), C.c_short), int(use_grad_weight),
                      C.c_float(sobelScale), shape[0], shape[1], _p(A), _p(b))
    return A.reshape(6, 6), b


def so3Step(lastImage, nextImage, imageBasis, kinv, krlr, shape=SO3_SHAPE):
    rows, cols = nextImage.shape
    A, b, res = np.zeros(9, np.float32), np.zeros(3, np.float32), np.zeros(2, np.float32)
    lib().ref_so3Step(rows, cols, _p(np.ascontiguousarray(lastImage), C.c_ubyte), _p(np.ascontiguousarray(nextImage), C.c_ubyte),
                      _p(_f(imageBasis)), _p(_f(kinv)), _p(_f(krlr)), shape[0], shape[1], _p(A), _p(b), _p(res))
    return A.reshape(3, 3), b, res
