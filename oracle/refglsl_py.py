"""ctypes binding of oracle/_ref/libref_glsl.so: the REFERENCE's own GLSL shaders compiled for the CPU (oracle/build_ref_glsl.py,
oracle/glsl_cpu.h).  TEST INFRASTRUCTURE ONLY.  Every function has the name, arguments and return value of the oracle function it
is compared with (oracle/orc_py.py); framebuffer stores are the driver-side conversions of the attachment formats
(RGBA8 unorm, R16UI), stated here and nowhere in the shader."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(_HERE, "_ref", "libref_glsl.so")
_LIB = None


def available():
    """the library exists, or can be built now (the reference's shader sources are on this machine)"""
    if not os.path.exists(PATH) or os.path.isdir("/root/reference/Core/src/Shaders"):
        subprocess.call([sys.executable, os.path.join(_HERE, "build_ref_glsl.py")])
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


def _unorm8(x):
    q = x * np.float32(255.0)
    return np.where(q <= 0, 0, np.where(q >= 255.0, 255, (q + np.float32(0.5)).astype(np.int32))).astype(np.uint8)


def predictHRBF(idx, cam, width, height, win=3, minNeighbors=6, maxNeighbors=10, confThreshold=3.0, icpWeightLambda=10.0):
    """Shaders/predict_hrbf.frag over the whole image; idx: dict from predictIndices -> the dict orc_py.predictHRBF returns"""
    f4 = lambda: np.zeros((height, width, 4), np.float32)
    image, vertex, normal, k1, k2 = f4(), f4(), f4(), f4(), f4()
    time, icpw = np.zeros((height, width), np.uint32), np.zeros((height, width), np.float32)
    lib().glsl_predict_hrbf(width, height, _p(np.ascontiguousarray(idx["index"], np.uint32), C.c_uint), _p(_f(idx["vertConf"])), _p(_f(idx["colorTime"])),
                            _p(_f(idx["normRad"])), _p(_f(idx["curvMax"])), _p(_f(idx["curvMin"])),
                            C.c_float(cam[2]), C.c_float(cam[3]), C.c_float(cam[0]), C.c_float(cam[1]), C.c_float(1.0), C.c_float(win),
                            int(minNeighbors), int(maxNeighbors), C.c_float(icpWeightLambda), C.c_float(confThreshold),
                            _p(image), _p(vertex), _p(normal), _p(k1), _p(k2), _p(time, C.c_uint), _p(icpw))
    return {"image": _unorm8(image), "vertex": vertex, "normal": normal, "curvk1": k1, "curvk2": k2, "time": time.astype(np.uint16), "icpw": icpw}
