"""ctypes binding of oracle/_ref/libref_odometry.so: the REFERENCE's own RGBDOdometry class (tracking loop of
Core/src/Utils/RGBDOdometry.cpp compiled verbatim on the reference's own CUDA kernels; see oracle/build_ref_odometry.py for the
stand-ins: Eigen, GL textures).  TEST INFRASTRUCTURE ONLY; needs a GPU.  Same call names and argument meaning as
oracle.orc_py.Odometry, so that the same driver code runs on the oracle, on the CUDA library and on the reference."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def path(ieee=False):
    return os.path.join(_HERE, "_ref", "libref_odometry_ieee.so" if ieee else "libref_odometry.so")


def available(ieee=False):
    return os.path.exists(path(ieee))


def lib(ieee=False):
    if ieee not in _LIBS:
        L = C.CDLL(path(ieee))
        L.refodom_create.restype = C.c_void_p
        _LIBS[ieee] = L
    return _LIBS[ieee]


def _f(a):
    return np.ascontiguousarray(a, np.float32)


class Odometry:
    SLOTS = dict(vm=0, nm=1, vc=2, nc=3, k1m=4, k2m=5, k1c=6, k2c=7, w=8, rgbm=9, rgbc=10, first=11)

    def __init__(self, width, height, cx, cy, fx, fy, ieee=False):
        self.L = lib(ieee)
        self.w, self.h = width, height
        self.o = C.c_void_p(self.L.refodom_create(width, height, C.c_float(cx), C.c_float(cy), C.c_float(fx), C.c_float(fy)))
        if not self.o.value:
            raise RuntimeError("refodom_create failed (no CUDA device?)")
        self.last_us = None

    def __del__(self):
        try:
            if getattr(self, "o", None) and self.o.value:
                self.L.refodom_destroy(self.o)
                self.o = C.c_void_p()
        except Exception:
            pass

    def _tex(self, name, a, kind):
        a = np.ascontiguousarray(a, np.uint8 if kind == 2 else np.float32)
        assert self.L.refodom_set_texture(self.o, self.SLOTS[name], a.ctypes.data_as(C.c_void_p), kind) == 0
        return self.SLOTS[name]

    def _pose(self, pose):
        return _f(pose).ctypes.data_as(C.POINTER(C.c_float))

    def initICP(self, v, n, cutoff=20.0):
        self.L.refodom_initICP(self.o, self._tex("vc", v, 0), self._tex("nc", n, 0), C.c_float(cutoff))

    def initICPModel(self, v, n, cutoff, pose):
        self.L.refodom_initICPModel(self.o, self._tex("vm", v, 0), self._tex("nm", n, 0), C.c_float(cutoff), self._pose(pose))

    def initRGB(self, rgba):
        self.L.refodom_initRGB(self.o, self._tex("rgbc", rgba, 2))

    def initRGBModel(self, rgba):
        self.L.refodom_initRGBModel(self.o, self._tex("rgbm", rgba, 2))

    def initFirstRGB(self, rgba):
        self.L.refodom_initFirstRGB(self.o, self._tex("first", rgba, 2))

    def initCurvature(self, k1, k2):
        self.L.refodom_initCurvature(self.o, self._tex("k1c", k1, 0), self._tex("k2c", k2, 0))

    def initCurvatureModel(self, k1, k2, pose):
        self.L.refodom_initCurvatureModel(self.o, self._tex("k1m", k1, 0), self._tex("k2m", k2, 0), self._pose(pose))

    def initICPweight(self, w):
        self.L.refodom_initICPweight(self.o, self._tex("w", w, 1))

    def getIncrementalTransformation(self, trans, rot, rgbOnly=False, icpWeight=10.0, pyramid=True, fastOdom=False, so3=True, if_curvature_info=True):
        t = _f(trans).copy().reshape(3)
        R = _f(rot).copy().reshape(9)
        st = np.zeros(8, np.float32)
        rc = self.L.refodom_track(self.o, t.ctypes.data_as(C.POINTER(C.c_float)), R.ctypes.data_as(C.POINTER(C.c_float)), int(rgbOnly), C.c_float(icpWeight),
                                  int(pyramid), int(fastOdom), int(so3), int(if_curvature_info), st.ctypes.data_as(C.POINTER(C.c_float)))
        assert rc == 0, "CUDA error inside the reference's tracking call"
        self.last_us = float(st[6])
        stats = dict(lastICPError=float(st[0]), lastICPCount=float(st[1]), lastRGBError=float(st[2]), lastRGBCount=float(st[3]),
                     lastSO3Error=float(st[4]), lastSO3Count=float(st[5]), wall_us=float(st[6]))
        return t, R.reshape(3, 3), stats

    def lastSystem(self):
        A, b = np.zeros(36, np.float64), np.zeros(6, np.float64)
        self.L.refodom_last_system(self.o, A.ctypes.data_as(C.POINTER(C.c_double)), b.ctypes.data_as(C.POINTER(C.c_double)))
        return A.reshape(6, 6), b
