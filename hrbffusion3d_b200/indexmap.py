"""Host mirror of the reference's IndexMap (Core/src/IndexMap.h:36-201) over the C ABI.
GPUTexture accessors return CUDA tensors that alias the object's device buffers (no copies)."""
import ctypes as C

import numpy as np
import torch

from ._lib import check, lib, ptr, stream_ptr

_TEX = ["index", "vertConf", "colorTime", "normRad", "curvMax", "curvMin",
        "imageHRBF", "vertexHRBF", "normalHRBF", "curvk1HRBF", "curvk2HRBF", "timeHRBF", "icpweightHRBF",
        "oldImageHRBF", "oldVertexHRBF", "oldNormalHRBF", "oldcurvk1HRBF", "oldcurvk2HRBF", "oldTimeHRBF", "oldicpweightHRBF"]
_DT = {"index": (torch.int32, 1), "imageHRBF": (torch.uint8, 4), "timeHRBF": (torch.int16, 1), "icpweightHRBF": (torch.float32, 1),
       "oldImageHRBF": (torch.uint8, 4), "oldTimeHRBF": (torch.int16, 1), "oldicpweightHRBF": (torch.float32, 1)}


class _Alias:
    """__cuda_array_interface__ view of a raw device pointer"""

    def __init__(self, p, shape, typestr):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (p, False), "version": 3}


def alias_tensor(p, shape, dtype):
    typestr = {torch.float32: "<f4", torch.int32: "<i4", torch.uint8: "|u1", torch.int16: "<i2"}[dtype]
    return torch.as_tensor(_Alias(p, tuple(shape), typestr), device="cuda")


class IndexMap:
    ACTIVE, INACTIVE = 0, 1
    FACTOR = 1
    ACTIVE_KEYFRAME_DIMENSION = 19200

    def __init__(self, width, height, cx, cy, fx, fy):
        self.width, self.height = width, height
        self._h = C.c_void_p()
        L = lib()
        L.hrbf_indexmap_texture.restype = C.c_void_p
        check(L.hrbf_indexmap_create(C.byref(self._h), width, height, C.c_float(cx), C.c_float(cy), C.c_float(fx), C.c_float(fy)))
        self.lActiveKFID = [0]

    def close(self):
        if getattr(self, "_h", None) and self._h.value and lib is not None:
            lib().hrbf_indexmap_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def setActiveKeyframes(self, ids):
        self.lActiveKFID = list(ids)
        arr = (C.c_int * len(ids))(*ids)
        check(lib().hrbf_indexmap_set_active_keyframes(self._h, arr, len(ids), stream_ptr()))

    # IndexMap.cpp:193-267
    def predictIndices(self, pose, time, maxTime, model, depthCutoff, insertSubmap=0, indexSubmap=0):
        """model = (surfel tensor float32 [capacity, 20] on the GPU, count)"""
        surfels, count = model
        pose = np.ascontiguousarray(pose, np.float32)
        check(lib().hrbf_indexmap_predict_indices(self._h, pose.ctypes.data_as(C.POINTER(C.c_float)), int(time), int(maxTime), ptr(surfels),
                                                  C.c_uint(int(count)), C.c_float(depthCutoff), int(insertSubmap), int(indexSubmap), stream_ptr()))

    # IndexMap.cpp:413-518
    def predictHRBF(self, predictionType=0, win=3, minNeighbors=6, maxNeighbors=10, confThreshold=3.0, icpWeightLambda=10.0):
        check(lib().hrbf_indexmap_predict_hrbf(self._h, int(predictionType), int(win), int(minNeighbors), int(maxNeighbors),
                                               C.c_float(confThreshold), C.c_float(icpWeightLambda), stream_ptr()))

    def tex(self, name):
        which = _TEX.index(name)
        dt, ch = _DT.get(name, (torch.float32, 4))
        p = lib().hrbf_indexmap_texture(self._h, which)
        shape = (self.height, self.width, ch) if ch > 1 else (self.height, self.width)
        return alias_tensor(p, shape, dt)

    def texPtr(self, name):
        return C.c_void_p(lib().hrbf_indexmap_texture(self._h, _TEX.index(name)))

    # reference accessor names
    def indexTex(self): return self.tex("index")
    def vertConfTex(self): return self.tex("vertConf")
    def colorTimeTex(self): return self.tex("colorTime")
    def normalRadTex(self): return self.tex("normRad")
    def curvMaxTex(self): return self.tex("curvMax")
    def curvMinTex(self): return self.tex("curvMin")
    def imageTexHRBF(self): return self.tex("imageHRBF")
    def vertexTexHRBF(self): return self.tex("vertexHRBF")
    def normalTexHRBF(self): return self.tex("normalHRBF")
    def curvk1TexHRBF(self): return self.tex("curvk1HRBF")
    def curvk2TexHRBF(self): return self.tex("curvk2HRBF")
    def icpweightTexHRBF(self): return self.tex("icpweightHRBF")
