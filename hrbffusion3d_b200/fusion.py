"""Host mirrors of the reference's per-frame classes over the C ABI (include/hrbf_b200.h):
Frame (HRBFFusion's textures[] + preprocessing ComputePacks), FillIn (Shaders/FillIn.h), GlobalModel
(GlobalModel.h) and the HRBFFusion orchestrator (HRBFFusion.cpp:991-1260).  Textures are CUDA tensors
aliasing the objects' device buffers."""
import ctypes as C

import numpy as np
import torch

from ._lib import check, lib, ptr, stream_ptr
from .indexmap import IndexMap, alias_tensor


class FrameParams(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("cx", C.c_float), ("cy", C.c_float), ("fx", C.c_float), ("fy", C.c_float),
                ("depthFactor", C.c_float), ("depthCutoff", C.c_float), ("radiusMultiplier", C.c_float), ("normalPCA", C.c_int),
                ("curvWindow", C.c_int), ("bilateral", C.c_int), ("useConfEval", C.c_int), ("confEvalEpsilon", C.c_float)]


class FusionParams(C.Structure):
    _fields_ = [("frame", FrameParams), ("confidenceThreshold", C.c_float), ("maxDepthProcessed", C.c_float), ("icpWeight", C.c_float),
                ("rgbOnly", C.c_int), ("pyramid", C.c_int), ("fastOdom", C.c_int), ("so3", C.c_int), ("weightedICP", C.c_int),
                ("predWindow", C.c_int), ("predMinNeighbors", C.c_int), ("predMaxNeighbors", C.c_int), ("predConfThreshold", C.c_float),
                ("icpWeightLambda", C.c_float), ("curvValidThreshold", C.c_float), ("denseEnoughThresh", C.c_float), ("cleanWindow", C.c_int),
                ("capacity", C.c_uint), ("trackerThreads", C.c_int)]


_FT = ["RGB", "RGBA", "DEPTH_RAW", "DEPTH_FILTERED", "DEPTH_METRIC", "DEPTH_METRIC_FILTERED", "VERTEX_RAW", "VERTEX_FILTERED", "NORMAL_PCA",
       "NORMAL", "PRINCIPAL_CURV1", "PRINCIPAL_CURV2", "GRADIENT_MAG", "RADIUS", "CONFIDENCE"]
_FT_DT = {"RGB": (torch.uint8, 3), "RGBA": (torch.uint8, 4), "DEPTH_RAW": (torch.int16, 1), "DEPTH_FILTERED": (torch.float32, 1),
          "DEPTH_METRIC": (torch.float32, 1), "DEPTH_METRIC_FILTERED": (torch.float32, 1), "GRADIENT_MAG": (torch.float32, 1),
          "RADIUS": (torch.float32, 1), "CONFIDENCE": (torch.float32, 1)}
_FILL = ["image", "vertex", "normal", "curvk1", "curvk2", "icpweight"]


def _setup():
    L = lib()
    for n in ("hrbf_indexmap_texture", "hrbf_frame_texture", "hrbf_fillin_texture", "hrbf_model_model", "hrbf_model_count_dev", "hrbf_fusion_trajectory_dev",
              "hrbf_fusion_odometry", "hrbf_fusion_indexmap", "hrbf_fusion_model", "hrbf_fusion_frame", "hrbf_fusion_fillin"):
        getattr(L, n).restype = C.c_void_p
    return L


def frame_params(width, height, cam, depthFactor=1.0 / 5000.0, depthCutoff=3.5, radiusMultiplier=4.0, normalPCA=1, curvWindow=3,
                 bilateral=1, useConfEval=0, confEvalEpsilon=1000.0):
    fx, fy, cx, cy = cam
    return FrameParams(width, height, cx, cy, fx, fy, depthFactor, depthCutoff, radiusMultiplier, normalPCA, curvWindow, bilateral, useConfEval, confEvalEpsilon)


class _TexOwner:
    _names, _dt, _getter = [], {}, None

    def tex(self, name):
        dt, ch = self._dt.get(name, (torch.float32, 4))
        p = getattr(lib(), self._getter)(self._h, self._names.index(name))
        shape = (self.height, self.width, ch) if ch > 1 else (self.height, self.width)
        return alias_tensor(p, shape, dt)

    def texPtr(self, name):
        return C.c_void_p(getattr(lib(), self._getter)(self._h, self._names.index(name)))


class Frame(_TexOwner):
    _names, _dt, _getter = _FT, _FT_DT, "hrbf_frame_texture"

    def __init__(self, params, handle=None):
        _setup()
        self.width, self.height = params.width, params.height
        self._own = handle is None
        self._h = C.c_void_p(handle) if handle else C.c_void_p()
        if self._own:
            check(lib().hrbf_frame_create(C.byref(self._h), C.byref(params)))

    def __del__(self):
        if getattr(self, "_own", False) and self._h.value:
            try:
                lib().hrbf_frame_destroy(self._h)
            except Exception:
                pass
            self._h = C.c_void_p()

    def upload(self, rgb, depth):
        """rgb uint8 [h,w,3], depth uint16 [h,w]: numpy (host) or CUDA tensors"""
        host = isinstance(rgb, np.ndarray)
        if host:
            rgb = np.ascontiguousarray(rgb, np.uint8); depth = np.ascontiguousarray(depth, np.uint16)
            check(lib().hrbf_frame_upload(self._h, rgb.ctypes.data_as(C.c_void_p), depth.ctypes.data_as(C.c_void_p), 1, stream_ptr()))
            torch.cuda.current_stream().synchronize()      # pageable source
        else:
            check(lib().hrbf_frame_upload(self._h, ptr(rgb), ptr(depth), 0, stream_ptr()))

    def preprocess(self):
        check(lib().hrbf_frame_preprocess(self._h, stream_ptr()))

    def vertexConfidence(self, weighting):
        check(lib().hrbf_frame_vertex_confidence(self._h, C.c_float(weighting), stream_ptr()))


class FillIn(_TexOwner):
    _names, _dt, _getter = _FILL, {"image": (torch.uint8, 4), "icpweight": (torch.float32, 1)}, "hrbf_fillin_texture"

    def __init__(self, width, height, handle=None):
        _setup()
        self.width, self.height = width, height
        self._own = handle is None
        self._h = C.c_void_p(handle) if handle else C.c_void_p()
        if self._own:
            check(lib().hrbf_fillin_create(C.byref(self._h), width, height))

    def __del__(self):
        if getattr(self, "_own", False) and self._h.value:
            try:
                lib().hrbf_fillin_destroy(self._h)
            except Exception:
                pass
            self._h = C.c_void_p()

    def run(self, indexMap, frame, passthrough=False, lamb=10.0, curvThr=300.0):
        check(lib().hrbf_fillin_run(self._h, indexMap._h, frame._h, int(passthrough), C.c_float(lamb), C.c_float(curvThr), stream_ptr()))


def _hp(a):
    return np.ascontiguousarray(a, np.float32).ctypes.data_as(C.POINTER(C.c_float))


class GlobalModel:
    def __init__(self, width, height, cam, capacity=1 << 20, handle=None):
        _setup()
        fx, fy, cx, cy = cam
        self.width, self.height = width, height
        self._own = handle is None
        self._h = C.c_void_p(handle) if handle else C.c_void_p()
        if self._own:
            check(lib().hrbf_model_create(C.byref(self._h), width, height, C.c_float(cx), C.c_float(cy), C.c_float(fx), C.c_float(fy), C.c_uint(capacity)))

    def __del__(self):
        if getattr(self, "_own", False) and self._h.value:
            try:
                lib().hrbf_model_destroy(self._h)
            except Exception:
                pass
            self._h = C.c_void_p()

    def setParams(self, radiusMultiplier=4.0, curvValidThreshold=300.0, normalPCA=1, cleanWindow=2, useConfEval=0, confEvalEpsilon=1000.0):
        check(lib().hrbf_model_set_params(self._h, C.c_float(radiusMultiplier), C.c_float(curvValidThreshold), normalPCA, cleanWindow, useConfEval, C.c_float(confEvalEpsilon)))

    def initialise(self, vertexMap, normalMap, colorMap, curv1Map, curv2Map, gradientMagMap, init_pose):
        check(lib().hrbf_model_initialise(self._h, ptr(vertexMap), ptr(normalMap), ptr(colorMap), ptr(curv1Map), ptr(curv2Map), ptr(gradientMagMap),
                                          _hp(init_pose), stream_ptr()))

    def fuse(self, pose, time, rgb, depthRaw, depthFiltered, curv1, curv2, confidence, indexMap, vertConfMap, colorTimeMap, normRadMap,
             depthCutoff=20.0, confThreshold=5.0, weighting=1.0, insertSubmap=0, indexSubmap=0):
        check(lib().hrbf_model_fuse(self._h, _hp(pose), int(time), ptr(rgb), ptr(depthRaw), ptr(depthFiltered), ptr(curv1), ptr(curv2), ptr(confidence),
                                    ptr(indexMap), ptr(vertConfMap), ptr(colorTimeMap), ptr(normRadMap), C.c_float(depthCutoff), C.c_float(confThreshold),
                                    C.c_float(weighting), int(insertSubmap), int(indexSubmap), stream_ptr()))

    def clean(self, pose, time, indexMap, vertConfMap, colorTimeMap, normRadMap, depthMap=None, confThreshold=5.0, maxDepth=20.0):
        check(lib().hrbf_model_clean(self._h, _hp(pose), int(time), ptr(indexMap), ptr(vertConfMap), ptr(colorTimeMap), ptr(normRadMap), ptr(depthMap),
                                     C.c_float(confThreshold), C.c_float(maxDepth), stream_ptr()))

    def setModel(self, surfels):
        """surfels: float32 [n, 20] numpy array or CUDA tensor"""
        if isinstance(surfels, np.ndarray):
            a = np.ascontiguousarray(surfels, np.float32)
            check(lib().hrbf_model_set_model(self._h, a.ctypes.data_as(C.c_void_p), C.c_uint(a.shape[0]), 1, stream_ptr()))
        else:
            check(lib().hrbf_model_set_model(self._h, ptr(surfels.contiguous()), C.c_uint(surfels.shape[0]), 0, stream_ptr()))

    def updateModel(self, deltaTransformKF):
        """GlobalModel::updateModel: deltaTransformKF [n, 4, 4] rigid corrections indexed by sub-map id"""
        d = np.ascontiguousarray(deltaTransformKF, np.float32).reshape(-1, 16)
        check(lib().hrbf_model_update_model(self._h, d.ctypes.data_as(C.POINTER(C.c_float)), int(d.shape[0]), stream_ptr()))

    def lastCount(self):
        c = C.c_uint(0)
        check(lib().hrbf_model_last_count(self._h, C.byref(c), stream_ptr()))
        return int(c.value)

    def model(self):
        """(surfel tensor [count, 20] aliasing the current VBO, count)"""
        n = self.lastCount()
        p = lib().hrbf_model_model(self._h)
        return alias_tensor(p, (max(n, 1), 20), torch.float32)[:n], n

    def downloadMap(self):
        """GlobalModel::downloadMap: numpy float32 [count, 20]"""
        n = C.c_uint(0)
        check(lib().hrbf_model_download_map(self._h, None, 0, C.byref(n), stream_ptr()))
        out = np.zeros((n.value, 20), np.float32)
        if n.value:
            check(lib().hrbf_model_download_map(self._h, out.ctypes.data_as(C.POINTER(C.c_float)), n.value, C.byref(n), stream_ptr()))
        return out

    def exportPly(self, confThreshold):
        """bytes of the PLY file HRBFFusion::savePly writes (header + 43-byte vertices, packed on the device)"""
        L = lib()
        L.hrbf_ply_header.restype = C.c_size_t
        n = C.c_uint(0)
        check(L.hrbf_model_export_ply(self._h, C.c_float(confThreshold), None, C.c_size_t(0), C.byref(n), stream_ptr()))
        rec = np.zeros(n.value * 43, np.uint8)
        if n.value:
            check(L.hrbf_model_export_ply(self._h, C.c_float(confThreshold), rec.ctypes.data_as(C.c_void_p), C.c_size_t(rec.size), C.byref(n), stream_ptr()))
        buf = C.create_string_buffer(512)
        k = L.hrbf_ply_header(n.value, buf, C.c_size_t(512))
        if k == 0:
            raise RuntimeError("hrbf_ply_header: buffer too small")
        return buf.raw[:k] + rec.tobytes()

    def overflowed(self):
        f = C.c_int(0)
        check(lib().hrbf_model_overflowed(self._h, C.byref(f), stream_ptr()))
        return bool(f.value)


class HRBFFusion:
    """HRBFFusion::processFrame / predict with the sparse back-end off."""

    def __init__(self, width, height, cam, capacity=1 << 21, **kw):
        L = _setup()
        fx, fy, cx, cy = cam
        self.width, self.height, self.cam = width, height, cam
        self.params = FusionParams()
        L.hrbf_fusion_default_params(C.byref(self.params), width, height, C.c_float(cx), C.c_float(cy), C.c_float(fx), C.c_float(fy))
        self.params.capacity = capacity
        for k, v in kw.items():
            if hasattr(self.params.frame, k) and k not in ("width", "height"):
                setattr(self.params.frame, k, v)
            elif hasattr(self.params, k):
                setattr(self.params, k, v)
            else:
                raise TypeError(f"unknown parameter {k}")
        self._h = C.c_void_p()
        check(L.hrbf_fusion_create(C.byref(self._h), C.byref(self.params)))
        self.frame = Frame(self.params.frame, handle=L.hrbf_fusion_frame(self._h))
        self.fillIn = FillIn(width, height, handle=L.hrbf_fusion_fillin(self._h))
        self.globalModel = GlobalModel(width, height, cam, handle=L.hrbf_fusion_model(self._h))
        self.indexMap = IndexMap.__new__(IndexMap)
        self.indexMap.width, self.indexMap.height, self.indexMap._h, self.indexMap.lActiveKFID = width, height, C.c_void_p(L.hrbf_fusion_indexmap(self._h)), [0]
        self.indexMap.close = lambda: None

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            try:
                lib().hrbf_fusion_destroy(self._h)
            except Exception:
                pass
            self._h = C.c_void_p()

    @property
    def tick(self):
        return lib().hrbf_fusion_tick(self._h)

    def processFrame(self, rgb, depth, timestamp=0, weightMultiplier=1.0):
        """host numpy inputs (pinned or pageable); returns the 4x4 pose (synchronous, like the reference)"""
        rgb = np.ascontiguousarray(rgb, np.uint8); depth = np.ascontiguousarray(depth, np.uint16)
        pose = np.zeros(16, np.float32)
        check(lib().hrbf_fusion_process_frame(self._h, rgb.ctypes.data_as(C.c_void_p), depth.ctypes.data_as(C.c_void_p), C.c_longlong(timestamp),
                                              C.c_float(weightMultiplier), _hp(pose) if False else pose.ctypes.data_as(C.POINTER(C.c_float)), stream_ptr()))
        return pose.reshape(4, 4)

    def processFramePinned(self, rgb_pinned, depth_pinned, pose_out, timestamp=0, weightMultiplier=1.0):
        """torch pinned host tensors in, numpy float32[16] out"""
        check(lib().hrbf_fusion_process_frame(self._h, C.c_void_p(rgb_pinned.data_ptr()), C.c_void_p(depth_pinned.data_ptr()), C.c_longlong(timestamp),
                                              C.c_float(weightMultiplier), pose_out.ctypes.data_as(C.POINTER(C.c_float)), stream_ptr()))

    def processFrameDev(self, rgb_dev, depth_dev, timestamp=0, weightMultiplier=1.0):
        """CUDA tensors in; enqueue only"""
        check(lib().hrbf_fusion_process_frame_dev(self._h, ptr(rgb_dev), ptr(depth_dev), C.c_longlong(timestamp), C.c_float(weightMultiplier), stream_ptr()))

    def stageFrame(self, rgb, depth):
        """upload + preprocess of the next frame on the internal staging stream; CUDA tensors or pinned host tensors"""
        host = 0 if rgb.is_cuda else 1
        check(lib().hrbf_fusion_stage_frame(self._h, C.c_void_p(rgb.data_ptr()), C.c_void_p(depth.data_ptr()), host, stream_ptr()))

    def processStaged(self, pose_out=None, timestamp=0, weightMultiplier=1.0):
        """processFrame of the oldest staged frame; pose_out: numpy float32[16] (synchronises) or None (enqueue only)"""
        p = pose_out.ctypes.data_as(C.POINTER(C.c_float)) if pose_out is not None else None
        check(lib().hrbf_fusion_process_staged(self._h, C.c_longlong(timestamp), C.c_float(weightMultiplier), p, stream_ptr()))

    def getPose(self):
        pose = np.zeros(16, np.float32)
        check(lib().hrbf_fusion_get_pose(self._h, pose.ctypes.data_as(C.POINTER(C.c_float)), stream_ptr()))
        return pose.reshape(4, 4)

    def savePly(self, filename, confThreshold=0.0):
        """HRBFFusion::savePly (HRBFFusion.cpp:1737-1853); confThreshold = globalOutputSavePointCloudConfThreshold"""
        with open(filename, "wb") as f:
            f.write(self.globalModel.exportPly(confThreshold))

    def odometry(self):
        """the pipeline's RGBDOdometry (borrowed view: pyramids, Sobel images, candidate masks of the frame last tracked)"""
        from .odometry import RGBDOdometry
        return RGBDOdometry.borrowed(lib().hrbf_fusion_odometry(self._h), self.width, self.height)

    def trajectory(self):
        n = C.c_int(0)
        p = lib().hrbf_fusion_trajectory_dev(self._h, C.byref(n))
        return alias_tensor(p, (max(n.value, 1), 12), torch.float32)[:n.value]

    def enableTimings(self, on=True):
        check(lib().hrbf_fusion_enable_timings(self._h, int(on)))

    def lastTimings(self):
        ms = (C.c_float * 4)()
        check(lib().hrbf_fusion_last_timings(self._h, ms))
        return dict(zip(("Initialization", "Registration", "Integration", "Prediction"), [float(x) for x in ms]))
