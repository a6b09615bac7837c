"""ctypes binding of the CPU oracle (oracle/liborc.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
u16p = np.ctypeslib.ndpointer(dtype=np.uint16, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
i16p = np.ctypeslib.ndpointer(dtype=np.int16, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


class Cam(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float)]


class IcpOpts(C.Structure):
    _fields_ = [("use_search", C.c_int), ("radius", C.c_int), ("use_weight", C.c_int),
                ("dist_thres", C.c_float), ("angle_thres", C.c_float)]


class TrackOpts(C.Structure):
    _fields_ = [("rgbOnly", C.c_int), ("icpWeight", C.c_float), ("pyramid", C.c_int), ("fastOdom", C.c_int),
                ("so3", C.c_int), ("if_curvature_info", C.c_int), ("use_search", C.c_int),
                ("search_radius", C.c_int), ("rgb_grad_weight", C.c_int)]


class TrackStats(C.Structure):
    _fields_ = [("lastICPError", C.c_float), ("lastICPCount", C.c_float), ("lastRGBError", C.c_float),
                ("lastRGBCount", C.c_float), ("lastSO3Error", C.c_float), ("lastSO3Count", C.c_float),
                ("lastA", C.c_double * 36), ("lastb", C.c_double * 6), ("icp_iterations_run", C.c_int)]


class SplatParams(C.Structure):
    _fields_ = [("cx", C.c_float), ("cy", C.c_float), ("fx", C.c_float), ("fy", C.c_float),
                ("cols", C.c_int), ("rows", C.c_int), ("maxDepth", C.c_float)]


class PredictParams(C.Structure):
    _fields_ = [("cx", C.c_float), ("cy", C.c_float), ("fx", C.c_float), ("fy", C.c_float),
                ("cols", C.c_int), ("rows", C.c_int), ("win", C.c_int), ("minNeighbors", C.c_int),
                ("maxNeighbors", C.c_int), ("confThreshold", C.c_float), ("icpWeightLambda", C.c_float)]


class ModelParams(C.Structure):
    _fields_ = [("cx", C.c_float), ("cy", C.c_float), ("fx", C.c_float), ("fy", C.c_float),
                ("cols", C.c_int), ("rows", C.c_int), ("maxDepth", C.c_float), ("confThreshold", C.c_float),
                ("radiusMultiplier", C.c_float), ("curvThr", C.c_float), ("pca", C.c_int), ("cleanWindow", C.c_int)]


class PrepParams(C.Structure):
    _fields_ = [("cx", C.c_float), ("cy", C.c_float), ("fx", C.c_float), ("fy", C.c_float),
                ("cols", C.c_int), ("rows", C.c_int), ("depthFactor", C.c_float), ("maxD", C.c_float),
                ("radiusMultiplier", C.c_float), ("pca", C.c_int), ("curvWindow", C.c_float), ("bilateral", C.c_int)]


def build():
    """(Re)build oracle/liborc.so with gcc."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "liborc.so"])


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(_HERE, "liborc.so")
    if not os.path.exists(path):
        build()
    L = C.CDLL(path)
    L.orc_odom_create.restype = C.c_void_p
    L.orc_odom_create.argtypes = [C.c_int, C.c_int] + [C.c_float] * 6
    L.orc_odom_map.restype = C.POINTER(C.c_float)
    L.orc_odom_map.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.orc_odom_image.restype = C.POINTER(C.c_ubyte)
    L.orc_odom_image.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.orc_odom_depth.restype = C.POINTER(C.c_float)
    L.orc_odom_depth.argtypes = [C.c_void_p, C.c_int, C.c_int]
    _LIB = L
    return L


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t=C.c_float):
    return a.ctypes.data_as(C.POINTER(t))


# ------------------------------------------------------------------ row 5 --
def copyMaps(v_aos, n_aos):
    rows, cols = v_aos.shape[:2]
    v, n = np.empty((4 * rows, cols), np.float32), np.empty((4 * rows, cols), np.float32)
    lib().orc_copyMaps(rows, cols, _p(_f(v_aos)), _p(_f(n_aos)), _p(v), _p(n))
    return v, n


def copyCurvatureMap(c_aos, thr):
    rows, cols = c_aos.shape[:2]
    c = np.empty((4 * rows, cols), np.float32)
    lib().orc_copyCurvatureMap(rows, cols, _p(_f(c_aos)), _p(c), C.c_float(thr))
    return c


def copyicpWeightMap(w):
    rows, cols = w.shape
    o = np.empty((rows, cols), np.float32)
    lib().orc_copyicpWeightMap(rows, cols, _p(_f(w)), _p(o))
    return o


def resizeMap(src, normalize, init=None):
    rows, cols = src.shape[0] // 4, src.shape[1]
    dst = np.full((4 * (rows // 2), cols // 2), np.nan, np.float32) if init is None else init.copy()
    lib().orc_resizeMap(rows // 2, cols // 2, _p(_f(src)), _p(dst), int(normalize))
    return dst


def resizeCMap(src):
    rows, cols = src.shape[0] // 4, src.shape[1]
    dst = np.full((4 * (rows // 2), cols // 2), np.nan, np.float32)
    lib().orc_resizeCMap(rows // 2, cols // 2, _p(_f(src)), _p(dst))
    return dst


def resizeicpWeightMap(src):
    rows, cols = src.shape
    dst = np.empty((rows // 2, cols // 2), np.float32)
    lib().orc_resizeicpWeightMap(rows // 2, cols // 2, _p(_f(src)), _p(dst))
    return dst


def tranformMaps(v, n, R, t):
    rows, cols = v.shape[0] // 4, v.shape[1]
    vd, nd = v.copy(), n.copy()
    lib().orc_tranformMaps(rows, cols, _p(_f(v)), _p(_f(n)), _p(_f(R)), _p(_f(t)), _p(vd), _p(nd))
    return vd, nd


def transformCurvMaps(k1, k2, R, t):
    rows, cols = k1.shape[0] // 4, k1.shape[1]
    a, b = k1.copy(), k2.copy()
    lib().orc_transformCurvMaps(rows, cols, _p(_f(k1)), _p(_f(k2)), _p(_f(R)), _p(_f(t)), _p(a), _p(b))
    return a, b


# --------------------------------------------------------------- rows 1-3 --
def icpStep(Rcurr, tcurr, vc, nc, k1c, k2c, Rprev_inv, tprev, cam, vg, ng, k1g, k2g, w,
            use_search=0, radius=2, use_weight=1, dist_thres=0.1, angle_thres=float(np.sin(np.deg2rad(20.0))),
            want_corres=False):
    rows, cols = vc.shape[0] // 4, vc.shape[1]
    A, b, res = np.zeros(36, np.float32), np.zeros(6, np.float32), np.zeros(2, np.float32)
    sums = np.zeros(29, np.float64)
    corres = np.zeros((rows, cols, 2), np.int32) if want_corres else None
    o = IcpOpts(use_search, radius, use_weight, dist_thres, angle_thres)
    lib().orc_icpStep(rows, cols, _p(_f(Rcurr)), _p(_f(tcurr)), _p(_f(vc)), _p(_f(nc)), _p(_f(k1c)), _p(_f(k2c)),
                      _p(_f(Rprev_inv)), _p(_f(tprev)), Cam(*cam), _p(_f(vg)), _p(_f(ng)), _p(_f(k1g)), _p(_f(k2g)),
                      _p(_f(w)), C.byref(o), _p(A), _p(b), _p(res), _p(sums, C.c_double),
                      _p(corres, C.c_int) if want_corres else None)
    return A.reshape(6, 6), b, res, sums, corres


DATATERM = np.dtype([("zx", np.int16), ("zy", np.int16), ("ox", np.int16), ("oy", np.int16),
                     ("diff", np.float32), ("valid", np.uint8), ("pad", np.uint8, 3)])


def computeRgbResidual(minScale, dIdx, dIdy, lastDepth, nextDepth, lastImage, nextImage, maxDepthDelta, kt, krkinv):
    rows, cols = nextImage.shape
    corr = np.zeros((rows, cols), DATATERM)
    sig, cnt = C.c_int(0), C.c_int(0)
    lib().orc_computeRgbResidual(rows, cols, C.c_float(minScale), _p(dIdx, C.c_short), _p(dIdy, C.c_short),
                                 _p(_f(lastDepth)), _p(_f(nextDepth)), _p(lastImage, C.c_ubyte), _p(nextImage, C.c_ubyte),
                                 corr.ctypes.data_as(C.c_void_p), C.c_float(maxDepthDelta), _p(_f(kt)), _p(_f(krkinv)),
                                 C.byref(sig), C.byref(cnt))
    return corr, sig.value, cnt.value


def rgbStep(corr, sigma, cloud3, fx, fy, dIdx, dIdy, use_grad_weight, sobelScale):
    rows, cols = corr.shape
    A, b, sums = np.zeros(36, np.float32), np.zeros(6, np.float32), np.zeros(29, np.float64)
    lib().orc_rgbStep(rows, cols, corr.ctypes.data_as(C.c_void_p), C.c_float(sigma), _p(_f(cloud3)), C.c_float(fx), C.c_float(fy),
                      _p(dIdx, C.c_short), _p(dIdy, C.c_short), int(use_grad_weight), C.c_float(sobelScale),
                      _p(A), _p(b), _p(sums, C.c_double))
    return A.reshape(6, 6), b, sums


def so3Step(lastImage, nextImage, imageBasis, kinv, krlr):
    rows, cols = nextImage.shape
    A, b, res, sums = np.zeros(9, np.float32), np.zeros(3, np.float32), np.zeros(2, np.float32), np.zeros(11, np.float64)
    lib().orc_so3Step(rows, cols, _p(lastImage, C.c_ubyte), _p(nextImage, C.c_ubyte), _p(_f(imageBasis)), _p(_f(kinv)), _p(_f(krlr)),
                      _p(A), _p(b), _p(res), _p(sums, C.c_double))
    return A.reshape(3, 3), b, res, sums


def sobel(img):
    rows, cols = img.shape
    dx, dy = np.zeros((rows, cols), np.int16), np.zeros((rows, cols), np.int16)
    lib().orc_sobel(rows, cols, _p(img, C.c_ubyte), _p(dx, C.c_short), _p(dy, C.c_short))
    return dx, dy


def projectToPointCloud(depth, cam_level):
    rows, cols = depth.shape
    cl = np.zeros((rows, cols, 3), np.float32)
    lib().orc_projectToPointCloud(rows, cols, _p(_f(depth)), _p(cl), Cam(*cam_level))
    return cl


# ------------------------------------------------------------------ row 4 --
class Odometry:
    """Mirror of the reference's RGBDOdometry on the CPU oracle."""
    MAPS = ["vmap_g_prev", "nmap_g_prev", "ck1_g_prev", "ck2_g_prev", "vmap_curr", "nmap_curr", "ck1_curr", "ck2_curr", "icpWeight"]

    def __init__(self, width, height, cx, cy, fx, fy, distThresh=0.1, angleThresh=float(np.sin(20.0 * 3.14159265 / 180.0))):
        self.w, self.h = width, height
        self.o = lib().orc_odom_create(width, height, cx, cy, fx, fy, distThresh, angleThresh)
        self.curvThr = 300.0

    def __del__(self):
        if getattr(self, "o", None) and _LIB is not None and C is not None:
            _LIB.orc_odom_destroy(C.c_void_p(self.o))
            self.o = None

    def initICP_depth(self, depth_f32, cutoff, factor):
        lib().orc_odom_initICP_depth(C.c_void_p(self.o), _p(_f(depth_f32)), C.c_float(cutoff), C.c_float(factor))

    def initICP(self, v, n, cutoff=20.0):
        lib().orc_odom_initICP(C.c_void_p(self.o), _p(_f(v)), _p(_f(n)), C.c_float(cutoff))

    def initICPModel(self, v, n, cutoff, pose):
        lib().orc_odom_initICPModel(C.c_void_p(self.o), _p(_f(v)), _p(_f(n)), C.c_float(cutoff), _p(_f(pose)))

    def initRGB(self, rgba):
        lib().orc_odom_initRGB(C.c_void_p(self.o), _p(np.ascontiguousarray(rgba, np.uint8), C.c_ubyte))

    def initRGBModel(self, rgba):
        lib().orc_odom_initRGBModel(C.c_void_p(self.o), _p(np.ascontiguousarray(rgba, np.uint8), C.c_ubyte))

    def initFirstRGB(self, rgba):
        lib().orc_odom_initFirstRGB(C.c_void_p(self.o), _p(np.ascontiguousarray(rgba, np.uint8), C.c_ubyte))

    def initCurvature(self, k1, k2):
        lib().orc_odom_initCurvature(C.c_void_p(self.o), _p(_f(k1)), _p(_f(k2)), C.c_float(self.curvThr))

    def initCurvatureModel(self, k1, k2, pose):
        lib().orc_odom_initCurvatureModel(C.c_void_p(self.o), _p(_f(k1)), _p(_f(k2)), _p(_f(pose)), C.c_float(self.curvThr))

    def initICPweight(self, w):
        lib().orc_odom_initICPweight(C.c_void_p(self.o), _p(_f(w)))

    def fillNeutralCurvature(self):
        lib().orc_odom_fillNeutralCurvature(C.c_void_p(self.o))

    def getIncrementalTransformation(self, trans, rot, rgbOnly=False, icpWeight=10.0, pyramid=True, fastOdom=False, so3=True,
                                     if_curvature_info=True, use_search=0, search_radius=2, rgb_grad_weight=0):
        t = _f(trans).copy().reshape(3)
        R = _f(rot).copy().reshape(9)
        o = TrackOpts(int(rgbOnly), icpWeight, int(pyramid), int(fastOdom), int(so3), int(if_curvature_info),
                      use_search, search_radius, rgb_grad_weight)
        st = TrackStats()
        lib().orc_odom_getIncrementalTransformation(C.c_void_p(self.o), _p(t), _p(R), C.byref(o), C.byref(st))
        return t, R.reshape(3, 3), st

    def map(self, which, level):
        idx = self.MAPS.index(which) if isinstance(which, str) else which
        rows, cols = self.h >> level, self.w >> level
        n = rows * cols * (1 if idx == 8 else 4)
        p = lib().orc_odom_map(C.c_void_p(self.o), idx, level)
        return np.ctypeslib.as_array(p, shape=(n,)).reshape(-1, cols).copy()

    def image(self, which, level):
        rows, cols = self.h >> level, self.w >> level
        p = lib().orc_odom_image(C.c_void_p(self.o), which, level)
        return np.ctypeslib.as_array(p, shape=(rows * cols,)).reshape(rows, cols).copy()

    def depth(self, which, level):
        rows, cols = self.h >> level, self.w >> level
        p = lib().orc_odom_depth(C.c_void_p(self.o), which, level)
        return np.ctypeslib.as_array(p, shape=(rows * cols,)).reshape(rows, cols).copy()


# --------------------------------------------------------------- rows 6-7 --
def predictIndices(pose, surfels, cam, width, height, maxDepth=20.0, active_kf=None):
    """-> dict(index u32 [h,w], vertConf, colorTime, normRad, curvMax, curvMin f32 [h,w,4])"""
    surfels = np.ascontiguousarray(surfels, np.float32).reshape(-1, 20)
    P = SplatParams(cam[2], cam[3], cam[0], cam[1], width, height, maxDepth)
    if active_kf is None:
        active_kf = np.zeros(19200, np.float32)
        active_kf[0] = 1.0
    out = {"index": np.zeros((height, width), np.uint32)}
    for k in ("vertConf", "colorTime", "normRad", "curvMax", "curvMin"):
        out[k] = np.zeros((height, width, 4), np.float32)
    lib().orc_predictIndices(_p(_f(pose)), _p(surfels), C.c_int(surfels.shape[0]), C.byref(P), _p(_f(active_kf)), C.c_int(len(active_kf)),
                             _p(out["index"], C.c_uint32), _p(out["vertConf"]), _p(out["colorTime"]), _p(out["normRad"]),
                             _p(out["curvMax"]), _p(out["curvMin"]))
    return out


def predictHRBF(idx, cam, width, height, win=3, minNeighbors=6, maxNeighbors=10, confThreshold=3.0, icpWeightLambda=10.0):
    """idx: dict from predictIndices -> dict(image u8 [h,w,4], vertex, normal, curvk1, curvk2 f32 [h,w,4], time u16, icpw f32)"""
    P = PredictParams(cam[2], cam[3], cam[0], cam[1], width, height, win, minNeighbors, maxNeighbors, confThreshold, icpWeightLambda)
    out = {"image": np.zeros((height, width, 4), np.uint8), "time": np.zeros((height, width), np.uint16),
           "icpw": np.zeros((height, width), np.float32)}
    for k in ("vertex", "normal", "curvk1", "curvk2"):
        out[k] = np.zeros((height, width, 4), np.float32)
    lib().orc_predictHRBF(C.byref(P), _p(_f(idx["vertConf"])), _p(_f(idx["colorTime"])), _p(_f(idx["normRad"])), _p(_f(idx["curvMax"])),
                          _p(_f(idx["curvMin"])), _p(out["image"], C.c_ubyte), _p(out["vertex"]), _p(out["normal"]), _p(out["curvk1"]),
                          _p(out["curvk2"]), _p(out["time"], C.c_ushort), _p(out["icpw"]))
    return out


# ----------------------------------------------------------- rows 8-10 --
def prep_params(cam, width, height, depthFactor=1.0 / 5000.0, maxD=3.5, radiusMultiplier=4.0, pca=1, curvWindow=3.0, bilateral=1):
    return PrepParams(cam[2], cam[3], cam[0], cam[1], width, height, depthFactor, maxD, radiusMultiplier, pca, curvWindow, bilateral)


def model_params(cam, width, height, maxDepth=20.0, confThreshold=5.0, radiusMultiplier=4.0, curvThr=300.0, pca=1, cleanWindow=2):
    return ModelParams(cam[2], cam[3], cam[0], cam[1], width, height, maxDepth, confThreshold, radiusMultiplier, curvThr, pca, cleanWindow)


def preprocess(pp, depth_u16):
    """filterDepth -> metriciseDepth -> computeVertexNormalRadius -> computeCurvatureGradient -> updateNormalRad
    (HRBFFusion.cpp:1017-1021).  Returns the textures dict (AoS float32)."""
    H, W = pp.rows, pp.cols
    depth_u16 = np.ascontiguousarray(depth_u16, np.uint16)
    t = {"filtered": np.zeros((H, W), np.float32), "metric": np.zeros((H, W), np.float32), "metric_filtered": np.zeros((H, W), np.float32)}
    for k in ("vertex_raw", "vertex_filtered", "normal_pca", "curv1", "curv2", "normal_opt"):
        t[k] = np.zeros((H, W, 4), np.float32)
    t["radius"] = np.zeros((H, W), np.float32)
    t["gradient_mag"] = np.zeros((H, W), np.float32)
    L = lib()
    L.orc_filterDepth(C.byref(pp), _p(depth_u16, C.c_ushort), _p(t["filtered"]))
    L.orc_metriciseDepth(C.byref(pp), _p(depth_u16, C.c_ushort), _p(t["filtered"]), _p(t["metric"]), _p(t["metric_filtered"]))
    L.orc_computeVertexNormalRadius(C.byref(pp), _p(t["metric"]), _p(t["metric_filtered"]), _p(t["vertex_raw"]), _p(t["vertex_filtered"]),
                                    _p(t["normal_pca"]), _p(t["radius"]))
    L.orc_computeCurvatureGradient(C.byref(pp), _p(t["vertex_filtered"]), _p(t["normal_pca"]), _p(t["curv1"]), _p(t["curv2"]),
                                   _p(t["gradient_mag"]), _p(t["normal_opt"]))
    t["normal"] = t["normal_opt"]          # updateNormalRad: NORMAL <- NORMAL_OPT
    return t


def vertexConfidence(pp, gradient_mag, weighting, useConfEval=0, epsilon=1000.0):
    out = np.zeros((pp.rows, pp.cols), np.float32)
    lib().orc_vertexConfidence(C.byref(pp), _p(_f(gradient_mag)), C.c_float(weighting), int(useConfEval), C.c_float(epsilon), _p(out))
    return out


def fillIn(pp, pred, frame, confidence, rgb, passthrough=0, lamb=10.0, curvThr=300.0):
    """pred: dict from predictHRBF; frame: dict from preprocess -> dict(vertex, icpw, normal, curvk1, curvk2, image)"""
    H, W = pp.rows, pp.cols
    o = {"vertex": np.zeros((H, W, 4), np.float32), "icpw": np.zeros((H, W), np.float32), "normal": np.zeros((H, W, 4), np.float32),
         "curvk1": np.zeros((H, W, 4), np.float32), "curvk2": np.zeros((H, W, 4), np.float32), "image": np.zeros((H, W, 4), np.uint8)}
    lib().orc_fillIn(C.byref(pp), int(passthrough), C.c_float(lamb), C.c_float(curvThr),
                     _p(_f(pred["vertex"])), _p(_f(pred["icpw"])), _p(_f(pred["normal"])), _p(_f(pred["curvk1"])), _p(_f(pred["curvk2"])),
                     _p(np.ascontiguousarray(pred["image"]), C.c_ubyte),
                     _p(_f(frame["vertex_filtered"])), _p(_f(frame["normal"])), _p(_f(frame["curv1"])), _p(_f(frame["curv2"])),
                     _p(_f(confidence)), _p(np.ascontiguousarray(rgb), C.c_ubyte),
                     _p(o["vertex"]), _p(o["icpw"]), _p(o["normal"]), _p(o["curvk1"]), _p(o["curvk2"]), _p(o["image"], C.c_ubyte))
    return o


def denseEnough(vertex, thresh=0.75):
    H, W = vertex.shape[:2]
    return bool(lib().orc_denseEnough(H, W, _p(_f(vertex)), C.c_float(thresh)))


def modelInitialise(mp, pose, frame, rgb, useConfEval=0, epsilon=1000.0):
    out = np.zeros((mp.cols * mp.rows, 20), np.float32)
    n = lib().orc_model_initialise(C.byref(mp), _p(_f(pose)), _p(_f(frame["vertex_raw"])), _p(_f(frame["normal"])),
                                   _p(np.ascontiguousarray(rgb), C.c_ubyte), _p(_f(frame["curv1"])), _p(_f(frame["curv2"])),
                                   _p(_f(frame["gradient_mag"])), int(useConfEval), C.c_float(epsilon), _p(out))
    return out[:n].copy()


def modelFuse(mp, pose, time, rgb, frame, confidence, idx, indexSubmap, surfels):
    count = surfels.shape[0]
    out = np.zeros((max(count, 1), 20), np.float32)
    un = np.zeros((mp.cols * mp.rows, 20), np.float32)
    n = lib().orc_model_fuse(C.byref(mp), _p(_f(pose)), int(time), _p(np.ascontiguousarray(rgb), C.c_ubyte), _p(_f(frame["metric"])),
                             _p(_f(frame["metric_filtered"])), _p(_f(frame["curv1"])), _p(_f(frame["curv2"])), _p(_f(confidence)),
                             _p(idx["index"], C.c_uint32), _p(_f(idx["vertConf"])), _p(_f(idx["colorTime"])), _p(_f(idx["normRad"])),
                             C.c_float(indexSubmap), _p(_f(surfels)), int(count), _p(out), _p(un))
    return out[:count].copy(), un[:n].copy()


def modelClean(mp, pose, time, idx, surfels, unstable, active_kf=None):
    if active_kf is None:
        active_kf = np.zeros(19200, np.float32)
        active_kf[0] = 1.0
    out = np.zeros((surfels.shape[0] + unstable.shape[0] + 1, 20), np.float32)
    n = lib().orc_model_clean(C.byref(mp), _p(_f(pose)), int(time), _p(idx["index"], C.c_uint32), _p(_f(idx["vertConf"])),
                              _p(_f(idx["colorTime"])), _p(_f(idx["normRad"])), _p(_f(active_kf)), int(len(active_kf)),
                              _p(_f(surfels)), int(surfels.shape[0]), _p(_f(unstable)), int(unstable.shape[0]), _p(out))
    return out[:n].copy()


def modelUpdate(surfels, delta):
    """GlobalModel::updateModel: delta [n, 4, 4] row-major rigid corrections indexed by sub-map id"""
    out = np.ascontiguousarray(surfels, np.float32).copy()
    d = np.ascontiguousarray(delta, np.float32).reshape(-1, 16)
    lib().orc_model_update(_p(out), int(out.shape[0]), _p(d), int(d.shape[0]))
    return out


def savePlyBytes(surfels, confThreshold=0.0):
    """HRBFFusion::savePly (HRBFFusion.cpp:1737-1853) restated byte for byte: header (:1760-1787), then per surfel with
    pos.w > globalOutputSavePointCloudConfThreshold (:1797) the 43-byte record x y z (:1811-1818), r g b from
    int(col[0]) >> 16 / >> 8 / & 0xFF (:1821-1827), the NEGATED normal (:1805-1807,1829-1836), curvature_max = c_max[3],
    curvature_min = c_min[3] (:1838-1842), radius = nor[3] (:1844), submapIndex = col[1] (:1847).  surfels: float32 [n, 20]."""
    s = np.ascontiguousarray(surfels, np.float32).reshape(-1, 20)
    keep = s[s[:, 3] > np.float32(confThreshold)]
    header = ("ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z"
              "\nproperty uchar red\nproperty uchar green\nproperty uchar blue\nproperty float nx\nproperty float ny\nproperty float nz"
              "\nproperty float curvature_max\nproperty float curvature_min\nproperty float radius\nproperty float submapIndex\nend_header\n" % keep.shape[0])
    out = bytearray(header.encode("ascii"))
    for v in keep:          # small maps only (test infrastructure)
        c = int(v[4])
        out += np.array([v[0], v[1], v[2]], "<f4").tobytes()
        out += bytes([(c >> 16) & 0xFF, (c >> 8) & 0xFF, c & 0xFF])
        out += np.array([-v[8], -v[9], -v[10], v[15], v[19], v[11], v[5]], "<f4").tobytes()
    return bytes(out)


# ---------------------------------------------------- row 5, the remaining single kernels (cudafuncs.cu) --
def pyrDownDepth(src):
    rows, cols = src.shape
    dst = np.zeros((rows // 2, cols // 2), np.float32)
    lib().orc_pyrDownDepth(rows, cols, _p(_f(src)), _p(dst))
    return dst


def createVMap(cam, depth, cutoff, factor):
    rows, cols = depth.shape
    v = np.zeros((4 * rows, cols), np.float32)
    lib().orc_createVMap(Cam(*cam), rows, cols, _p(_f(depth)), _p(v), C.c_float(cutoff), C.c_float(factor))
    return v


def createNMap(vmap):
    rows, cols = vmap.shape[0] // 4, vmap.shape[1]
    n = np.zeros((4 * rows, cols), np.float32)
    lib().orc_createNMap(rows, cols, _p(_f(vmap)), _p(n))
    return n


def verticesToDepth(v_aos, cutoff):
    rows, cols = v_aos.shape[:2]
    d = np.zeros((rows, cols), np.float32)
    lib().orc_verticesToDepth(rows, cols, _p(_f(v_aos)), _p(d), C.c_float(cutoff))
    return d


def pyrDownGaussF(src):
    rows, cols = src.shape
    dst = np.zeros((rows // 2, cols // 2), np.float32)
    lib().orc_pyrDownGaussF(rows, cols, _p(_f(src)), _p(dst))
    return dst


def pyrDownUcharGauss(src):
    rows, cols = src.shape
    dst = np.zeros((rows // 2, cols // 2), np.uint8)
    lib().orc_pyrDownUcharGauss(rows, cols, _p(np.ascontiguousarray(src, np.uint8), C.c_ubyte), _p(dst, C.c_ubyte))
    return dst


def rgbaToIntensity(rgba):
    rows, cols = rgba.shape[:2]
    dst = np.zeros((rows, cols), np.uint8)
    lib().orc_rgbaToIntensity(rows, cols, _p(np.ascontiguousarray(rgba, np.uint8), C.c_ubyte), _p(dst, C.c_ubyte))
    return dst
