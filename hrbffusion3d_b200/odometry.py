"""Host mirror of the reference's RGBDOdometry (Core/src/Utils/RGBDOdometry.h:57-107) and of the
cudafuncs.cuh step functions, over the C ABI.  GPUTexture* arguments become dense CUDA tensors:
RGBA32F -> float32 [h, w, 4]; R32F -> float32 [h, w]; RGBA8 -> uint8 [h, w, 4]."""
import ctypes as C
import math

import numpy as np
import torch

from ._lib import Camera, IcpOptions, TrackStats, check, lib, ptr, stream_ptr

MAP_NAMES = ["vmap_g_prev", "nmap_g_prev", "ck1_g_prev", "ck2_g_prev", "vmap_curr", "nmap_curr", "ck1_curr", "ck2_curr", "icpWeight"]


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _hp(a, t=C.c_float):
    return a.ctypes.data_as(C.POINTER(t))


class RGBDOdometry:
    NUM_PYRS = 3

    def __init__(self, width, height, cx, cy, fx, fy, distThresh=0.1, angleThresh=math.sin(20.0 * 3.14159265 / 180.0)):
        self.width, self.height = width, height
        self._h = C.c_void_p()
        check(lib().hrbf_odometry_create(C.byref(self._h), width, height, C.c_float(cx), C.c_float(cy), C.c_float(fx), C.c_float(fy),
                                         C.c_float(distThresh), C.c_float(angleThresh)))
        self.last_stats = None

    @classmethod
    def borrowed(cls, handle, width, height):
        """a view of an odometry object somebody else owns (hrbf_fusion_odometry): the accessors work, nothing is destroyed"""
        o = cls.__new__(cls)
        o.width, o.height, o._h, o._borrowed, o.last_stats = width, height, C.c_void_p(handle), True, None
        return o

    def close(self):
        if getattr(self, "_h", None) and self._h.value and not getattr(self, "_borrowed", False):
            lib().hrbf_odometry_destroy(self._h)
        self._h = C.c_void_p()

    def __del__(self):
        # at interpreter exit module globals (lib, C) may already be gone: nothing to release then, the process is ending
        try:
            self.close()
        except Exception:
            pass

    def setTracker(self, use_kernel_graph=False):
        """False (default): one persistent cooperative kernel; True: one kernel per reduction in a CUDA graph"""
        check(lib().hrbf_odometry_set_tracker(self._h, int(use_kernel_graph)))

    def setTrackerTiles(self, resident=True):
        """persistent tracker: keep every level's ICP tile in shared memory across its iterations (default) or re-read the maps"""
        check(lib().hrbf_odometry_set_tracker_tiles(self._h, int(resident)))

    def setTrackerThreads(self, threads=512):
        """512 (default): the persistent tracker fills every SM; 256: leaves half of each SM to other sequences' kernels"""
        check(lib().hrbf_odometry_set_tracker_threads(self._h, int(threads)))

    def setParams(self, curvValidThreshold=300.0, useCorrespondenceSearch=False, searchRadius=2, rgbUseGradientWeight=False):
        check(lib().hrbf_odometry_set_params(self._h, C.c_float(curvValidThreshold), int(useCorrespondenceSearch), int(searchRadius), int(rgbUseGradientWeight)))

    # ---- RGBDOdometry.cpp:161-247, 689-794 ----
    def initICP_depth(self, depth, depthCutoff, depthMapFactor):
        check(lib().hrbf_odometry_init_icp_depth(self._h, ptr(depth), C.c_float(depthCutoff), C.c_float(depthMapFactor), stream_ptr()))

    def initICP(self, vertices, normals, depthCutoff=20.0):
        check(lib().hrbf_odometry_init_icp(self._h, ptr(vertices), ptr(normals), C.c_float(depthCutoff), stream_ptr()))

    def initICPModel(self, vertices, normals, depthCutoff, modelPose):
        check(lib().hrbf_odometry_init_icp_model(self._h, ptr(vertices), ptr(normals), C.c_float(depthCutoff), _hp(_f32(modelPose)), stream_ptr()))

    def initRGB(self, rgba):
        check(lib().hrbf_odometry_init_rgb(self._h, ptr(rgba), stream_ptr()))

    def initRGBModel(self, rgba):
        check(lib().hrbf_odometry_init_rgb_model(self._h, ptr(rgba), stream_ptr()))

    def initFirstRGB(self, rgba):
        check(lib().hrbf_odometry_init_first_rgb(self._h, ptr(rgba), stream_ptr()))

    def initCurvature(self, k1, k2):
        check(lib().hrbf_odometry_init_curvature(self._h, ptr(k1), ptr(k2), stream_ptr()))

    def initCurvatureModel(self, k1, k2, modelPose):
        check(lib().hrbf_odometry_init_curvature_model(self._h, ptr(k1), ptr(k2), _hp(_f32(modelPose)), stream_ptr()))

    def initICPweight(self, w):
        check(lib().hrbf_odometry_init_icp_weight(self._h, ptr(w), stream_ptr()))

    def fillNeutralCurvature(self):
        check(lib().hrbf_odometry_fill_neutral_curvature(self._h, stream_ptr()))

    # ---- RGBDOdometry.cpp:796-1249 ----
    def getIncrementalTransformation(self, trans, rot, rgbOnly=False, icpWeight=10.0, pyramid=True, fastOdom=False, so3=True,
                                     if_curvature_info=True, index_frame=0):
        t = _f32(trans).reshape(3).copy()
        R = _f32(rot).reshape(9).copy()
        st = TrackStats()
        check(lib().hrbf_odometry_get_incremental_transformation(self._h, _hp(t), _hp(R), int(rgbOnly), C.c_float(icpWeight), int(pyramid),
                                                                 int(fastOdom), int(so3), int(if_curvature_info), int(index_frame),
                                                                 C.byref(st), stream_ptr()))
        self.last_stats = st
        return t, R.reshape(3, 3), st

    def trackAsync(self, prev_pose_dev, pose_out_dev, rgbOnly=False, icpWeight=10.0, pyramid=True, fastOdom=False, so3=True, if_curvature_info=True):
        check(lib().hrbf_odometry_track_async(self._h, ptr(prev_pose_dev), ptr(pose_out_dev), int(rgbOnly), C.c_float(icpWeight), int(pyramid),
                                              int(fastOdom), int(so3), int(if_curvature_info), stream_ptr()))

    def icpStepLevel(self, level, Rcurr, tcurr, Rprev_inv, tprev, use_weight=True, tiled=True):
        """icpStep (reduce.cu:580-693) on this object's own pyramid level; tiled = the TMA-staged tile form of the reduction"""
        A, b, res, sums = np.zeros(36, np.float32), np.zeros(6, np.float32), np.zeros(2, np.float32), np.zeros(29, np.float64)
        check(lib().hrbf_odometry_icp_step(self._h, int(level), _hp(_f32(Rcurr)), _hp(_f32(tcurr)), _hp(_f32(Rprev_inv)), _hp(_f32(tprev)), int(use_weight),
                                           int(tiled), _hp(A), _hp(b), _hp(res), _hp(sums, C.c_double), stream_ptr()))
        return A.reshape(6, 6), b, res, sums

    def timeKernel(self, which, level, with_update=0, reps=200):
        us = C.c_float()
        check(lib().hrbf_odometry_time_kernel(self._h, int(which), int(level), int(with_update), int(reps), C.byref(us), stream_ptr()))
        return us.value

    # ---- test / chaining views (copies) ----
    def map(self, which, level):
        idx = MAP_NAMES.index(which) if isinstance(which, str) else which
        rows, cols = self.height >> level, self.width >> level
        step = C.c_size_t()
        p = lib().hrbf_odometry_map(self._h, idx, level, C.byref(step))
        n = rows * cols * (1 if idx == 8 else 4)
        out = torch.empty(n, dtype=torch.float32, device="cuda")
        _memcpy_d2d(out.data_ptr(), p, n * 4)
        return out.view(-1, cols)

    def image(self, which, level):
        rows, cols = self.height >> level, self.width >> level
        p = lib().hrbf_odometry_image(self._h, which, level)
        out = torch.empty(rows * cols, dtype=torch.uint8, device="cuda")
        _memcpy_d2d(out.data_ptr(), p, rows * cols)
        return out.view(rows, cols)

    def gradient(self, axis, level):
        """next-image Sobel derivative (int16), axis 0 = dI/dx, 1 = dI/dy"""
        rows, cols = self.height >> level, self.width >> level
        p = lib().hrbf_odometry_gradient(self._h, axis, level)
        out = torch.empty(rows * cols, dtype=torch.int16, device="cuda")
        _memcpy_d2d(out.data_ptr(), p, rows * cols * 2)
        return out.view(rows, cols)

    def candidates(self, level):
        """pose-independent candidate mask of computeRgbResidual (uint8)"""
        rows, cols = self.height >> level, self.width >> level
        p = lib().hrbf_odometry_candidates(self._h, level)
        out = torch.empty(rows * cols, dtype=torch.uint8, device="cuda")
        _memcpy_d2d(out.data_ptr(), p, rows * cols)
        return out.view(rows, cols)

    def depth(self, which, level):
        rows, cols = self.height >> level, self.width >> level
        p = lib().hrbf_odometry_depth(self._h, which, level)
        out = torch.empty(rows * cols, dtype=torch.float32, device="cuda")
        _memcpy_d2d(out.data_ptr(), p, rows * cols * 4)
        return out.view(rows, cols)


def _memcpy_d2d(dst, src, nbytes):
    check(lib().hrbf_copy_device(C.c_void_p(dst), C.c_void_p(src), C.c_size_t(nbytes), stream_ptr()))


class ReduceWorkspace:
    def __init__(self):
        self.buf = torch.zeros(lib().hrbf_reduce_workspace_bytes(), dtype=torch.uint8, device="cuda")


def icpStep(Rcurr, tcurr, vmap_curr, nmap_curr, ck1_curr, ck2_curr, Rprev_inv, tprev, intr, vmap_g_prev, nmap_g_prev,
            ck1_g_prev, ck2_g_prev, icpWeightmap_g_prev, use_search=False, search_radius=2, use_weight=True,
            distThres=0.1, angleThres=math.sin(20.0 * 3.14159265 / 180.0), work=None, want_corres=False):
    """icpStep (Cuda/cudafuncs.cuh:82-116): SoA maps are CUDA float32 [4*rows, cols] (dense)."""
    rows, cols = vmap_curr.shape[0] // 4, vmap_curr.shape[1]
    work = work or ReduceWorkspace()
    A, b, res, sums = np.zeros(36, np.float32), np.zeros(6, np.float32), np.zeros(2, np.float32), np.zeros(29, np.float64)
    corres = torch.zeros((rows, cols, 2), dtype=torch.int32, device="cuda") if want_corres else None
    o = IcpOptions(int(use_search), int(search_radius), int(use_weight), distThres, angleThres)
    step = C.c_size_t(cols * 4)
    check(lib().hrbf_icp_step(_hp(_f32(Rcurr)), _hp(_f32(tcurr)), ptr(vmap_curr), ptr(nmap_curr), ptr(ck1_curr), ptr(ck2_curr), step,
                              _hp(_f32(Rprev_inv)), _hp(_f32(tprev)), Camera(*intr), ptr(vmap_g_prev), ptr(nmap_g_prev), ptr(ck1_g_prev),
                              ptr(ck2_g_prev), step, ptr(icpWeightmap_g_prev), step, rows, cols, C.byref(o), ptr(corres), ptr(work.buf),
                              _hp(A), _hp(b), _hp(res), _hp(sums, C.c_double), stream_ptr()))
    return A.reshape(6, 6), b, res, sums, corres


def computeRgbResidual(minScale, dIdx, dIdy, lastDepth, nextDepth, lastImage, nextImage, maxDepthDelta, kt, krkinv, work=None):
    rows, cols = nextImage.shape
    work = work or ReduceWorkspace()
    corr = torch.zeros((rows, cols, 16), dtype=torch.uint8, device="cuda")
    sig, cnt = C.c_int(0), C.c_int(0)
    check(lib().hrbf_compute_rgb_residual(C.c_float(minScale), ptr(dIdx), ptr(dIdy), ptr(lastDepth), ptr(nextDepth), ptr(lastImage), ptr(nextImage),
                                          ptr(corr), C.c_float(maxDepthDelta), _hp(_f32(kt)), _hp(_f32(krkinv)), rows, cols, ptr(work.buf),
                                          C.byref(sig), C.byref(cnt), stream_ptr()))
    return corr, sig.value, cnt.value


def rgbStep(corr, sigma, cloud3, fx, fy, dIdx, dIdy, use_grad_weight, sobelScale, work=None):
    rows, cols = corr.shape[:2]
    work = work or ReduceWorkspace()
    A, b, sums = np.zeros(36, np.float32), np.zeros(6, np.float32), np.zeros(29, np.float64)
    check(lib().hrbf_rgb_step(ptr(corr), C.c_float(sigma), ptr(cloud3), C.c_float(fx), C.c_float(fy), ptr(dIdx), ptr(dIdy), int(use_grad_weight),
                              C.c_float(sobelScale), rows, cols, ptr(work.buf), _hp(A), _hp(b), _hp(sums, C.c_double), stream_ptr()))
    return A.reshape(6, 6), b, sums


def so3Step(lastImage, nextImage, imageBasis, kinv, krlr, work=None):
    rows, cols = nextImage.shape
    work = work or ReduceWorkspace()
    A, b, res, sums = np.zeros(9, np.float32), np.zeros(3, np.float32), np.zeros(2, np.float32), np.zeros(11, np.float64)
    check(lib().hrbf_so3_step(ptr(lastImage), ptr(nextImage), _hp(_f32(imageBasis)), _hp(_f32(kinv)), _hp(_f32(krlr)), rows, cols, ptr(work.buf),
                              _hp(A), _hp(b), _hp(res), _hp(sums, C.c_double), stream_ptr()))
    return A.reshape(3, 3), b, res, sums
