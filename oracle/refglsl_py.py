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


def preprocess(pp, depth_u16):
    """The reference's preprocessing shaders in the order of HRBFFusion::processFrame (HRBFFusion.cpp:1017-1021): depth_bilateral.frag ->
    depth_metric_raw / depth_metric_filtered.frag -> depth_vertex_normal_radius.frag -> depth_curvature_gradient.frag (-> updateNormalRad).
    pp: orc_py.prep_params(...).  Returns the dict orc_py.preprocess returns.  Like there, every pass reads the PREVIOUS pass's output of
    the same implementation (this is the shader chain end to end)."""
    H, W = pp.rows, pp.cols
    raw = np.ascontiguousarray(depth_u16, np.uint16).astype(np.uint32)
    cam4 = (C.c_float(pp.cx), C.c_float(pp.cy), C.c_float(pp.fx), C.c_float(pp.fy))
    t = {k: np.zeros((H, W), np.float32) for k in ("filtered", "metric", "metric_filtered", "radius", "gradient_mag")}
    for k in ("vertex_raw", "vertex_filtered", "normal_pca", "curv1", "curv2", "normal_opt"):
        t[k] = np.zeros((H, W, 4), np.float32)
    L = lib()
    if pp.bilateral:
        L.glsl_depth_bilateral(W, H, _p(raw, C.c_uint), C.c_float(pp.depthFactor), C.c_float(pp.maxD), _p(t["filtered"]))
    else:
        raise NotImplementedError("depth_guass.frag is not wired (preprocessingUsebilateralFilter defaults to true)")
    L.glsl_depth_metric_raw(W, H, _p(raw, C.c_uint), C.c_float(pp.depthFactor), C.c_float(pp.maxD), _p(t["metric"]))
    L.glsl_depth_metric_filtered(W, H, _p(t["filtered"]), C.c_float(pp.depthFactor), C.c_float(pp.maxD), _p(t["metric_filtered"]))
    L.glsl_depth_vertex_normal_radius(W, H, _p(t["metric"]), _p(t["metric_filtered"]), *cam4, C.c_float(pp.radiusMultiplier), C.c_float(pp.pca),
                                      _p(t["vertex_raw"]), _p(t["vertex_filtered"]), _p(t["normal_pca"]), _p(t["radius"]))
    L.glsl_depth_curvature_gradient(W, H, _p(t["vertex_filtered"]), _p(t["normal_pca"]), *cam4, C.c_float(pp.maxD), C.c_float(pp.curvWindow),
                                    _p(t["curv1"]), _p(t["curv2"]), _p(t["gradient_mag"]), _p(t["normal_opt"]))
    t["normal"] = np.zeros((H, W, 4), np.float32)          # updateNormalRad: depth_update_normalrad.frag
    L.glsl_depth_update_normalrad(W, H, _p(t["normal_opt"]), _p(t["vertex_filtered"]), _p(t["normal"]))
    return t


def preprocess_stage(pp, name, src):
    """ONE shader pass on given inputs (a dict with the oracle's texture names) -> dict of that pass's outputs"""
    H, W = pp.rows, pp.cols
    cam4 = (C.c_float(pp.cx), C.c_float(pp.cy), C.c_float(pp.fx), C.c_float(pp.fy))
    L = lib()
    f1, f4 = (lambda: np.zeros((H, W), np.float32)), (lambda: np.zeros((H, W, 4), np.float32))
    if name == "metric_filtered":
        o = f1()
        L.glsl_depth_metric_filtered(W, H, _p(_f(src["filtered"])), C.c_float(pp.depthFactor), C.c_float(pp.maxD), _p(o))
        return {"metric_filtered": o}
    if name == "vertex_normal_radius":
        o = {"vertex_raw": f4(), "vertex_filtered": f4(), "normal_pca": f4(), "radius": f1()}
        L.glsl_depth_vertex_normal_radius(W, H, _p(_f(src["metric"])), _p(_f(src["metric_filtered"])), *cam4, C.c_float(pp.radiusMultiplier), C.c_float(pp.pca),
                                          _p(o["vertex_raw"]), _p(o["vertex_filtered"]), _p(o["normal_pca"]), _p(o["radius"]))
        return o
    if name == "curvature_gradient":
        o = {"curv1": f4(), "curv2": f4(), "gradient_mag": f1(), "normal_opt": f4()}
        L.glsl_depth_curvature_gradient(W, H, _p(_f(src["vertex_filtered"])), _p(_f(src["normal_pca"])), *cam4, C.c_float(pp.maxD), C.c_float(pp.curvWindow),
                                        _p(o["curv1"]), _p(o["curv2"]), _p(o["gradient_mag"]), _p(o["normal_opt"]))
        return o
    raise ValueError(name)


def vertexConfidence(pp, gradient_mag, metric, weighting, useConfEval=0, epsilon=1000.0):
    out = np.zeros((pp.rows, pp.cols), np.float32)
    lib().glsl_depth_confidence_evaluation(pp.cols, pp.rows, _p(_f(gradient_mag)), _p(_f(metric)), C.c_float(pp.cx), C.c_float(pp.cy), C.c_float(pp.fx),
                                           C.c_float(pp.fy), C.c_float(weighting), C.c_float(useConfEval), C.c_float(epsilon), _p(out))
    return out


def fillIn(pp, pred, frame, confidence, rgb, passthrough=0, lamb=10.0, curvThr=300.0):
    """Shaders/fill_vertex.frag, fill_normal.frag, fill_curvature.frag, fill_rgb.frag as HRBFFusion::predict runs them
    (HRBFFusion.cpp:1253-1259); arguments and result as orc_py.fillIn"""
    H, W = pp.rows, pp.cols
    cam4 = (C.c_float(pp.cx), C.c_float(pp.cy), C.c_float(pp.fx), C.c_float(pp.fy))
    f4 = lambda: np.zeros((H, W, 4), np.float32)
    o = {"vertex": f4(), "icpw": np.zeros((H, W), np.float32), "normal": f4(), "curvk1": f4(), "curvk2": f4()}
    L = lib()
    L.glsl_fill_vertex(W, H, _p(_f(pred["vertex"])), _p(_f(frame["vertex_filtered"])), _p(_f(frame["curv1"])), _p(_f(frame["curv2"])), _p(_f(pred["icpw"])),
                       _p(_f(confidence)), *cam4, int(passthrough), C.c_float(lamb), C.c_float(curvThr), _p(o["vertex"]), _p(o["icpw"]))
    L.glsl_fill_normal(W, H, _p(_f(pred["normal"])), _p(_f(frame["normal"])), *cam4, int(passthrough), _p(o["normal"]))
    L.glsl_fill_curvature(W, H, _p(_f(pred["curvk1"])), _p(_f(pred["curvk2"])), _p(_f(frame["curv1"])), _p(_f(frame["curv2"])), *cam4, int(passthrough),
                          _p(o["curvk1"]), _p(o["curvk2"]))
    img = f4()
    e = np.ascontiguousarray(pred["image"], np.uint8).astype(np.float32) / np.float32(255.0)
    r = np.ascontiguousarray(rgb, np.uint8).astype(np.float32) / np.float32(255.0)
    L.glsl_fill_rgb(W, H, _p(_f(e)), _p(_f(r)), int(passthrough), _p(img))
    o["image"] = _unorm8(img)
    return o


def predictIndices(pose, surfels, cam, width, height, maxDepth=20.0, active_kf=None):
    """Shaders/index_map.vert per surfel + the fixed-function point rasterisation restated in the driver; arguments and result as
    orc_py.predictIndices"""
    surfels = np.ascontiguousarray(surfels, np.float32).reshape(-1, 20)
    if active_kf is None:
        active_kf = np.zeros(19200, np.float32)
        active_kf[0] = 1.0
    active_kf = _f(active_kf)
    pinv = np.linalg.inv(np.asarray(pose, np.float64)).astype(np.float32)          # Eigen: pose.inverse() (IndexMap.cpp:207)
    out = {"index": np.zeros((height, width), np.uint32)}
    for k in ("vertConf", "colorTime", "normRad", "curvMax", "curvMin"):
        out[k] = np.zeros((height, width, 4), np.float32)
    lib().glsl_index_map(surfels.shape[0], _p(surfels), _p(_f(pinv)), C.c_float(cam[2]), C.c_float(cam[3]), C.c_float(cam[0]), C.c_float(cam[1]),
                         width, height, C.c_float(maxDepth), _p(active_kf), len(active_kf),
                         _p(out["index"], C.c_uint), _p(out["vertConf"]), _p(out["colorTime"]), _p(out["normRad"]), _p(out["curvMax"]), _p(out["curvMin"]))
    return out


def modelFuse(mp, pose, time, rgb, frame, confidence, idx, indexSubmap, surfels):
    """GlobalModel::fuse through the reference's own shaders: data.vert per pixel (association + candidate record), the scatter of the
    merge candidates into the update textures (fixed function: GL_LESS at constant depth = the first fragment of a texel wins;
    restated here), update.vert per surfel (merge).  Arguments and result as orc_py.modelFuse: (surfels after fuse, recorded vertices)."""
    H, W = mp.rows, mp.cols
    surfels = np.ascontiguousarray(surfels, np.float32).reshape(-1, 20)
    count = surfels.shape[0]
    tex = max(64, int(np.ceil(np.sqrt(count + 1))))
    tex += tex % 2
    rec = np.zeros((W * H, 20), np.float32)
    uid, best = np.zeros(W * H, np.int32), np.zeros(W * H, np.uint32)
    rgbf = np.ascontiguousarray(rgb, np.uint8).astype(np.float32) / np.float32(255.0)
    n = lib().glsl_data_vert(W, H, _p(_f(rgbf)), _p(_f(frame["metric"])), _p(_f(frame["metric_filtered"])), _p(_f(frame["curv1"])), _p(_f(frame["curv2"])),
                             _p(_f(confidence)), _p(np.ascontiguousarray(idx["index"], np.uint32), C.c_uint), _p(_f(idx["vertConf"])), _p(_f(idx["colorTime"])),
                             _p(_f(idx["normRad"])), C.c_float(mp.cx), C.c_float(mp.cy), C.c_float(mp.fx), C.c_float(mp.fy), _p(_f(pose)), C.c_float(mp.maxDepth),
                             C.c_float(time), C.c_float(indexSubmap), C.c_float(mp.radiusMultiplier), C.c_float(mp.pca), C.c_float(tex),
                             _p(rec), _p(uid, C.c_int), _p(best, C.c_uint))
    rec, uid, best = rec[:n], uid[:n], best[:n]
    # data.frag writes a merge candidate's record at its surfel's texel; all fragments have the same depth, GL_LESS keeps the first
    upd = np.zeros((5, tex * tex, 4), np.float32)
    cand = np.flatnonzero(uid == 1)
    _, first = np.unique(best[cand], return_index=True)
    win = cand[first]
    for k in range(5):
        upd[k, best[win]] = rec[win, 4 * k:4 * k + 4]
    out = np.zeros((max(count, 1), 20), np.float32)
    if count:
        lib().glsl_update_vert(count, _p(surfels), tex, _p(upd[0]), _p(upd[1]), _p(upd[2]), _p(upd[3]), _p(upd[4]), int(time), _p(out))
    return out[:count].copy(), rec.copy()


def modelClean(mp, pose, time, idx, surfels, unstable, active_kf=None):
    """GlobalModel::clean through Shaders/copy_unstable.vert / .geom: the model's surfels, then the vertices fuse recorded; as orc_py.modelClean"""
    if active_kf is None:
        active_kf = np.zeros(19200, np.float32)
        active_kf[0] = 1.0
    active_kf = _f(active_kf)
    verts = np.ascontiguousarray(np.concatenate([np.asarray(surfels, np.float32).reshape(-1, 20), np.asarray(unstable, np.float32).reshape(-1, 20)]))
    out = np.zeros((verts.shape[0] + 1, 20), np.float32)
    pinv = np.linalg.inv(np.asarray(pose, np.float64)).astype(np.float32)
    n = lib().glsl_copy_unstable(verts.shape[0], _p(verts), mp.cols, mp.rows, _p(np.ascontiguousarray(idx["index"], np.uint32), C.c_uint), _p(_f(idx["vertConf"])),
                                 _p(_f(idx["colorTime"])), _p(_f(idx["normRad"])), _p(_f(pose)), _p(_f(pinv)), C.c_float(mp.cx), C.c_float(mp.cy),
                                 C.c_float(mp.fx), C.c_float(mp.fy), int(time), C.c_float(mp.confThreshold), C.c_float(mp.cleanWindow), C.c_float(mp.curvThr),
                                 C.c_float(mp.maxDepth), _p(active_kf), len(active_kf), _p(out))
    return out[:n].copy()


def modelInitialise(mp, pose, frame, rgb, useConfEval=0, epsilon=1000.0):
    """GlobalModel::initialise through Shaders/init_unstableTex.vert / .geom; as orc_py.modelInitialise"""
    H, W = mp.rows, mp.cols
    out = np.zeros((W * H, 20), np.float32)
    rgbf = np.ascontiguousarray(rgb, np.uint8).astype(np.float32) / np.float32(255.0)
    n = lib().glsl_init_unstable(W, H, _p(_f(frame["vertex_raw"])), _p(_f(frame["normal"])), _p(_f(rgbf)), _p(_f(frame["curv1"])), _p(_f(frame["curv2"])),
                                 _p(_f(frame["gradient_mag"])), C.c_float(mp.cx), C.c_float(mp.cy), C.c_float(mp.fx), C.c_float(mp.fy), _p(_f(pose)),
                                 C.c_float(mp.curvThr), C.c_float(useConfEval), C.c_float(epsilon), _p(out))
    return out[:n].copy()


def denseEnough(vertex, thresh=0.75):
    """Resize::vertex (resize.frag at 1/20 of the resolution) + HRBFFusion::denseEnough's count (HRBFFusion.cpp:974-987, host code restated:
    sum of z > 0 over the sampled image, float(sum) / float(n) > thresh); as orc_py.denseEnough"""
    H, W = vertex.shape[:2]
    w, h = W // 20, H // 20
    out = np.zeros((h, w, 4), np.float32)
    lib().glsl_resize(W, H, _p(_f(vertex)), w, h, _p(out))
    per = np.float32(int((out[..., 2] > 0).sum())) / np.float32(h * w)
    return bool(per > np.float32(thresh)), out


def modelUpdate(surfels, delta):
    """GlobalModel::updateModel through Shaders/update_delta_trans.vert; as orc_py.modelUpdate (delta [n, 4, 4] row-major)"""
    s = np.ascontiguousarray(surfels, np.float32).reshape(-1, 20)
    d = np.ascontiguousarray(delta, np.float32).reshape(-1, 16)
    out = np.zeros_like(s)
    lib().glsl_update_delta_trans(s.shape[0], _p(s), _p(d), d.shape[0], _p(out))
    return out
