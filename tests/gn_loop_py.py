"""The Gauss-Newton loop of RGBDOdometry::getIncrementalTransformation (Core/src/Utils/RGBDOdometry.cpp:796-1249) driven from
Python over the STEP functions (so3Step / computeRgbResidual / icpStep / rgbStep) of one or two back-ends -- test infrastructure.

`drive` = the back-end whose results move the pose (normally the CPU oracle).  `shadow` (optional) = a second back-end evaluated at
the SAME pose in every iteration ("teacher forcing"): its integer outputs and float sums are recorded beside the driver's, so that a
whole default-configuration frame can be compared iteration by iteration without the two trackers drifting apart.

A back-end is an object with
    maps(level) -> (vc, nc, k1c, k2c, vg, ng, k1g, k2g, w)        SoA maps of that pyramid level
    images(level) -> (lastImage, nextImage, lastNextImage)         u8
    depths(level) -> (lastDepth, nextDepth)                        f32
    sobel(level), cloud(level, cam_level)
    so3Step / computeRgbResidual / rgbStep / icpStep               as oracle.orc_py / hrbffusion3d_b200.odometry
Host arithmetic follows oracle/orc_odometry.c (double for K R K^-1 and the 6x6 solve, float for the pose composition)."""
import numpy as np

ITER = {"default": (10, 5, 4)}
MIN_GRAD = (5.0, 3.0, 1.0)
SOBEL_SCALE = 0.125
MAX_DEPTH_DELTA = 0.07


def rodrigues(w):
    w = np.asarray(w, np.float64)
    th = float(np.sqrt((w * w).sum()))
    R = np.eye(3)
    if th >= np.finfo(np.float64).eps:
        c, s = np.cos(th), np.sin(th)
        r = w / th
        rx = np.array([[0, -r[2], r[1]], [r[2], 0, -r[0]], [-r[1], r[0], 0]])
        R = c * np.eye(3) + (1 - c) * np.outer(r, r) + s * rx
    return R


def cam_level(cam, level):
    return tuple(np.float32(c) / np.float32(1 << level) for c in cam)      # (fx, fy, cx, cy), Cuda/types.cuh:93-97


def _K(cl):
    return np.array([[cl[0], 0, cl[2]], [0, cl[1], cl[3]], [0, 0, 1]], np.float64)


class OracleBackend:
    """step functions + pyramid views of an oracle.orc_py.Odometry whose init* calls have been made"""

    def __init__(self, orc, odom):
        self.orc, self.o = orc, odom

    def maps(self, l):
        return tuple(self.o.map(k, l) for k in ("vmap_curr", "nmap_curr", "ck1_curr", "ck2_curr", "vmap_g_prev", "nmap_g_prev", "ck1_g_prev", "ck2_g_prev", "icpWeight"))

    def images(self, l):
        return self.o.image(0, l), self.o.image(1, l), self.o.image(2, l)

    def depths(self, l):
        return self.o.depth(0, l), self.o.depth(1, l)

    def sobel(self, img):
        return self.orc.sobel(img)

    def cloud(self, depth, cl):
        return self.orc.projectToPointCloud(depth, cl)

    def so3Step(self, *a):
        return self.orc.so3Step(*a)

    def computeRgbResidual(self, *a):
        return self.orc.computeRgbResidual(*a)

    def rgbStep(self, *a):
        return self.orc.rgbStep(*a)

    def icpStep(self, Rc, tc, m, Rpi, tp, cl, use_weight):
        return self.orc.icpStep(Rc, tc, *m[:4], Rpi, tp, cl, *m[4:], use_search=0, radius=2, use_weight=int(use_weight))


class CudaBackend:
    """the same over hrbffusion3d_b200.odometry (device tensors); sobel / cloud come from the oracle's functions, uploaded"""

    def __init__(self, od, odom, orc, torch):
        self.od, self.o, self.orc, self.torch = od, odom, orc, torch

    def _dev(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a)).cuda()

    def maps(self, l):
        return tuple(self.o.map(k, l) for k in ("vmap_curr", "nmap_curr", "ck1_curr", "ck2_curr", "vmap_g_prev", "nmap_g_prev", "ck1_g_prev", "ck2_g_prev", "icpWeight"))

    def images(self, l):
        return self.o.image(0, l), self.o.image(1, l), self.o.image(2, l)

    def depths(self, l):
        return self.o.depth(0, l), self.o.depth(1, l)

    def sobel(self, img):
        dx, dy = self.orc.sobel(img.cpu().numpy())
        return self._dev(dx), self._dev(dy)

    def cloud(self, depth, cl):
        return self._dev(self.orc.projectToPointCloud(depth.cpu().numpy(), cl))

    def so3Step(self, *a):
        return self.od.so3Step(*a)

    def computeRgbResidual(self, *a):
        return self.od.computeRgbResidual(*a)

    def rgbStep(self, *a):
        return self.od.rgbStep(*a)

    def icpStep(self, Rc, tc, m, Rpi, tp, cl, use_weight):
        return self.od.icpStep(Rc, tc, *m[:4], Rpi, tp, cl, *m[4:], use_search=False, search_radius=2, use_weight=bool(use_weight))


def run(drive, cam, trans, rot, shadow=None, rgbOnly=False, icpWeight=10.0, pyramid=True, fastOdom=False, so3=True, if_curvature_info=True):
    """-> (trans, rot, log).  log: one dict per reduction with the driver's ("d_*") and the shadow's ("s_*") results."""
    icp = (not rgbOnly) and icpWeight > 0
    rgb = rgbOnly or icpWeight < 100
    Rprev, tprev = np.asarray(rot, np.float32).reshape(3, 3).copy(), np.asarray(trans, np.float32).reshape(3).copy()
    Rcurr, tcurr = Rprev.copy(), tprev.copy()
    backs = [("d", drive)] + ([("s", shadow)] if shadow is not None else [])
    log = []
    sob = {}
    if rgb:
        for tag, b in backs:
            sob[tag] = [b.sobel(b.images(l)[1]) for l in range(3)]
    resultR = np.eye(3)
    if so3:
        l = 2
        cl = cam_level(cam, l)
        K = _K(cl)
        Kinv = np.linalg.inv(K)
        R_lr = np.eye(3, dtype=np.float32)
        lastError = lastCount = np.float32(np.finfo(np.float32).max / 2)
        lastResultR = resultR.copy()
        for it in range(10):
            H = (K @ resultR @ Kinv).astype(np.float32)
            kinv = Kinv.astype(np.float32)
            krlr = (K @ resultR).astype(np.float32)
            rec = dict(kind="so3", level=l, it=it)
            for tag, b in backs:
                last, nxt, lastnext = b.images(l)
                A, bb, res, sums = b.so3Step(lastnext, nxt, H, kinv, krlr)
                rec[tag + "_sums"], rec[tag + "_res"] = np.array(sums), np.array(res)
                if tag == "d":
                    jtj, jtr, residual = A, bb, res
            log.append(rec)
            err = np.float32(np.sqrt(np.float32(residual[0])) / np.float32(residual[1]))
            cnt = np.float32(residual[1])
            if err < lastError and lastCount == cnt:
                break
            if np.float64(err) > np.float64(lastError) + 0.001:
                resultR = lastResultR
                break
            lastError, lastCount, lastResultR = err, cnt, resultR.copy()
            delta = np.linalg.solve(jtj.astype(np.float64), jtr.astype(np.float64)).astype(np.float32)
            R_lr = (rodrigues(delta.astype(np.float64)).astype(np.float32) @ R_lr).astype(np.float32)
            resultR = R_lr.astype(np.float64)
    iters = [3 if fastOdom else 10, 5 if pyramid else 0, 4 if pyramid else 0]
    Rprev_inv = np.linalg.inv(Rprev.astype(np.float64)).astype(np.float32)
    resultRt = np.eye(4)
    if so3:
        resultRt[:3, :3] = resultR
    for l in (2, 1, 0):
        cl = cam_level(cam, l)
        K = _K(cl)
        Kinv = np.linalg.inv(K)
        cloud = {tag: b.cloud(b.depths(l)[0], cl) for tag, b in backs} if rgb else {}
        lastRGBError = np.float32(np.finfo(np.float32).max)
        for j in range(iters[l]):
            Rt = np.linalg.inv(resultRt)
            krkinv = (K @ Rt[:3, :3] @ Kinv).astype(np.float32)
            kt = (K @ Rt[:3, 3]).astype(np.float32)
            rec = dict(kind="se3", level=l, it=j)
            sigma = rgbSize = 0
            corr = {}
            if rgb:
                minScale = float(np.float32(MIN_GRAD[l] ** 2 / SOBEL_SCALE ** 2))
                for tag, b in backs:
                    lastD, nextD = b.depths(l)
                    lastI, nextI, _ = b.images(l)
                    corr[tag], sg, ct = b.computeRgbResidual(minScale, sob[tag][l][0], sob[tag][l][1], lastD, nextD, lastI, nextI, MAX_DEPTH_DELTA, kt, krkinv)
                    rec[tag + "_sigma"], rec[tag + "_count"] = sg, ct
                    if tag == "d":
                        sigma, rgbSize = sg, ct
            with np.errstate(divide="ignore", invalid="ignore"):
                sigmaVal = np.float32(np.sqrt(np.float32(1 if (np.float32(sigma) / np.float32(rgbSize) == 0) else rgbSize)))
            rgbError = np.float32(np.sqrt(np.float32(sigma)) / np.float32(1 if rgbSize == 0 else rgbSize))
            if rgbOnly and rgbError > lastRGBError:
                log.append(rec)
                break
            lastRGBError = rgbError
            if rgbOnly:
                sigmaVal = np.float32(-1)
            A_icp = np.zeros((6, 6), np.float32); b_icp = np.zeros(6, np.float32)
            A_rgb = np.zeros((6, 6), np.float32); b_rgb = np.zeros(6, np.float32)
            if icp:
                for tag, b in backs:
                    A, bb, res, sums, _ = b.icpStep(Rcurr, tcurr, b.maps(l), Rprev_inv, tprev, cl, if_curvature_info)
                    rec[tag + "_icp_sums"], rec[tag + "_icp_res"] = np.array(sums), np.array(res)
                    if tag == "d":
                        A_icp, b_icp = A, bb
            if rgb:
                for tag, b in backs:
                    A, bb, sums = b.rgbStep(corr[tag], float(sigmaVal), cloud[tag], float(cl[0]), float(cl[1]), sob[tag][l][0], sob[tag][l][1], 0, SOBEL_SCALE)
                    rec[tag + "_rgb_sums"] = np.array(sums)
                    if tag == "d":
                        A_rgb, b_rgb = A, bb
            if icp and rgb:
                w = float(icpWeight)
                lastA = A_rgb.astype(np.float64) + w * w * A_icp.astype(np.float64)
                lastb = b_rgb.astype(np.float64) + w * b_icp.astype(np.float64)
            elif icp:
                lastA, lastb = A_icp.astype(np.float64), b_icp.astype(np.float64)
            else:
                lastA, lastb = A_rgb.astype(np.float64), b_rgb.astype(np.float64)
            result = np.linalg.solve(lastA, lastb)
            upd = np.eye(4)
            upd[:3, :3] = rodrigues(result[3:])
            upd[:3, 3] = result[:3]
            resultRt = upd @ resultRt
            Rf, tf = resultRt[:3, :3].astype(np.float32), resultRt[:3, 3].astype(np.float32)
            ti = -(Rf.T @ tf)
            Rcurr = (Rprev @ Rf.T).astype(np.float32)
            tcurr = (Rprev @ ti + tprev).astype(np.float32)
            rec["pose_t"], rec["pose_R"] = tcurr.copy(), Rcurr.copy()
            log.append(rec)
    if rgb and np.linalg.norm(tcurr - tprev) > 0.3:
        Rcurr, tcurr = Rprev, tprev
    return tcurr, Rcurr, log
