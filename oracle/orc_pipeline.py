"""CPU ORACLE of the per-frame orchestration (TEST INFRASTRUCTURE ONLY): HRBFFusion::processFrame / predict
(Core/src/HRBFFusion.cpp:991-1260) restated on top of oracle/orc_py.py, with the sparse-SLAM back-end
(ORB keyframes, BA, loop closure: SURVEY section 2 rows 10-11, out of scope) switched off as in
optimizationUseLocalBA = optimizationUseGlobalBA = false."""
import numpy as np

from . import orc_py as orc


def rodrigues2_norm(R):
    """|log(R)| as HRBFFusion::rodrigues2 returns it (rotation angle)"""
    c = (np.trace(R.astype(np.float64)) - 1.0) * 0.5
    return float(np.arccos(np.clip(c, -1.0, 1.0)))


class HRBFFusion:
    def __init__(self, width, height, cam, depthFactor=1.0 / 5000.0, depthCutoff=3.5, confidence=5.0, icpWeight=10.0,
                 so3=True, pyramid=True, fastOdom=False, rgbOnly=False, weightedICP=True,
                 win=3, minNeighbors=6, maxNeighbors=10, predConfThreshold=3.0, icpWeightLambda=10.0):
        self.W, self.H, self.cam = width, height, cam
        self.pp = orc.prep_params(cam, width, height, depthFactor, depthCutoff)
        self.mp = orc.model_params(cam, width, height, 20.0, confidence)
        self.maxDepthProcessed = 20.0
        self.kw = dict(rgbOnly=rgbOnly, icpWeight=icpWeight, pyramid=pyramid, fastOdom=fastOdom, so3=so3, if_curvature_info=weightedICP)
        self.pred_kw = dict(win=win, minNeighbors=minNeighbors, maxNeighbors=maxNeighbors, confThreshold=predConfThreshold,
                            icpWeightLambda=icpWeightLambda)
        self.lamb = icpWeightLambda
        self.odom = orc.Odometry(width, height, cam[2], cam[3], cam[0], cam[1])
        self.tick = 1
        self.currPose = np.eye(4, dtype=np.float32)
        self.surfels = np.zeros((0, 20), np.float32)
        self.pred = None
        self.fill = None
        self.trajectory = []
        self.indexSubmap = 0
        self.last = {}

    @staticmethod
    def rgba(rgb):
        return np.ascontiguousarray(np.concatenate([rgb, np.full(rgb.shape[:2] + (1,), 255, np.uint8)], -1))

    def processFrame(self, rgb, depth, weightMultiplier=1.0):
        fr = orc.preprocess(self.pp, depth)
        weighting = 1.0
        if self.tick == 1:
            self.surfels = orc.modelInitialise(self.mp, self.currPose, fr, rgb)
            self.odom.initFirstRGB(self.rgba(rgb))
        else:
            lastPose = self.currPose.copy()
            fill = not orc.denseEnough(self.pred["vertex"])
            src = self.fill if fill else self.pred
            self.odom.initICPModel(src["vertex"], src["normal"], self.maxDepthProcessed, self.currPose)
            self.odom.initRGBModel(src["image"])
            self.odom.initCurvatureModel(src["curvk1"], src["curvk2"], self.currPose)
            self.odom.initICP(fr["vertex_filtered"], fr["normal"], self.maxDepthProcessed)
            self.odom.initRGB(self.rgba(rgb))
            self.odom.initCurvature(fr["curv1"], fr["curv2"])
            self.odom.initICPweight(src["icpw"])
            t, R, st = self.odom.getIncrementalTransformation(self.currPose[:3, 3], self.currPose[:3, :3], **self.kw)
            self.currPose[:3, 3] = t
            self.currPose[:3, :3] = R
            self.last["stats"] = st
            self.last["filled"] = fill
            # weight by velocity (HRBFFusion.cpp:1112-1123)
            diff = np.linalg.inv(self.currPose.astype(np.float64)) @ lastPose.astype(np.float64)
            w = max(float(np.linalg.norm(diff[:3, 3])), rodrigues2_norm(diff[:3, :3]))
            w = min(w, 0.01)
            weighting = max(1.0 - w / 0.01, 0.5) * weightMultiplier
        # VertexConfidence runs only from the second frame on (HRBFFusion.cpp:1126 sits in the tick > 1 branch);
        # on the first frame the CONFIDENCE texture still holds its initial contents, defined as zeros
        conf = orc.vertexConfidence(self.pp, fr["gradient_mag"], weighting) if self.tick > 1 else np.zeros((self.H, self.W), np.float32)
        if self.tick > 1 and not self.kw["rgbOnly"]:
            idx = orc.predictIndices(self.currPose, self.surfels, self.cam, self.W, self.H, self.maxDepthProcessed)
            self.surfels, unstable = orc.modelFuse(self.mp, self.currPose, self.tick, rgb, fr, conf, idx, self.indexSubmap, self.surfels)
            idx = orc.predictIndices(self.currPose, self.surfels, self.cam, self.W, self.H, self.maxDepthProcessed)
            self.surfels = orc.modelClean(self.mp, self.currPose, self.tick, idx, self.surfels, unstable)
        self.trajectory.append(self.currPose.copy())
        # predict (HRBFFusion.cpp:1244-1260)
        idx = orc.predictIndices(self.currPose, self.surfels, self.cam, self.W, self.H, self.maxDepthProcessed)
        self.pred = orc.predictHRBF(idx, self.cam, self.W, self.H, **self.pred_kw)
        self.fill = orc.fillIn(self.pp, self.pred, fr, conf, rgb, 0, self.lamb)
        self.last.update(frame=fr, conf=conf, idx=idx, weighting=weighting)
        self.tick += 1
        return self.currPose.copy()
