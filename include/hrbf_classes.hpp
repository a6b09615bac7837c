/*
 * hrbf_classes.hpp -- GL-free C++ classes with the reference's CLASS and METHOD names and argument order for the hot path, over the
 * C ABI of hrbf_b200.h.  Header-only.  A maintainer of the reference replaces, inside Core/src/HRBFFusion.cpp, the members
 *     RGBDOdometry frameToModel;  IndexMap indexMap;  GlobalModel* globalModel;  FillIn fillIn;  std::map<std::string, GPUTexture*> textures
 * by the classes below (namespace hrbf_b200) and keeps the call sites of HRBFFusion::processFrame / predict
 * (Core/src/HRBFFusion.cpp:1016-1021, 1043-1052, 1069-1100, 1126, 1195-1227, 1244-1260) as they are: tests/test_abi.py compiles and
 * RUNS a translation unit that makes exactly those calls in that order (tests/classes_tu.cpp).
 *
 *   reference type                         here
 *   GPUTexture* (GL texture + CUDA handle) hrbf_b200::GPUTexture : { device pointer, width, height, format } -- a VIEW, not an owner
 *   Eigen::Matrix4f pose                   hrbf_b200::Mat4 (row-major float[16]); any type with operator()(i, j) converts through Mat4::from
 *   std::pair<GLuint, GLuint> model()      hrbf_b200::ModelRef { const float* surfels (device, 80-B records), unsigned int count }
 *   cudaSafeCall -> exit(0)                std::runtime_error carrying hrbf_last_error()
 * All work is enqueued on the CUDA stream given to the constructors (default: the legacy default stream, like the reference).
 */
#ifndef HRBF_CLASSES_HPP_
#define HRBF_CLASSES_HPP_

#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "hrbf_b200.h"

namespace hrbf_b200 {

inline void check(int rc, const char* what)
{
    if (rc != HRBF_OK) throw std::runtime_error(std::string(what) + ": " + hrbf_last_error());
}

/* row-major 4x4, the element order of hrbf_b200.h's pose16 arguments */
struct Mat4 {
    float m[16];
    Mat4() { for (int k = 0; k < 16; ++k) m[k] = (k % 5 == 0) ? 1.f : 0.f; }
    float& operator()(int i, int j) { return m[i * 4 + j]; }
    float operator()(int i, int j) const { return m[i * 4 + j]; }
    template <class M>
    static Mat4 from(const M& e) { Mat4 r; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r(i, j) = e(i, j); return r; }      // e.g. Eigen::Matrix4f
    const float* data() const { return m; }
};

/* Core/src/GPUTexture.h:27-59 without GL: a typed view of a dense device buffer */
struct GPUTexture {
    enum Format { RGBA32F, R32F, RGBA8, RGB8, R16UI, R32UI };
    void* dev = nullptr;
    int width = 0, height = 0;
    Format format = RGBA32F;
    GPUTexture() {}
    GPUTexture(void* d, int w, int h, Format f) : dev(d), width(w), height(h), format(f) {}
    const float* f32() const { return static_cast<const float*>(dev); }
    const unsigned char* u8() const { return static_cast<const unsigned char*>(dev); }
    const unsigned int* u32() const { return static_cast<const unsigned int*>(dev); }
};

struct ModelRef { const float* surfels; unsigned int count; };

/* Core/src/Utils/RGBDOdometry.h:57-107 */
class RGBDOdometry {
public:
    RGBDOdometry(int width, int height, float cx, float cy, float fx, float fy, float distThresh = 0.10f, float angleThresh = 0.3420201433f /* sin 20 deg */,
                 void* stream = nullptr)
        : stream_(stream)
    {
        check(hrbf_odometry_create(&h_, width, height, cx, cy, fx, fy, distThresh, angleThresh), "RGBDOdometry");
    }
    virtual ~RGBDOdometry() { hrbf_odometry_destroy(h_); }
    RGBDOdometry(const RGBDOdometry&) = delete;
    RGBDOdometry& operator=(const RGBDOdometry&) = delete;

    void initICP(GPUTexture* filteredDepth, const float depthCutoff, const float mDepthMapFactor)
    { check(hrbf_odometry_init_icp_depth(h_, filteredDepth->f32(), depthCutoff, mDepthMapFactor, stream_), "initICP(depth)"); }
    void initICP(GPUTexture* predictedVertices, GPUTexture* predictedNormals, const float depthCutoff)
    { check(hrbf_odometry_init_icp(h_, predictedVertices->f32(), predictedNormals->f32(), depthCutoff, stream_), "initICP"); }
    void initICPModel(GPUTexture* predictedVertices, GPUTexture* predictedNormals, const float depthCutoff, const Mat4& modelPose)
    { check(hrbf_odometry_init_icp_model(h_, predictedVertices->f32(), predictedNormals->f32(), depthCutoff, modelPose.data(), stream_), "initICPModel"); }
    void initRGB(GPUTexture* rgb) { check(hrbf_odometry_init_rgb(h_, rgb->u8(), stream_), "initRGB"); }
    void initRGBModel(GPUTexture* rgb) { check(hrbf_odometry_init_rgb_model(h_, rgb->u8(), stream_), "initRGBModel"); }
    void initFirstRGB(GPUTexture* rgb) { check(hrbf_odometry_init_first_rgb(h_, rgb->u8(), stream_), "initFirstRGB"); }
    void initCurvature(GPUTexture* curvk1, GPUTexture* curvk2) { check(hrbf_odometry_init_curvature(h_, curvk1->f32(), curvk2->f32(), stream_), "initCurvature"); }
    void initCurvatureModel(GPUTexture* curvk1Model, GPUTexture* curvk2Model, const Mat4& modelPose)
    { check(hrbf_odometry_init_curvature_model(h_, curvk1Model->f32(), curvk2Model->f32(), modelPose.data(), stream_), "initCurvatureModel"); }
    void initICPweight(GPUTexture* icpWeight) { check(hrbf_odometry_init_icp_weight(h_, icpWeight->f32(), stream_), "initICPweight"); }

    /* trans: float[3] (Eigen::Vector3f::data()), rot: row-major float[9] (Eigen::Matrix<float, 3, 3, RowMajor>::data()); both in/out */
    void getIncrementalTransformation(float* trans, float* rot, const bool& rgbOnly, const float& icpWeight, const bool& pyramid, const bool& fastOdom,
                                      const bool& so3, const bool& if_curvature_info, const int index_frame)
    {
        hrbf_track_stats st;
        check(hrbf_odometry_get_incremental_transformation(h_, trans, rot, rgbOnly, icpWeight, pyramid, fastOdom, so3, if_curvature_info, index_frame, &st, stream_),
              "getIncrementalTransformation");
        lastICPError = st.lastICPError; lastICPCount = st.lastICPCount; lastRGBError = st.lastRGBError; lastRGBCount = st.lastRGBCount;
        lastSO3Error = st.lastSO3Error; lastSO3Count = st.lastSO3Count;
        for (int k = 0; k < 36; ++k) lastA[k] = st.lastA[k];
        for (int k = 0; k < 6; ++k) lastb[k] = st.lastb[k];
    }
    /* any vector / matrix types with data() (Eigen::Vector3f, Eigen::Matrix<float, 3, 3, Eigen::RowMajor>): the reference's own signature */
    template <class Vec3, class Mat3>
    void getIncrementalTransformation(Vec3& trans, Mat3& rot, const bool& rgbOnly, const float& icpWeight, const bool& pyramid, const bool& fastOdom,
                                      const bool& so3, const bool& if_curvature_info, const int index_frame)
    { getIncrementalTransformation(trans.data(), rot.data(), rgbOnly, icpWeight, pyramid, fastOdom, so3, if_curvature_info, index_frame); }

    float lastICPError = 0, lastICPCount = 0, lastRGBError = 0, lastRGBCount = 0, lastSO3Error = 0, lastSO3Count = 0;      // RGBDOdometry.h:124-134
    double lastA[36] = {}, lastb[6] = {};
    hrbf_odometry* handle() { return h_; }

private:
    hrbf_odometry* h_ = nullptr;
    void* stream_;
};

/* Core/src/IndexMap.h:36-201 */
class IndexMap {
public:
    enum Prediction { ACTIVE, INACTIVE };
    IndexMap(int width, int height, float cx, float cy, float fx, float fy, void* stream = nullptr) : w_(width), h2_(height), stream_(stream)
    {
        check(hrbf_indexmap_create(&h_, width, height, cx, cy, fx, fy), "IndexMap");
        for (int k = 0; k < HRBF_TEX_COUNT; ++k) tex_[k] = GPUTexture(hrbf_indexmap_texture(h_, k), width, height, format_of(k));
    }
    virtual ~IndexMap() { hrbf_indexmap_destroy(h_); }
    IndexMap(const IndexMap&) = delete;
    IndexMap& operator=(const IndexMap&) = delete;

    void predictIndices(const Mat4& pose, const int& time, const int maxTime, const ModelRef& model, const float depthCutoff, const int insertSubmap, const int indexSubmap)
    { check(hrbf_indexmap_predict_indices(h_, pose.data(), time, maxTime, model.surfels, model.count, depthCutoff, insertSubmap, indexSubmap, stream_), "predictIndices"); }
    /* the GlobalStateParam knobs the reference reads inside predictHRBF (IndexMap.cpp:449-470) are members here, reference defaults */
    void predictHRBF(IndexMap::Prediction predictionType)
    { check(hrbf_indexmap_predict_hrbf(h_, predictionType == ACTIVE ? 0 : 1, preictionWindowMultiplier, preictionMinNeighbors, preictionMaxNeighbors,
                                       preictionConfThreshold, registrationICPCurvWeightImpactControl, stream_), "predictHRBF"); }
    void setActiveKeyframes(const std::vector<int>& lActiveKFID) { check(hrbf_indexmap_set_active_keyframes(h_, lActiveKFID.data(), (int)lActiveKFID.size(), stream_), "lActiveKFID"); }

    GPUTexture* indexTex() { return &tex_[HRBF_TEX_INDEX]; }
    GPUTexture* vertConfTex() { return &tex_[HRBF_TEX_VERTCONF]; }
    GPUTexture* colorTimeTex() { return &tex_[HRBF_TEX_COLORTIME]; }
    GPUTexture* normalRadTex() { return &tex_[HRBF_TEX_NORMRAD]; }
    GPUTexture* curvMaxTex() { return &tex_[HRBF_TEX_CURVMAX]; }
    GPUTexture* curvMinTex() { return &tex_[HRBF_TEX_CURVMIN]; }
    GPUTexture* depthTex() { return &depth_; }      /* never rendered in the reference either (no caller of synthesizeDepth): an empty view */
    GPUTexture* imageTexHRBF() { return &tex_[HRBF_TEX_IMAGE_HRBF]; }
    GPUTexture* vertexTexHRBF() { return &tex_[HRBF_TEX_VERTEX_HRBF]; }
    GPUTexture* normalTexHRBF() { return &tex_[HRBF_TEX_NORMAL_HRBF]; }
    GPUTexture* curvk1TexHRBF() { return &tex_[HRBF_TEX_CURVK1_HRBF]; }
    GPUTexture* curvk2TexHRBF() { return &tex_[HRBF_TEX_CURVK2_HRBF]; }
    GPUTexture* icpweightTexHRBF() { return &tex_[HRBF_TEX_ICPW_HRBF]; }
    GPUTexture* oldImageTexHRBF() { return &tex_[HRBF_TEX_OLD_IMAGE_HRBF]; }
    GPUTexture* oldVertexTexHRBF() { return &tex_[HRBF_TEX_OLD_VERTEX_HRBF]; }
    GPUTexture* oldNormalTexHRBF() { return &tex_[HRBF_TEX_OLD_NORMAL_HRBF]; }

    int preictionWindowMultiplier = 3, preictionMinNeighbors = 6, preictionMaxNeighbors = 10;      // [sic] GUI/GlobalStateParam.txt
    float preictionConfThreshold = 3.f, registrationICPCurvWeightImpactControl = 10.f;
    hrbf_indexmap* handle() { return h_; }

private:
    static GPUTexture::Format format_of(int k)
    {
        if (k == HRBF_TEX_INDEX) return GPUTexture::R32UI;
        if (k == HRBF_TEX_IMAGE_HRBF || k == HRBF_TEX_OLD_IMAGE_HRBF) return GPUTexture::RGBA8;
        if (k == HRBF_TEX_TIME_HRBF || k == HRBF_TEX_OLD_TIME_HRBF) return GPUTexture::R16UI;
        if (k == HRBF_TEX_ICPW_HRBF || k == HRBF_TEX_OLD_ICPW_HRBF) return GPUTexture::R32F;
        return GPUTexture::RGBA32F;
    }
    hrbf_indexmap* h_ = nullptr;
    int w_, h2_;
    void* stream_;
    GPUTexture tex_[HRBF_TEX_COUNT];
    GPUTexture depth_;
};

/* Core/src/GlobalModel.h:35-150 */
class GlobalModel {
public:
    GlobalModel(int width, int height, float cx, float cy, float fx, float fy, unsigned int capacity = 0 /* reference: 4596^2 */, void* stream = nullptr)
        : stream_(stream)
    { check(hrbf_model_create(&h_, width, height, cx, cy, fx, fy, capacity), "GlobalModel"); }
    virtual ~GlobalModel() { hrbf_model_destroy(h_); }
    GlobalModel(const GlobalModel&) = delete;
    GlobalModel& operator=(const GlobalModel&) = delete;

    void initialise(GPUTexture* vertexMap, GPUTexture* normalMap, GPUTexture* colorMap /* RGB8 */, GPUTexture* curv1Map, GPUTexture* curv2Map, GPUTexture* gradientMagMap,
                    const Mat4& init_pose)
    { check(hrbf_model_initialise(h_, vertexMap->f32(), normalMap->f32(), colorMap->u8(), curv1Map->f32(), curv2Map->f32(), gradientMagMap->f32(), init_pose.data(), stream_),
            "initialise"); }
    ModelRef model() { return ModelRef{ hrbf_model_model(h_), lastCount() }; }
    void fuse(const Mat4& pose, const int& time, GPUTexture* rgb /* RGB8 */, GPUTexture* depthRaw, GPUTexture* depthFiltered, GPUTexture* curv_map_max, GPUTexture* curv_map_min,
              GPUTexture* confidence, GPUTexture* indexMap, GPUTexture* vertConfMap, GPUTexture* colorTimeMap, GPUTexture* normRadMap, const float depthCutoff,
              const float confThreshold, const float weighting, const bool insertSubmap, const float indexsubmap)
    {
        check(hrbf_model_fuse(h_, pose.data(), time, rgb->u8(), depthRaw->f32(), depthFiltered->f32(), curv_map_max->f32(), curv_map_min->f32(), confidence->f32(),
                              indexMap->u32(), vertConfMap->f32(), colorTimeMap->f32(), normRadMap->f32(), depthCutoff, confThreshold, weighting, insertSubmap ? 1 : 0,
                              (int)indexsubmap, stream_), "fuse");
    }
    void clean(const Mat4& pose, const int& time, GPUTexture* indexMap, GPUTexture* vertConfMap, GPUTexture* colorTimeMap, GPUTexture* normRadMap, GPUTexture* depthMap,
               const float confThreshold, const float maxDepth)
    {
        check(hrbf_model_clean(h_, pose.data(), time, indexMap->u32(), vertConfMap->f32(), colorTimeMap->f32(), normRadMap->f32(), depthMap ? depthMap->f32() : nullptr,
                               confThreshold, maxDepth, stream_), "clean");
    }
    /* GlobalModel::updateModel (GlobalModel.cpp:690-767): DeltaTransformKF is a member of the reference class, an argument here */
    void updateModel(const std::vector<Mat4>& DeltaTransformKF)
    { check(hrbf_model_update_model(h_, DeltaTransformKF.empty() ? nullptr : DeltaTransformKF[0].data(), (int)DeltaTransformKF.size(), stream_), "updateModel"); }
    unsigned int lastCount()
    {
        unsigned int n = 0;
        check(hrbf_model_last_count(h_, &n, stream_), "lastCount");
        return n;
    }
    /* GlobalModel::downloadMap (GlobalModel.cpp:775-804): count x 5 float4, caller owns the vector */
    std::vector<float> downloadMap()
    {
        const unsigned int n = lastCount();
        std::vector<float> out((size_t)n * 20);
        unsigned int got = 0;
        check(hrbf_model_download_map(h_, out.data(), n, &got, stream_), "downloadMap");
        out.resize((size_t)got * 20);
        return out;
    }
    void setActiveKeyframes(const std::vector<int>& lActiveKFID) { check(hrbf_model_set_active_keyframes(h_, lActiveKFID.data(), (int)lActiveKFID.size(), stream_), "lActiveKFID"); }
    hrbf_model* handle() { return h_; }

private:
    hrbf_model* h_ = nullptr;
    void* stream_;
};

/* the `textures[GPUTexture::...]` map of HRBFFusion plus filterDepth / metriciseDepth / computeVertexNormalRadius / computeCurvatureGradient /
 * updateNormalRad / VertexConfidence (Core/src/HRBFFusion.cpp:1006-1021, 1262-1346) */
class FrameTextures {
public:
    FrameTextures(int width, int height, float cx, float cy, float fx, float fy, float depthFactor = 1.0f / 5000.0f, float depthCutoff = 3.5f, void* stream = nullptr)
        : stream_(stream)
    {
        hrbf_frame_params p{};
        p.width = width; p.height = height; p.cx = cx; p.cy = cy; p.fx = fx; p.fy = fy; p.depthFactor = depthFactor; p.depthCutoff = depthCutoff;
        p.radiusMultiplier = 4.f; p.normalPCA = 1; p.curvWindow = 3; p.bilateral = 1; p.useConfEval = 0; p.confEvalEpsilon = 1000.f;
        check(hrbf_frame_create(&h_, &p), "FrameTextures");
        static const GPUTexture::Format fmt[HRBF_FT_COUNT] = { GPUTexture::RGB8, GPUTexture::RGBA8, GPUTexture::R16UI, GPUTexture::R32F, GPUTexture::R32F, GPUTexture::R32F,
                                                               GPUTexture::RGBA32F, GPUTexture::RGBA32F, GPUTexture::RGBA32F, GPUTexture::RGBA32F, GPUTexture::RGBA32F,
                                                               GPUTexture::RGBA32F, GPUTexture::R32F, GPUTexture::R32F, GPUTexture::R32F };
        for (int k = 0; k < HRBF_FT_COUNT; ++k) tex_[k] = GPUTexture(hrbf_frame_texture(h_, k), width, height, fmt[k]);
    }
    ~FrameTextures() { hrbf_frame_destroy(h_); }
    FrameTextures(const FrameTextures&) = delete;
    FrameTextures& operator=(const FrameTextures&) = delete;
    /* textures[DEPTH_RAW]->texture->Upload(depth, ...), textures[RGB]->texture->Upload(rgb, ...): host buffers */
    void Upload(const unsigned char* rgb, const unsigned short* depth) { check(hrbf_frame_upload(h_, rgb, depth, 1, stream_), "Upload"); }
    void preprocess() { check(hrbf_frame_preprocess(h_, stream_), "filterDepth..updateNormalRad"); }
    void VertexConfidence(float weighting) { check(hrbf_frame_vertex_confidence(h_, weighting, stream_), "VertexConfidence"); }
    GPUTexture* operator[](int which) { return &tex_[which]; }      /* which: enum hrbf_frame_tex */
    hrbf_frame* handle() { return h_; }

private:
    hrbf_frame* h_ = nullptr;
    void* stream_;
    GPUTexture tex_[HRBF_FT_COUNT];
};

/* Core/src/Shaders/FillIn.h: the four fill passes, fused */
class FillIn {
public:
    FillIn(int width, int height, void* stream = nullptr) : stream_(stream)
    {
        check(hrbf_fillin_create(&h_, width, height), "FillIn");
        static const GPUTexture::Format fmt[HRBF_FILL_COUNT] = { GPUTexture::RGBA8, GPUTexture::RGBA32F, GPUTexture::RGBA32F, GPUTexture::RGBA32F, GPUTexture::RGBA32F, GPUTexture::R32F };
        GPUTexture* t[HRBF_FILL_COUNT] = { &imageTexture, &vertexTexture, &normalTexture, &curvk1Texture, &curvk2Texture, &icpweightTexture };
        for (int k = 0; k < HRBF_FILL_COUNT; ++k) *t[k] = GPUTexture(hrbf_fillin_texture(h_, k), width, height, fmt[k]);
    }
    ~FillIn() { hrbf_fillin_destroy(h_); }
    FillIn(const FillIn&) = delete;
    FillIn& operator=(const FillIn&) = delete;
    /* fillIn.vertex + normal + curvature + image of HRBFFusion::predict (HRBFFusion.cpp:1252-1259) in one call */
    void run(IndexMap& prediction, FrameTextures& frame, bool passthrough, float lambda = 10.f, float curvThr = 300.f)
    { check(hrbf_fillin_run(h_, prediction.handle(), frame.handle(), passthrough ? 1 : 0, lambda, curvThr, stream_), "FillIn"); }
    GPUTexture imageTexture, vertexTexture, normalTexture, curvk1Texture, curvk2Texture, icpweightTexture;      // FillIn.h member names

private:
    hrbf_fillin* h_ = nullptr;
    void* stream_;
};

/* HRBFFusion::denseEnough(resize.vertex(...)), HRBFFusion.cpp:974-987, 1069-1070 */
inline bool denseEnough(GPUTexture* vertexTexHRBF, float thresh = 0.75f, void* stream = nullptr)
{
    int dense = 0;
    check(hrbf_dense_enough(vertexTexHRBF->f32(), vertexTexHRBF->width, vertexTexHRBF->height, thresh, &dense, stream), "denseEnough");
    return dense != 0;
}

}  // namespace hrbf_b200

#endif /* HRBF_CLASSES_HPP_ */
