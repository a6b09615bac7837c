/*
 * oracle/orc.h -- CPU ORACLE for the HRBFFusion3D per-frame hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under hrbffusion3d_b200/ may include, link
 * or call this; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, as the checker.
 *
 * It is a plain-C restatement (no Eigen, no GL, no CUDA) of the reference
 * algorithm, written from the reference's behaviour; every function cites the
 * reference file:line (relative to /root/reference) it follows.
 *
 * PARITY PINNING
 *   rows 1-3 (icpStep / rgbStep / computeRgbResidual / so3Step): pinned against
 *     the reference's OWN CUDA kernels (Core/src/Cuda/reduce.cu compiled
 *     unmodified into oracle/_ref/, run on a B200; golden vectors in
 *     tests/golden/ref_reduce.npz, generator oracle/gen_ref_golden.py).
 *   rows 6-10, GlobalModel::initialise, FillIn, VertexConfidence (the GLSL passes):
 *     pinned against the reference's OWN SHADER TEXT, compiled for the CPU
 *     (oracle/build_ref_glsl.py + glsl_cpu.h -> oracle/_ref/libref_glsl.so) and run
 *     per fragment / vertex on the same inputs: tests/test_oracle_vs_reference_glsl.py.
 *     Bit-identical where the pass is not an exp / a root search, with the shaders'
 *     float-counter window loops taken literally (the default; orc_set_float_loops(0)
 *     = the intended integer windows round 1 restated -- DESIGN.md).
 *     Fixed-function GL between the shaders (point rasterisation, depth test,
 *     transform-feedback order, framebuffer formats) is restated, not executed.
 *   row 5 (cudafuncs.cu map / pyramid kernels): the reference's kernels are built
 *     (oracle/_ref/libref_cudafuncs.so, texture-reference shim), ran on a B200:
 *     golden vectors tests/golden/ref_cudafuncs.npz (oracle/gen_ref5_golden.py), live
 *     comparison in tests/test_oracle_vs_reference_row5.py.
 *   row 4 (host Gauss-Newton loop, RGBDOdometry.cpp + Eigen ldlt): Eigen is not
 *     installed and the class is inseparable from GL -> "parity unpinned", except the
 *     pose update (OdometryProvider.h rodrigues / computeUpdateSE3), which is pinned to
 *     the reference's own header compiled against oracle/eigen_mini
 *     (tests/test_oracle_vs_reference_host.py, bit-identical); the oracle's own outputs
 *     on the reference's GPUTest pair are the golden vectors.
 *
 * Layouts (identical to the reference so buffers are interchangeable):
 *   SoA map    : float[4*rows][cols], planes x,y,z,w stacked row-wise
 *                (RGBDOdometry.cpp:128-136); dense here (pitch == cols).
 *   AoS texture: float[rows][cols][4]  (GL RGBA32F as read by cudaMemcpyFromArray)
 *   surfel     : 5 x float4 = 80 B  {pos.xyz,conf | colour,submap,initTime,lastTime |
 *                n.xyz,radius | k1dir.xyz,k1 | k2dir.xyz,k2}  (Shaders/Vertex.cpp:20-44)
 */
#ifndef ORC_H_
#define ORC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NUM_PYRS 3

typedef struct { float fx, fy, cx, cy; } orc_cam;

/* Cuda/types.cuh:74-80 (DataTerm, 16 B) */
typedef struct { short zx, zy; short ox, oy; float diff; unsigned char valid; unsigned char pad[3]; } orc_dataterm;

/* ---------------------------------------------------------------- row 5 -- */
/* Cuda/cudafuncs.cu:344-383 */
void orc_copyMaps(int rows, int cols, const float* v_aos, const float* n_aos, float* vmap, float* nmap);
/* Cuda/cudafuncs.cu:405-431 */
void orc_copyCurvatureMap(int rows, int cols, const float* c_aos, float* cmap, float thr);
/* Cuda/cudafuncs.cu:452-470 */
void orc_copyicpWeightMap(int rows, int cols, const float* w_src, float* w_dst);
/* Cuda/cudafuncs.cu:526-587; normalize=1 -> resizeNMap, 0 -> resizeVMap.  dst is
 * pre-filled by the caller (stale planes of NaN pixels are left untouched,
 * as in the reference) */
void orc_resizeMap(int drows, int dcols, const float* src, float* dst, int normalize);
/* Cuda/cudafuncs.cu:618-674 */
void orc_resizeCMap(int drows, int dcols, const float* src, float* dst);
/* Cuda/cudafuncs.cu:694-726 */
void orc_resizeicpWeightMap(int drows, int dcols, const float* src, float* dst);
/* Cuda/cudafuncs.cu:213-257 (in place allowed) */
void orc_tranformMaps(int rows, int cols, const float* vsrc, const float* nsrc,
                      const float R[9], const float t[3], float* vdst, float* ndst);
/* Cuda/cudafuncs.cu:279-322 */
void orc_transformCurvMaps(int rows, int cols, const float* k1src, const float* k2src,
                           const float R[9], const float t[3], float* k1dst, float* k2dst);
/* Cuda/cudafuncs.cu:57-94 (pyrDown, sigma_color 30), :109-136, :154-195 */
void orc_pyrDownDepth(int srows, int scols, const float* src, float* dst);
void orc_createVMap(orc_cam k, int rows, int cols, const float* depth, float* vmap, float cutoff, float factor);
void orc_createNMap(int rows, int cols, const float* vmap, float* nmap);
/* RGB branch prep: cudafuncs.cu:874-885, 493-524, 818-848, 898-911, 930-954, 995-1013 */
void orc_verticesToDepth(int rows, int cols, const float* v_aos, float* depth, float cutoff);
void orc_pyrDownGaussF(int srows, int scols, const float* src, float* dst);
void orc_pyrDownUcharGauss(int srows, int scols, const unsigned char* src, unsigned char* dst);
void orc_rgbaToIntensity(int rows, int cols, const unsigned char* rgba, unsigned char* dst);
void orc_sobel(int rows, int cols, const unsigned char* src, short* dx, short* dy);
void orc_projectToPointCloud(int rows, int cols, const float* depth, float* cloud3, orc_cam k_level);

/* ------------------------------------------------------------ rows 1-3 -- */
typedef struct {
    int use_search;      /* registrationICPUseCoorespondenceSearch */
    int radius;          /* registrationICPNeighborSearchRadius    */
    int use_weight;      /* icp_if_use_weight                      */
    float dist_thres, angle_thres;
} orc_icp_opts;

/* Cuda/reduce.cu:253-693.  A[36] row-major symmetric, b[6], residual[2] =
 * {sum w r^2, inliers}.  sums29 (optional) receives the 29 raw sums in
 * JtJJtrSE3 field order (types.cuh:100-151).  corres (optional) int[rows*cols*2]. */
void orc_icpStep(int rows, int cols,
                 const float Rcurr[9], const float tcurr[3],
                 const float* vmap_curr, const float* nmap_curr,
                 const float* ck1_curr, const float* ck2_curr,
                 const float Rprev_inv[9], const float tprev[3], orc_cam intr,
                 const float* vmap_g_prev, const float* nmap_g_prev,
                 const float* ck1_g_prev, const float* ck2_g_prev,
                 const float* icpw_g_prev, const orc_icp_opts* o,
                 float A[36], float b[6], float residual[2], double* sums29, int* corres);

/* Cuda/reduce.cu:957-1154 */
void orc_computeRgbResidual(int rows, int cols, float minScale,
                            const short* dIdx, const short* dIdy,
                            const float* lastDepth, const float* nextDepth,
                            const unsigned char* lastImage, const unsigned char* nextImage,
                            orc_dataterm* corresImg, float maxDepthDelta,
                            const float kt[3], const float krkinv[9],
                            int* sigmaSum, int* count);
/* Cuda/reduce.cu:697-896 */
void orc_rgbStep(int rows, int cols, const orc_dataterm* corresImg, float sigma,
                 const float* cloud3, float fx, float fy,
                 const short* dIdx, const short* dIdy, int use_grad_weight, float sobelScale,
                 float A[36], float b[6], double* sums29);
/* Cuda/reduce.cu:1156-1359 */
void orc_so3Step(int rows, int cols, const unsigned char* lastImage, const unsigned char* nextImage,
                 const float imageBasis[9], const float kinv[9], const float krlr[9],
                 float A[9], float b[3], float residual[2], double* sums11);

/* -------------------------------------------------------------- row 4 -- */
typedef struct orc_odom orc_odom;
/* Utils/RGBDOdometry.cpp:35-154 */
orc_odom* orc_odom_create(int width, int height, float cx, float cy, float fx, float fy,
                          float distThresh, float angleThresh);
void orc_odom_destroy(orc_odom*);
/* RGBDOdometry.cpp:161-181 (GPUTest path), :183-206, :208-247 */
void orc_odom_initICP_depth(orc_odom*, const float* depth_raw_f32, float depthCutoff, float depthFactor);
void orc_odom_initICP(orc_odom*, const float* vert_aos, const float* norm_aos, float depthCutoff);
void orc_odom_initICPModel(orc_odom*, const float* vert_aos, const float* norm_aos, float depthCutoff, const float pose[16]);
/* :689-699 (rgba: RGBA8 bytes, R first) */
void orc_odom_initRGB(orc_odom*, const unsigned char* rgba);
void orc_odom_initRGBModel(orc_odom*, const unsigned char* rgba);
void orc_odom_initFirstRGB(orc_odom*, const unsigned char* rgba);
/* :701-759 */
void orc_odom_initCurvature(orc_odom*, const float* k1_aos, const float* k2_aos, float curvThr);
void orc_odom_initCurvatureModel(orc_odom*, const float* k1_aos, const float* k2_aos, const float pose[16], float curvThr);
/* :761-775 */
void orc_odom_initICPweight(orc_odom*, const float* w);
/* curvature planes = 0 (valid), weights = 1: the GPUTest-pair convention */
void orc_odom_fillNeutralCurvature(orc_odom*);

typedef struct {
    int rgbOnly; float icpWeight; int pyramid; int fastOdom; int so3; int if_curvature_info;
    int use_search; int search_radius; int rgb_grad_weight;
} orc_track_opts;
typedef struct {
    float lastICPError, lastICPCount, lastRGBError, lastRGBCount, lastSO3Error, lastSO3Count;
    double lastA[36], lastb[6];
    int icp_iterations_run;
} orc_track_stats;
/* RGBDOdometry.cpp:796-1249.  trans[3], rot[9] row-major: in = previous pose, out = new pose */
void orc_odom_getIncrementalTransformation(orc_odom*, float trans[3], float rot[9],
                                           const orc_track_opts*, orc_track_stats*);
/* test access to pyramid level maps: which = 0..8 -> vmap_g_prev,nmap_g_prev,ck1_g_prev,ck2_g_prev,
 * vmap_curr,nmap_curr,ck1_curr,ck2_curr,icpWeight */
const float* orc_odom_map(const orc_odom*, int which, int level);
const unsigned char* orc_odom_image(const orc_odom*, int which, int level); /* 0 last,1 next,2 lastNext */
const float* orc_odom_depth(const orc_odom*, int which, int level);         /* 0 last,1 next */

/* Utils/OdometryProvider.h:35-93 and the fp64 6x6 / fp32 3x3 LDLT standing in for Eigen's ldlt() */
void orc_rodrigues(const double w[3], double R[9]);
void orc_ldlt_solve6(const double A[36], const double b[6], double x[6]);
void orc_computeUpdateSE3(double resultRt[16], const double result[6], float iso16[16]);   /* OdometryProvider.h:71-93 */
void orc_ldlt_solve3f(const float A[9], const float b[3], float x[3]);

/* ---------------------------------------------------------- rows 6-7 -- */
typedef struct {
    float cx, cy, fx, fy; int cols, rows;
    float maxDepth;
} orc_splat_params;
/* IndexMap.cpp:193-267 + Shaders/index_map.vert:34-66, .frag.  active_kf: float[kf_dim] 0/1 mask.
 * Outputs: index u32[rows*cols] (0 = empty, as glClear), 5 x RGBA32F AoS maps (camera frame).
 * Depth test: nearest z wins; ties -> lowest surfel id (GL_LESS + in-order rasterisation). */
void orc_predictIndices(const float pose[16], const float* surfels, int count,
                        const orc_splat_params* p, const float* active_kf, int kf_dim,
                        uint32_t* index, float* vertConf, float* colorTime, float* normRad,
                        float* curvMax, float* curvMin);

typedef struct {
    float cx, cy, fx, fy; int cols, rows;
    int win; int minNeighbors; int maxNeighbors; float confThreshold; float icpWeightLambda;
} orc_predict_params;
/* IndexMap.cpp:413-518 + Shaders/predict_hrbf.frag:40-311 + hrbfbase.glsl:7-166.
 * Outputs (AoS): image RGBA8, vertex(xyz,conf), normal(xyz,radius), curvMax, curvMin (RGBA32F),
 * time u16, icp_weight f32. */
void orc_predictHRBF(const orc_predict_params* p,
                     const float* vertConf, const float* colorTime, const float* normRad,
                     const float* curvMax, const float* curvMin,
                     unsigned char* image, float* vertex, float* normal,
                     float* ocurvMax, float* ocurvMin, unsigned short* time, float* icp_weight);

/* ---------------------------------------------------------- rows 8-9 -- */
typedef struct {
    float cx, cy, fx, fy; int cols, rows;
    float maxDepth; float confThreshold; float radiusMultiplier; float curvThr; int pca;
    int cleanWindow;
} orc_model_params;
/* GlobalModel.cpp:214-288 + Shaders/init_unstableTex.vert:51-89, .geom.  Returns count.
 * Pixel order = uv VBO order: x outer, y inner (GlobalModel.cpp:89-96). */
int orc_model_initialise(const orc_model_params* p, const float pose[16],
                         const float* vertexRaw, const float* normal, const unsigned char* rgb /* RGB8 */,
                         const float* curv1, const float* curv2, const float* gradientMag, int useConfEval, float epsilon,
                         float* surfels_out);
/* GlobalModel.cpp:355-549 + Shaders/data.vert:63-198, data.geom, data.frag, update.vert:51-115.
 * surfels_in[count] -> surfels_out[count] (ping-pong), unstable_out[<= rows*cols] ; returns n_unstable */
int orc_model_fuse(const orc_model_params* p, const float pose[16], int time,
                   const unsigned char* rgb, const float* depthRaw, const float* depthFiltered,
                   const float* curv1, const float* curv2, const float* confidence,
                   const uint32_t* index, const float* vertConf, const float* colorTime, const float* normRad,
                   float indexSubmap,
                   const float* surfels_in, int count, float* surfels_out, float* unstable_out);
/* GlobalModel.cpp:551-688 + Shaders/copy_unstable.vert:62-166, .geom:37-50.  Returns new count. */
int orc_model_clean(const orc_model_params* p, const float pose[16], int time,
                    const uint32_t* index, const float* vertConf, const float* colorTime, const float* normRad,
                    const float* active_kf, int kf_dim,
                    const float* surfels_in, int count, const float* unstable, int n_unstable,
                    float* surfels_out);

/* ------------------------------------------------------------ row 10 -- */
typedef struct {
    float cx, cy, fx, fy; int cols, rows;
    float depthFactor;      /* metres per raw unit (1/5000 for TUM) */
    float maxD;             /* globalDepthCutoff */
    float radiusMultiplier; int pca; float curvWindow; int bilateral;
} orc_prep_params;
/* Shaders/depth_bilateral.frag ; orc_exp_bilateral: the deterministic exp() this pass is defined with */
float orc_exp_bilateral(float x);
void orc_filterDepth(const orc_prep_params* p, const unsigned short* raw, float* filtered);
/* Shaders/depth_metric_raw.frag, depth_metric_filtered.frag */
void orc_metriciseDepth(const orc_prep_params* p, const unsigned short* raw, const float* filtered,
                        float* metric, float* metric_filtered);
/* geometry.glsl:190-244 (getNormalPCA, window 3) at pixel (px,py); surfels.glsl:19-34, 37-46 */
/* 0 (default): intended integer windows; 1: the shaders' literal float-counter window loops (see orc_prep.c) */
void orc_set_float_loops(int on);
int orc_get_float_loops(void);
int orc_texel_of(float u, int n);      /* GL_NEAREST texel of texture coordinate u on an axis of n texels (8 fractional bits of fixed point) */
int orc_float_window(int p, int n, float win, int uv, int* texels /*[16]*/, float* coords /*[16]*/);   /* test hook */
void orc_set_uv_vbo_coords(int on);      /* per thread; used by orc_model_fuse around its PCA normal (literal mode only) */
void orc_getNormalPCA(const orc_prep_params* p, const float* depth, int px, int py, float vz, float n[3]);
float orc_getRadius(float icx, float icy, float depth, float norm_z);
float orc_confidence(float cx, float cy, float x, float y, float max_dist, float w);
/* Shaders/depth_vertex_normal_radius.frag:23-68, geometry.glsl:190-244, surfels.glsl:19-34 */
void orc_computeVertexNormalRadius(const orc_prep_params* p, const float* metric, const float* metric_filtered,
                                   float* vertex_raw, float* vertex_filtered, float* normal, float* radius);
/* Shaders/depth_curvature_gradient.frag:28-142 */
void orc_computeCurvatureGradient(const orc_prep_params* p, const float* vertex_filtered, const float* normal,
                                  float* curv1, float* curv2, float* gradient_mag, float* normal_opt);
/* Shaders/depth_confidence_evaluation.frag (useConfidenceEvaluation honoured) */
void orc_vertexConfidence(const orc_prep_params* p, const float* gradient_mag, float weighting,
                          int useConfEval, float epsilon, float* confidence);
/* Shaders/fill_vertex.frag, fill_normal.frag, fill_curvature.frag, fill_rgb.frag (FillIn.cpp) */
void orc_fillIn(const orc_prep_params* p, int passthrough, float lambda, float curvThr,
                const float* eVertex, const float* eIcpW, const float* eNormal,
                const float* eK1, const float* eK2, const unsigned char* eImage,
                const float* vertexFiltered, const float* normal, const float* k1, const float* k2,
                const float* confidence, const unsigned char* rgb,
                float* oVertex, float* oIcpW, float* oNormal, float* oK1, float* oK2, unsigned char* oImage);
/* HRBFFusion.cpp:974-988 + Shaders/Resize.cpp (1/20 nearest sample at texel centres) */
int orc_denseEnough(int rows, int cols, const float* vertex, float thresh);

/* GlobalModel::updateModel, GlobalModel.cpp:690-767 + Shaders/update_delta_trans.vert:41-91 (in place) */
void orc_model_update(float* surfels, int count, const float* delta, int n_delta);

#ifdef __cplusplus
}
#endif
#endif
