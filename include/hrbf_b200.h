/*
 * hrbf_b200.h -- C ABI of the B200-native (sm_100a) HRBFFusion3D hot path.
 *
 * Drop-in boundary for the reference's per-frame path:
 *   free functions of Core/src/Cuda/cudafuncs.cuh           -> hrbf_* map / step functions
 *   class RGBDOdometry  (Core/src/Utils/RGBDOdometry.h)      -> hrbf_odometry_*
 *   class IndexMap      (Core/src/IndexMap.h)                -> hrbf_indexmap_*
 *   class GlobalModel   (Core/src/GlobalModel.h)             -> hrbf_model_*
 *   HRBFFusion::processFrame / predict (Core/src/HRBFFusion.cpp:991-1260) -> hrbf_fusion_*
 *
 * Conventions
 *   - plain pointers and sizes only; `dev` pointers are CUDA device pointers,
 *     `host` pointers are host memory.  No torch / Eigen / GL types.
 *   - every function returns an int status (HRBF_OK == 0); nothing calls exit()
 *     (the reference's cudaSafeCall does: Cuda/convenience.cuh:64-71).
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *   - no allocation inside step functions: objects allocate once at create().
 *   - layouts are the reference's:
 *       SoA map     float[4*rows][cols], planes x,y,z,w stacked row-wise, `step` = row
 *                   pitch in BYTES (DeviceArray2D::step(), RGBDOdometry.cpp:128-136)
 *       AoS texture float[rows][cols][4] (RGBA32F as cudaMemcpyFromArray delivers it)
 *       surfel      5 x float4 = 80 B (Shaders/Vertex.cpp:20-44)
 *       A 6x6 row-major float, b 6 float, residual[2] = {sum w r^2, inliers}
 *       sums29      JtJJtrSE3 field order (Cuda/types.cuh:100-151)
 */
#ifndef HRBF_B200_H_
#define HRBF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HRBF_OK                0
#define HRBF_ERR_INVALID_ARG  -1
#define HRBF_ERR_CUDA         -2
#define HRBF_ERR_NO_DEVICE    -3
#define HRBF_ERR_CAPACITY     -4

#define HRBF_NUM_PYRS 3
#define HRBF_SURFEL_FLOATS 20

/* last error text of the calling thread ("" if none) */
const char* hrbf_last_error(void);
/* library / build identification, e.g. "hrbf_b200 0.1 sm_100a" */
const char* hrbf_version(void);
/* number of kernel launches issued by this library since load (bench.py's gpu_launches) */
unsigned long long hrbf_launch_count(void);

/* stream-ordered device-to-device copy (host mirrors use it to snapshot internal maps) */
int hrbf_copy_device(void* dst_dev, const void* src_dev, size_t bytes, void* stream);

typedef struct { float fx, fy, cx, cy; } hrbf_camera;   /* CameraModel, Cuda/types.cuh:82-98 */

/* ------------------------------------------------------------------------
 * Row 5 : pyramid / map preparation  (replaces Cuda/cudafuncs.cuh:148-214)
 * ---------------------------------------------------------------------- */
/* copyMaps, cudafuncs.cu:344-403 */
int hrbf_copy_maps(const float* v_aos_dev, const float* n_aos_dev,
                   float* vmap_dev, size_t vstep, float* nmap_dev, size_t nstep,
                   int rows, int cols, void* stream);
/* copyCurvatureMap, cudafuncs.cu:405-449 */
int hrbf_copy_curvature_map(const float* c_aos_dev, float* cmap_dev, size_t cstep,
                            int rows, int cols, float curvatureThreshold, void* stream);
/* copyicpWeightMap, cudafuncs.cu:452-491 */
int hrbf_copy_icpweight_map(const float* w_src_dev, float* w_dst_dev, size_t wstep,
                            int rows, int cols, void* stream);
/* resizeVMap / resizeNMap, cudafuncs.cu:526-615 (in_rows/in_cols = source size) */
int hrbf_resize_vmap(const float* in_dev, size_t in_step, float* out_dev, size_t out_step,
                     int in_rows, int in_cols, void* stream);
int hrbf_resize_nmap(const float* in_dev, size_t in_step, float* out_dev, size_t out_step,
                     int in_rows, int in_cols, void* stream);
/* resizeCMap, cudafuncs.cu:618-692 */
int hrbf_resize_cmap(const float* in_dev, size_t in_step, float* out_dev, size_t out_step,
                     int in_rows, int in_cols, void* stream);
/* resizeicpWeightMap, cudafuncs.cu:694-743 */
int hrbf_resize_icpweight_map(const float* in_dev, size_t in_step, float* out_dev, size_t out_step,
                              int in_rows, int in_cols, void* stream);
/* tranformMaps, cudafuncs.cu:213-277 (in place allowed); R row-major 3x3, t 3 (host) */
int hrbf_transform_maps(const float* vsrc_dev, size_t vsstep, const float* nsrc_dev, size_t nsstep,
                        const float* R_host, const float* t_host,
                        float* vdst_dev, size_t vdstep, float* ndst_dev, size_t ndstep,
                        int rows, int cols, void* stream);
/* transformCurvMaps, cudafuncs.cu:279-342 */
int hrbf_transform_curv_maps(const float* k1src_dev, size_t k1sstep, const float* k2src_dev, size_t k2sstep,
                             const float* R_host, const float* t_host,
                             float* k1dst_dev, size_t k1dstep, float* k2dst_dev, size_t k2dstep,
                             int rows, int cols, void* stream);

/* The RGB-branch / GPUTest-branch preparation functions of the same seam, one to one, on pitched arrays (step in bytes):
 *   pyrDown               Cuda/cudafuncs.cuh:177 (cudafuncs.cu:57-107)     depth pyramid level with a 3-sigma colour gate
 *   createVMap / NMap     cudafuncs.cuh:140-145 (cudafuncs.cu:109-211)     depth -> SoA vertex map, forward-difference normals
 *   verticesToDepth       cudafuncs.cuh:218 (cudafuncs.cu:874-894)         z of an RGBA32F vertex texture, NaN beyond maxDepth
 *   pyrDownGaussF         cudafuncs.cuh:181 (cudafuncs.cu:493-524,794-816) 5x5 Gaussian pyramid level of a float image
 *   pyrDownUcharGauss     cudafuncs.cuh:183 (cudafuncs.cu:818-871)         the same on u8 intensities
 *   imageBGRToIntensity   cudafuncs.cuh:213 (cudafuncs.cu:896-928)         RGBA8 texture -> u8 intensity (0.114 / 0.299 / 0.587)
 *   computeDerivativeImages cudafuncs.cuh:185 (cudafuncs.cu:930-993)       3x3 Sobel pair -> short dIdx, dIdy
 *   projectToPointCloud   cudafuncs.cuh:216 (cudafuncs.cu:995-1028)        depth -> float3 cloud with K of pyramid `level` */
int hrbf_pyr_down(const float* src_dev, size_t src_step, float* dst_dev, size_t dst_step, int src_rows, int src_cols, void* stream);
int hrbf_create_vmap(hrbf_camera intr, const float* depth_dev, size_t depth_step /* dense: cols * 4 */, float* vmap_dev, size_t vmap_step,
                     int rows, int cols, float depthCutoff, float depthMapFactor, void* stream);
int hrbf_create_nmap(const float* vmap_dev, size_t vmap_step, float* nmap_dev, size_t nmap_step, int rows, int cols, void* stream);
int hrbf_vertices_to_depth(const float* vert_aos_dev, float* depth_dev, size_t depth_step, int rows, int cols, float maxDepth, void* stream);
int hrbf_pyr_down_gauss_f(const float* src_dev, size_t src_step, float* dst_dev, size_t dst_step, int src_rows, int src_cols, void* stream);
int hrbf_pyr_down_uchar_gauss(const unsigned char* src_dev, size_t src_step, unsigned char* dst_dev, size_t dst_step, int src_rows, int src_cols, void* stream);
int hrbf_image_bgr_to_intensity(const unsigned char* rgba8_dev, unsigned char* dst_dev, size_t dst_step, int rows, int cols, void* stream);
int hrbf_compute_derivative_images(const unsigned char* src_dev, size_t src_step, short* dIdx_dev, size_t dx_step, short* dIdy_dev, size_t dy_step,
                                   int rows, int cols, void* stream);
int hrbf_project_to_point_cloud(const float* depth_dev, size_t depth_step, float* cloud3_dev, size_t cloud_step, hrbf_camera intr, int level,
                                int rows, int cols, void* stream);

/* ------------------------------------------------------------------------
 * Rows 1-3 : Jacobian-product reductions  (replaces Cuda/cudafuncs.cuh:82-147)
 * Synchronous like the reference: results are on the host when the call returns.
 * `work_dev` : >= hrbf_reduce_workspace_bytes() of device scratch.
 * ---------------------------------------------------------------------- */
size_t hrbf_reduce_workspace_bytes(void);

typedef struct {
    int use_search;          /* registrationICPUseCoorespondenceSearch */
    int search_radius;       /* registrationICPNeighborSearchRadius    */
    int use_weight;          /* icp_if_use_weight                      */
    float dist_thres;        /* RGBDOdometry.h:65 (0.1 m)              */
    float angle_thres;       /* RGBDOdometry.h:66 (sin 20 deg)         */
} hrbf_icp_options;

/* icpStep, reduce.cu:580-693.  corres_dev (optional) int2[rows*cols] */
int hrbf_icp_step(const float* Rcurr_host, const float* tcurr_host,
                  const float* vmap_curr_dev, const float* nmap_curr_dev,
                  const float* ck1_curr_dev, const float* ck2_curr_dev, size_t curr_step,
                  const float* Rprev_inv_host, const float* tprev_host, hrbf_camera intr,
                  const float* vmap_g_prev_dev, const float* nmap_g_prev_dev,
                  const float* ck1_g_prev_dev, const float* ck2_g_prev_dev, size_t prev_step,
                  const float* icpw_g_prev_dev, size_t icpw_step,
                  int rows, int cols, const hrbf_icp_options* opts,
                  int* corres_dev, void* work_dev,
                  float* A_host, float* b_host, float* residual_host, double* sums29_host,
                  void* stream);

/* DataTerm, Cuda/types.cuh:74-80 (16 B) */
typedef struct { short zero_x, zero_y, one_x, one_y; float diff; unsigned char valid; unsigned char pad[3]; } hrbf_dataterm;

/* computeRgbResidual, reduce.cu:1088-1154 (images dense, pitch == cols) */
int hrbf_compute_rgb_residual(float minScale, const short* dIdx_dev, const short* dIdy_dev,
                              const float* lastDepth_dev, const float* nextDepth_dev,
                              const unsigned char* lastImage_dev, const unsigned char* nextImage_dev,
                              hrbf_dataterm* corresImg_dev, float maxDepthDelta,
                              const float* kt_host, const float* krkinv_host,
                              int rows, int cols, void* work_dev,
                              int* sigmaSum_host, int* count_host, void* stream);
/* rgbStep, reduce.cu:842-896 */
int hrbf_rgb_step(const hrbf_dataterm* corresImg_dev, float sigma, const float* cloud3_dev,
                  float fx, float fy, const short* dIdx_dev, const short* dIdy_dev,
                  int use_gradient_weight, float sobelScale, int rows, int cols, void* work_dev,
                  float* A_host, float* b_host, double* sums29_host, void* stream);
/* so3Step, reduce.cu:1301-1359 */
int hrbf_so3_step(const unsigned char* lastImage_dev, const unsigned char* nextImage_dev,
                  const float* imageBasis_host, const float* kinv_host, const float* krlr_host,
                  int rows, int cols, void* work_dev,
                  float* A_host, float* b_host, float* residual_host, double* sums11_host, void* stream);

/* ------------------------------------------------------------------------
 * Row 4 : RGBDOdometry  (Utils/RGBDOdometry.h:57-107)
 * "Textures" are dense device buffers: RGBA32F AoS float[h][w][4], R32F float[h][w],
 * RGBA8 uchar[h][w][4].
 * ---------------------------------------------------------------------- */
typedef struct hrbf_odometry hrbf_odometry;

/* RGBDOdometry::RGBDOdometry, RGBDOdometry.cpp:35-154 */
int hrbf_odometry_create(hrbf_odometry** out, int width, int height,
                         float cx, float cy, float fx, float fy,
                         float distThresh, float angleThresh);
int hrbf_odometry_destroy(hrbf_odometry*);
/* knobs read from GlobalStateParam inside the reference path */
int hrbf_odometry_set_params(hrbf_odometry*, float curvValidThreshold, int useCorrespondenceSearch,
                             int searchRadius, int rgbUseGradientWeight);
/* Tracker implementation: 0 (default) = the whole coarse-to-fine loop as ONE persistent cooperative kernel;
 * 1 = one kernel per reduction, replayed as a CUDA graph (kept for comparison and for the per-kernel timing hook) */
int hrbf_odometry_set_tracker(hrbf_odometry*, int use_kernel_graph);
/* Threads per CTA of the persistent tracker: 512 (default; 512 x 128 registers fill an SM's register file -- lowest latency
 * for ONE sequence) or 256 (leaves half of every SM to the kernels of OTHER fusion objects running on their own streams:
 * several sequences per GPU, offline throughput runs).  Replaces the launch-shape table of Utils/GPUConfig.h:53-78. */
int hrbf_odometry_set_tracker_threads(hrbf_odometry*, int threads);
/* persistent tracker: resident != 0 keeps, per pyramid level, every CTA's tile of packed ICP records and the model window its
 * associations fall into in shared memory for all Gauss-Newton iterations of the level (staged by 2-D tensor-map TMA loads); 0 (default)
 * = every iteration re-reads the maps through L2.  Measured on a B200 (profiles/r2_*): the pass is bound by instruction issue, not by
 * the L2 round trips the streaming form already overlaps, and the resident form's extra index arithmetic makes it ~20 % slower. */
int hrbf_odometry_set_tracker_tiles(hrbf_odometry*, int resident);
/* initICP(depth) [GPUTest path], RGBDOdometry.cpp:161-181 : depth_dev = float[h][w] raw units */
int hrbf_odometry_init_icp_depth(hrbf_odometry*, const float* depth_dev, float depthCutoff, float depthMapFactor, void* stream);
/* initICP(vertices, normals), RGBDOdometry.cpp:183-206 */
int hrbf_odometry_init_icp(hrbf_odometry*, const float* vert_aos_dev, const float* norm_aos_dev, float depthCutoff, void* stream);
/* initICPModel, RGBDOdometry.cpp:208-247 : modelPose row-major 4x4 (host) */
int hrbf_odometry_init_icp_model(hrbf_odometry*, const float* vert_aos_dev, const float* norm_aos_dev,
                                 float depthCutoff, const float* modelPose_host, void* stream);
/* initRGB / initRGBModel / initFirstRGB, RGBDOdometry.cpp:689-699, 777-794 */
int hrbf_odometry_init_rgb(hrbf_odometry*, const unsigned char* rgba_dev, void* stream);
int hrbf_odometry_init_rgb_model(hrbf_odometry*, const unsigned char* rgba_dev, void* stream);
int hrbf_odometry_init_first_rgb(hrbf_odometry*, const unsigned char* rgba_dev, void* stream);
/* initCurvature / initCurvatureModel, RGBDOdometry.cpp:701-759 */
int hrbf_odometry_init_curvature(hrbf_odometry*, const float* k1_aos_dev, const float* k2_aos_dev, void* stream);
int hrbf_odometry_init_curvature_model(hrbf_odometry*, const float* k1_aos_dev, const float* k2_aos_dev,
                                       const float* modelPose_host, void* stream);
/* initICPweight, RGBDOdometry.cpp:761-775 */
int hrbf_odometry_init_icp_weight(hrbf_odometry*, const float* w_dev, void* stream);
/* GPUTest-pair convention: curvature planes 0 (valid), weights 1 */
int hrbf_odometry_fill_neutral_curvature(hrbf_odometry*, void* stream);

typedef struct {
    float lastICPError, lastICPCount, lastRGBError, lastRGBCount, lastSO3Error, lastSO3Count;
    double lastA[36], lastb[6];
    int icp_iterations_run;
    int kernel_launches;
} hrbf_track_stats;

/* getIncrementalTransformation, RGBDOdometry.cpp:796-1249.
 * trans_host[3], rot_host[9] (row-major): in = previous pose, out = estimated pose.
 * The whole coarse-to-fine Gauss-Newton loop (reductions, 6x6 LDLT in fp64, SE3 update)
 * runs on the device without host round trips; the call returns after one sync. */
int hrbf_odometry_get_incremental_transformation(hrbf_odometry*, float* trans_host, float* rot_host,
                                                 int rgbOnly, float icpWeight, int pyramid, int fastOdom,
                                                 int so3, int if_curvature_info, int index_frame,
                                                 hrbf_track_stats* stats_host, void* stream);
/* asynchronous variant: the pose stays on the device (pose_dev: float[12] = rot[9], trans[3]) */
int hrbf_odometry_track_async(hrbf_odometry*, const float* prev_pose_dev, float* pose_out_dev,
                              int rgbOnly, float icpWeight, int pyramid, int fastOdom,
                              int so3, int if_curvature_info, void* stream);
/* Measurement hook (bench.py roofline): average duration in microseconds of `reps` back-to-back launches of
 * one reduction kernel of the tracking loop at pyramid `level`, timed with CUDA events on `stream`.
 * which: 0 ICP JTJ/JTr reduction, 1 RGB residual, 2 RGB step, 3 SO3 step.  with_update != 0 keeps the
 * in-kernel Gauss-Newton solve.  which = 4: ONE launch of the persistent tracker running `reps` ICP-only Gauss-Newton
 * iterations at `level` (reduction + cross-CTA exchange + fp64 solve); avg_us is the time per iteration.
 * Needs initialised pyramids and one prior tracking call. */
int hrbf_odometry_time_kernel(hrbf_odometry*, int which, int level, int with_update, int reps, float* avg_us, void* stream);
/* icpStep (reduce.cu:580-693) on the object's OWN pyramid level (the maps the init* calls built) for a caller-given pose:
 * tiled != 0 runs the TMA-staged tile form of the reduction (the one the tracking loop uses when correspondence search is off),
 * 0 the per-pixel-gather form.  Outputs as hrbf_icp_step.  which = 5 / 6 / 7 of hrbf_odometry_time_kernel time these two:
 * 5 = tile form, back-to-back launches (L2-hot); 6 = tile form, 7 = gather form, each launch timed alone after an L2 flush (cold). */
int hrbf_odometry_icp_step(hrbf_odometry*, int level, const float* Rcurr_host, const float* tcurr_host, const float* Rprev_inv_host,
                           const float* tprev_host, int use_weight, int tiled, float* A_host, float* b_host, float* residual_host,
                           double* sums29_host, void* stream);
/* device views of the internal pyramid maps (tests, chaining):
 * which = 0..8 -> vmap_g_prev,nmap_g_prev,ck1_g_prev,ck2_g_prev,vmap_curr,nmap_curr,ck1_curr,ck2_curr,icpWeight */
const float* hrbf_odometry_map(const hrbf_odometry*, int which, int level, size_t* step_bytes);
const unsigned char* hrbf_odometry_image(const hrbf_odometry*, int which, int level); /* 0 last, 1 next, 2 lastNext, 3 (level 2, frame pipeline) the staged SO3 pre-alignment's image */
const float* hrbf_odometry_depth(const hrbf_odometry*, int which, int level);        /* 0 last, 1 next */
/* next-image Sobel derivatives (computeDerivativeImages, RGBDOdometry.cpp:951-956; short[rows][cols], axis 0 = dI/dx, 1 = dI/dy) and the
 * pose-independent candidate mask of computeRgbResidual (reduce.cu:1000-1023; bytes) of the bank the tracker last used */
const short* hrbf_odometry_gradient(const hrbf_odometry*, int axis, int level);
const unsigned char* hrbf_odometry_candidates(const hrbf_odometry*, int level);

/* ------------------------------------------------------------------------
 * Rows 6-7 : IndexMap  (Core/src/IndexMap.h:36-201)
 * Every GPUTexture of the reference is a dense device buffer: RGBA32F -> float[h][w][4],
 * R32UI -> uint32[h][w], RGBA8 -> uchar[h][w][4], R16UI -> uint16[h][w], R32F -> float[h][w].
 * ---------------------------------------------------------------------- */
typedef struct hrbf_indexmap hrbf_indexmap;
#define HRBF_ACTIVE_KEYFRAME_DIMENSION 19200      /* IndexMap::ACTIVE_KEYFRAME_DIMENSION, IndexMap.cpp:24 */

enum hrbf_indexmap_tex {                           /* accessor of IndexMap.h it replaces */
    HRBF_TEX_INDEX = 0,        /* indexTex()      u32  */
    HRBF_TEX_VERTCONF,         /* vertConfTex()   f32x4: camera-frame xyz, confidence */
    HRBF_TEX_COLORTIME,        /* colorTimeTex()  f32x4 */
    HRBF_TEX_NORMRAD,          /* normalRadTex()  f32x4 */
    HRBF_TEX_CURVMAX,          /* curvMaxTex()    f32x4 */
    HRBF_TEX_CURVMIN,          /* curvMinTex()    f32x4 */
    HRBF_TEX_IMAGE_HRBF,       /* imageTexHRBF()  u8x4  */
    HRBF_TEX_VERTEX_HRBF,      /* vertexTexHRBF() f32x4: xyz, confidence */
    HRBF_TEX_NORMAL_HRBF,      /* normalTexHRBF() f32x4: xyz, radius */
    HRBF_TEX_CURVK1_HRBF,      /* curvk1TexHRBF() f32x4 */
    HRBF_TEX_CURVK2_HRBF,      /* curvk2TexHRBF() f32x4 */
    HRBF_TEX_TIME_HRBF,        /* (timeTexture)   u16   */
    HRBF_TEX_ICPW_HRBF,        /* icpweightTexHRBF() f32 */
    HRBF_TEX_OLD_IMAGE_HRBF, HRBF_TEX_OLD_VERTEX_HRBF, HRBF_TEX_OLD_NORMAL_HRBF, HRBF_TEX_OLD_CURVK1_HRBF,
    HRBF_TEX_OLD_CURVK2_HRBF, HRBF_TEX_OLD_TIME_HRBF, HRBF_TEX_OLD_ICPW_HRBF,   /* old*TexHRBF(): INACTIVE prediction */
    HRBF_TEX_COUNT
};

/* IndexMap::IndexMap, IndexMap.cpp:25-191 (camera from the Intrinsics singleton there) */
int hrbf_indexmap_create(hrbf_indexmap** out, int width, int height, float cx, float cy, float fx, float fy);
int hrbf_indexmap_destroy(hrbf_indexmap*);
/* IndexMap::lActiveKFID (IndexMap.h:199, used at IndexMap.cpp:222-237): ids of the active sub-maps */
int hrbf_indexmap_set_active_keyframes(hrbf_indexmap*, const int* ids_host, int n, void* stream);
/* IndexMap::predictIndices, IndexMap.cpp:193-267.  pose16_host: row-major 4x4 camera-to-world;
 * surfels_dev: float[count][20] (the model VBO).  Writes the 6 index-map textures. */
int hrbf_indexmap_predict_indices(hrbf_indexmap*, const float* pose16_host, int time, int maxTime,
                                  const float* surfels_dev, unsigned int count, float depthCutoff,
                                  int insertSubmap, int indexSubmap, void* stream);
/* same, fully device-driven (no host pose / count): inv_pose_dev = R^T[9], -R^T t[3]; count read on the device */
int hrbf_indexmap_predict_indices_dev(hrbf_indexmap*, const float* inv_pose_dev, const float* surfels_dev,
                                      const unsigned int* count_dev, unsigned int count_bound, float depthCutoff, void* stream);
/* IndexMap::predictHRBF, IndexMap.cpp:413-518 (predictionType 0 = ACTIVE, 1 = INACTIVE); the GlobalStateParam
 * knobs the reference reads there are arguments: preictionWindowMultiplier (<= 3), preictionMinNeighbors,
 * preictionMaxNeighbors (<= 16), preictionConfThreshold, registrationICPCurvWeightImpactControl */
int hrbf_indexmap_predict_hrbf(hrbf_indexmap*, int predictionType, int win, int minNeighbors, int maxNeighbors,
                               float confThreshold, float icpWeightLambda, void* stream);
/* device pointer of one texture (enum hrbf_indexmap_tex) */
void* hrbf_indexmap_texture(hrbf_indexmap*, int which);

/* ------------------------------------------------------------------------
 * Row 10 : per-frame preprocessing (HRBFFusion::filterDepth / metriciseDepth / computeVertexNormalRadius /
 * computeCurvatureGradient / updateNormalRad / VertexConfidence, Core/src/HRBFFusion.cpp:1016-1021, 1262-1346)
 * and FillIn (Core/src/Shaders/FillIn.{h,cpp}).  A hrbf_frame owns the reference's `textures[...]` map.
 * ---------------------------------------------------------------------- */
typedef struct hrbf_frame hrbf_frame;
enum hrbf_frame_tex {            /* GPUTexture:: name (GPUTexture.cpp:20-37) */
    HRBF_FT_RGB = 0,             /* RGB               u8 x3 (as uploaded)   */
    HRBF_FT_RGBA,                /* RGB as the GL_RGBA texture reads: u8 x4, alpha 255 */
    HRBF_FT_DEPTH_RAW,           /* DEPTH             u16                   */
    HRBF_FT_DEPTH_FILTERED,      /* DEPTH_FILTERED    f32 (raw units)       */
    HRBF_FT_DEPTH_METRIC,        /* DEPTH_METRIC      f32 metres            */
    HRBF_FT_DEPTH_METRIC_FILTERED,
    HRBF_FT_VERTEX_RAW,          /* f32x4 xyz, radial confidence            */
    HRBF_FT_VERTEX_FILTERED,     /* f32x4 xyz, 1                            */
    HRBF_FT_NORMAL_PCA,          /* NORMAL before updateNormalRad: PCA normal, radius */
    HRBF_FT_NORMAL,              /* NORMAL (= NORMAL_OPT after updateNormalRad): HRBF-gradient normal, radius */
    HRBF_FT_PRINCIPAL_CURV1,     /* f32x4 direction, k1                     */
    HRBF_FT_PRINCIPAL_CURV2,
    HRBF_FT_GRADIENT_MAG,        /* f32                                     */
    HRBF_FT_RADIUS,              /* f32                                     */
    HRBF_FT_CONFIDENCE,          /* f32                                     */
    HRBF_FT_COUNT
};
typedef struct {
    int width, height;
    float cx, cy, fx, fy;
    float depthFactor;           /* metres per raw unit = 1 / DepthMapFactor (HRBFFusion.cpp:771-781), TUM: 1/5000 */
    float depthCutoff;           /* globalDepthCutoff (3.5)                 */
    float radiusMultiplier;      /* preprocessingInitRadiusMultiplier (4)   */
    int normalPCA;               /* preprocessingNormalEstimationPCA (1)    */
    int curvWindow;              /* preprocessingCurvEstimationWindow (3)   */
    int bilateral;               /* preprocessingUsebilateralFilter (1)     */
    int useConfEval;             /* preprocessingUseConfEval (0)            */
    float confEvalEpsilon;       /* preprocessingConfEvalEpsilon (1000)     */
} hrbf_frame_params;
int hrbf_frame_create(hrbf_frame** out, const hrbf_frame_params* p);
int hrbf_frame_destroy(hrbf_frame*);
/* textures[DEPTH_RAW/RGB]->Upload (HRBFFusion.cpp:1006-1010); host != 0: pointers are host memory (pinned for async) */
int hrbf_frame_upload(hrbf_frame*, const unsigned char* rgb8, const unsigned short* depth16, int host, void* stream);
/* filterDepth + metriciseDepth + computeVertexNormalRadius + computeCurvatureGradient + updateNormalRad */
int hrbf_frame_preprocess(hrbf_frame*, void* stream);
/* VertexConfidence(weighting), HRBFFusion.cpp:1310-1327 */
int hrbf_frame_vertex_confidence(hrbf_frame*, float weighting, void* stream);
void* hrbf_frame_texture(hrbf_frame*, int which);

typedef struct hrbf_fillin hrbf_fillin;
enum hrbf_fillin_tex { HRBF_FILL_IMAGE = 0, HRBF_FILL_VERTEX, HRBF_FILL_NORMAL, HRBF_FILL_CURVK1, HRBF_FILL_CURVK2, HRBF_FILL_ICPWEIGHT, HRBF_FILL_COUNT };
int hrbf_fillin_create(hrbf_fillin** out, int width, int height);
int hrbf_fillin_destroy(hrbf_fillin*);
/* FillIn::vertex + normal + curvature + image (FillIn.cpp:78-380), one fused pass.  lambda =
 * registrationICPCurvWeightImpactControl, curvThr = preprocessingCurvValidThreshold */
int hrbf_fillin_run(hrbf_fillin*, hrbf_indexmap* prediction, hrbf_frame* frame, int passthrough, float lambda, float curvThr, void* stream);
void* hrbf_fillin_texture(hrbf_fillin*, int which);
/* HRBFFusion::denseEnough(resize.vertex(...)) (HRBFFusion.cpp:974-987,1069-1070): 1 if more than `thresh` of the
 * 1/20-subsampled vertex texture has z > 0.  Synchronous (reads one int back). */
int hrbf_dense_enough(const float* vertex_tex_dev, int width, int height, float thresh, int* dense_host, void* stream);

/* ------------------------------------------------------------------------
 * Rows 8-9 : GlobalModel  (Core/src/GlobalModel.h:35-150)
 * The model "VBO" is a device array of 80-B surfels; the count lives on the device.
 * ---------------------------------------------------------------------- */
typedef struct hrbf_model hrbf_model;
/* capacity in surfels (reference: TEXTURE_DIMENSION^2 = 4596^2, GlobalModel.cpp:21-22); 0 -> that default */
int hrbf_model_create(hrbf_model** out, int width, int height, float cx, float cy, float fx, float fy, unsigned int capacity);
int hrbf_model_destroy(hrbf_model*);
/* GlobalStateParam knobs read inside initialise / fuse / clean */
int hrbf_model_set_params(hrbf_model*, float radiusMultiplier, float curvValidThreshold, int normalPCA, int cleanWindow,
                          int useConfEval, float confEvalEpsilon);
int hrbf_model_set_active_keyframes(hrbf_model*, const int* ids_host, int n, void* stream);   /* GlobalModel::lActiveKFID */
/* GlobalModel::initialise, GlobalModel.cpp:214-288 (colorMap: RGB8) */
int hrbf_model_initialise(hrbf_model*, const float* vertexMap, const float* normalMap, const unsigned char* colorMap_rgb8,
                          const float* curv1Map, const float* curv2Map, const float* gradientMagMap,
                          const float* init_pose16_host, void* stream);
/* GlobalModel::fuse, GlobalModel.cpp:355-549 */
int hrbf_model_fuse(hrbf_model*, const float* pose16_host, int time, const unsigned char* rgb8,
                    const float* depthRaw_metric, const float* depthFiltered_metric, const float* curv1, const float* curv2,
                    const float* confidence, const unsigned int* indexMap, const float* vertConfMap, const float* colorTimeMap,
                    const float* normRadMap, float depthCutoff, float confThreshold, float weighting,
                    int insertSubmap, int indexSubmap, void* stream);
/* GlobalModel::clean, GlobalModel.cpp:551-688 */
int hrbf_model_clean(hrbf_model*, const float* pose16_host, int time, const unsigned int* indexMap, const float* vertConfMap,
                     const float* colorTimeMap, const float* normRadMap, const float* depthMap, float confThreshold,
                     float maxDepth, void* stream);
/* GlobalModel::updateModel, GlobalModel.cpp:690-767 (+ Shaders/update_delta_trans.vert): rigid correction of every surfel by
 * the matrix of its sub-map, DeltaTransformKF[(int)colour.y]; delta16_host = n_delta row-major 4x4 matrices.  In place. */
int hrbf_model_update_model(hrbf_model*, const float* delta16_host, int n_delta, void* stream);
/* load a surfel array (e.g. a saved map) as the current model; host != 0: `surfels` is host memory.  Synchronous. */
int hrbf_model_set_model(hrbf_model*, const float* surfels, unsigned int count, int host, void* stream);
/* GlobalModel::model() / lastCount(): device pointer of the current surfel array; lastCount synchronises */
const float* hrbf_model_model(hrbf_model*);
const unsigned int* hrbf_model_count_dev(hrbf_model*);
int hrbf_model_last_count(hrbf_model*, unsigned int* count_host, void* stream);
/* 1 if a compaction ever ran out of capacity (survivors beyond it were dropped) */
int hrbf_model_overflowed(hrbf_model*, int* flag_host, void* stream);
/* The samples a shader's float-counter window loop visits along one axis of n texels (geometry.glsl:198-212,
 * depth_curvature_gradient.frag:62-75; win = 3): per pixel the first texel, the sample count (6 or 7 in the interior) and the
 * coordinates i * n.  uv_vbo_coords = 1: the texcoords of GlobalModel::fuse's uv VBO (GlobalModel.cpp:87-96).  Host code. */
#define HRBF_WINDOW_MAX 8
int hrbf_window_table(int n, float win, int uv_vbo_coords, int* first_host, int* count_host, float* coords_host /* [n][HRBF_WINDOW_MAX] */);
/* GlobalModel::downloadMap, GlobalModel.cpp:775-804: the surfel array (count x 20 floats) to surfels_host.
 * max_count = capacity of surfels_host in surfels; 0 = size query (only *count_out is written). */
int hrbf_model_download_map(hrbf_model*, float* surfels_host, unsigned int max_count, unsigned int* count_out, void* stream);
/* HRBFFusion::savePly, HRBFFusion.cpp:1737-1853.  hrbf_ply_header writes the reference's header text for n vertices
 * (returns its length, 0 if buf is too small).  hrbf_model_export_ply filters the map (confidence >
 * globalOutputSavePointCloudConfThreshold) and packs the 43-byte binary_little_endian vertices {x y z, r g b, -n, curvature_max,
 * curvature_min, radius, submapIndex} ON THE DEVICE, in map order, and copies them to records_host; records_host = NULL is a
 * size query.  File = header + records. */
#include <stddef.h>
size_t hrbf_ply_header(unsigned int n_vertices, char* buf, size_t buf_len);
int hrbf_model_export_ply(hrbf_model*, float confThreshold, void* records_host, size_t host_bytes, unsigned int* n_vertices_out, void* stream);

/* ------------------------------------------------------------------------
 * The per-frame orchestrator: HRBFFusion::processFrame / predict (Core/src/HRBFFusion.cpp:991-1260)
 * with the sparse back-end off (optimizationUseLocalBA = optimizationUseGlobalBA = false).
 * Everything between the input upload and the pose read-back is enqueued on one stream without a host
 * round trip: pose, fusion weight, fill-in decision and surfel count all stay on the device.
 * ---------------------------------------------------------------------- */
typedef struct hrbf_fusion hrbf_fusion;
typedef struct {
    hrbf_frame_params frame;
    float confidenceThreshold;    /* globalConfidenceThreshold (5)           */
    float maxDepthProcessed;      /* HRBFFusion.cpp:37 (20)                  */
    float icpWeight;              /* registrationJointICPWeight (10)         */
    int rgbOnly, pyramid, fastOdom, so3, weightedICP;   /* false, true, false, registrationPreAlignSO3, registrationICPUseWeightedICP */
    int predWindow, predMinNeighbors, predMaxNeighbors; /* preictionWindowMultiplier 3, preictionMinNeighbors 6, preictionMaxNeighbors 10 */
    float predConfThreshold;      /* preictionConfThreshold (3)              */
    float icpWeightLambda;        /* registrationICPCurvWeightImpactControl (10) */
    float curvValidThreshold;     /* preprocessingCurvValidThreshold (300)   */
    float denseEnoughThresh;      /* globalDenseEnoughThresh (0.75)          */
    int cleanWindow;              /* fusionCleanWindowMultiplier (2)         */
    unsigned int capacity;        /* surfel capacity, 0 = reference default  */
    int trackerThreads;           /* hrbf_odometry_set_tracker_threads: 0 = 384 = one sequence, replayed through stage_frame or not (the staged
                                     preprocessing of frame t+1 co-resides with the tracker of frame t), 256 = several sequences per GPU */
} hrbf_fusion_params;
void hrbf_fusion_default_params(hrbf_fusion_params* p, int width, int height, float cx, float cy, float fx, float fy);
int hrbf_fusion_create(hrbf_fusion** out, const hrbf_fusion_params* p);
int hrbf_fusion_destroy(hrbf_fusion*);
/* processFrame with HOST input buffers (rgb8: h*w*3, depth16: h*w); blocks until the pose is on the host.
 * pose16_out_host (optional): row-major 4x4 currPose */
int hrbf_fusion_process_frame(hrbf_fusion*, const unsigned char* rgb8_host, const unsigned short* depth16_host,
                              long long timestamp, float weightMultiplier, float* pose16_out_host, void* stream);
/* same, inputs already on the device, nothing is read back: enqueue-only.  Use hrbf_fusion_get_pose to synchronise. */
int hrbf_fusion_process_frame_dev(hrbf_fusion*, const unsigned char* rgb8_dev, const unsigned short* depth16_dev,
                                  long long timestamp, float weightMultiplier, void* stream);
int hrbf_fusion_get_pose(hrbf_fusion*, float* pose16_out_host, void* stream);
/* Pipelined processFrame for log replay, where frame t+1 is available while frame t is still being processed (the reference's
 * MainController loop reads and processes strictly in turn; this is an addition, not a mirrored call).
 * stage_frame: upload (host != 0: pinned host memory) + everything that depends on the camera frame alone -- preprocessing, the
 * current-frame pyramids of initICP / initRGB / initCurvature, Sobel images and candidate masks, the SO3 pre-alignment against the
 * previous camera frame -- on internal streams of the LOWEST priority, concurrently with whatever `stream` is still doing for the
 * previous frame; at most two frames may be staged.  Give `stream` a higher priority (cudaStreamCreateWithPriority) so that the
 * staged work only takes what the frame being processed leaves free (measured: +2.5 % frames/s over equal priorities).
 * process_staged: processFrame of the oldest staged frame on `stream`; pose16_out_host == NULL -> enqueue only.
 * Results are identical to hrbf_fusion_process_frame on the same frames. */
int hrbf_fusion_stage_frame(hrbf_fusion*, const unsigned char* rgb8, const unsigned short* depth16, int host, void* stream);
int hrbf_fusion_process_staged(hrbf_fusion*, long long timestamp, float weightMultiplier, float* pose16_out_host, void* stream);
int hrbf_fusion_tick(const hrbf_fusion*);
/* per-frame poses since creation, device array float[frames][12] (R row-major, t): gathered by the multi-GPU bench */
const float* hrbf_fusion_trajectory_dev(hrbf_fusion*, int* n_frames);
/* the components (reference members frameToModel, indexMap, globalModel, textures, fillIn) */
hrbf_odometry* hrbf_fusion_odometry(hrbf_fusion*);
hrbf_indexmap* hrbf_fusion_indexmap(hrbf_fusion*);
hrbf_model* hrbf_fusion_model(hrbf_fusion*);
hrbf_frame* hrbf_fusion_frame(hrbf_fusion*);
hrbf_fillin* hrbf_fusion_fillin(hrbf_fusion*);
/* CUDA-event timings of the last frame in ms, the reference's Stopwatch spans (HRBFFusion.cpp:1016,1063,1196,1248):
 * [0] Initialization [1] Registration [2] Integration [3] Prediction.  Only recorded when enabled; the spans of the last enqueued frame
 * once it has finished (synchronise first), else of the frame before.  With staged frames the Initialization span is empty: that work ran
 * ahead on the staging streams. */
int hrbf_fusion_enable_timings(hrbf_fusion*, int on);
int hrbf_fusion_last_timings(hrbf_fusion*, float ms4_host[4]);

#ifdef __cplusplus
}
#endif
#endif /* HRBF_B200_H_ */
