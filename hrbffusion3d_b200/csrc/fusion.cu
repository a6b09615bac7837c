// fusion.cu -- host side of SURVEY.md section 8 rows 8-10 and the per-frame orchestrator:
//   hrbf_frame   : the reference's textures[] + preprocessing ComputePacks (Core/src/HRBFFusion.cpp:785-933, 1262-1346)
//   hrbf_fillin  : Core/src/Shaders/FillIn.{h,cpp}
//   hrbf_model   : Core/src/GlobalModel.{h,cpp} (initialise / fuse / clean)
//   hrbf_fusion  : HRBFFusion::processFrame / predict (HRBFFusion.cpp:991-1260), sparse back-end off
// One stream, no host round trip between the input upload and the pose read-back: the pose, the fusion
// weight, the fill-in decision and the surfel count all stay in device memory.
#include "model_kernels.cuh"
#include "hrbf_internal.h"
#include <new>
#include <vector>

using namespace hrbf;

// ======================================================================== hrbf_frame
struct hrbf_frame {
    hrbf_frame_params p{};
    PrepArgs a{};
    char* wintab = nullptr;          // device tables of the literal window loops (PrepArgs::wx / wy point into it)
    char* slab = nullptr;
    void* tex[HRBF_FT_COUNT] = {};
    float* weighting = nullptr;      // device scalar (VertexConfidence's uniform)
    float* h_w = nullptr;            // pinned ring of 8
    int slot = 0;
    float4* fuse_normals = nullptr;  // frame pipeline: fuse_normals_kernel's output for frame number fuse_normals_time (per fuse slot)
    int fuse_normals_time = -1;
};

static size_t ft_bytes(int which, size_t P)
{
    switch (which) {
    case HRBF_FT_RGB: return P * 3;
    case HRBF_FT_RGBA: return P * 4;
    case HRBF_FT_DEPTH_RAW: return P * 2;
    case HRBF_FT_DEPTH_FILTERED: case HRBF_FT_DEPTH_METRIC: case HRBF_FT_DEPTH_METRIC_FILTERED:
    case HRBF_FT_GRADIENT_MAG: case HRBF_FT_RADIUS: case HRBF_FT_CONFIDENCE: return P * 4;
    default: return P * 16;
    }
}
static PrepArgs prep_args(const hrbf_frame_params& p)
{
    PrepArgs a;
    a.cols = p.width; a.rows = p.height; a.cx = p.cx; a.cy = p.cy;
    a.icx = (float)(1.0 / (double)p.fx); a.icy = (float)(1.0 / (double)p.fy);
    a.depthFactor = p.depthFactor; a.maxD = p.depthCutoff; a.radiusMultiplier = p.radiusMultiplier;
    a.pca = p.normalPCA; a.curvWin = p.curvWindow; a.bilateral = p.bilateral;
    return a;
}

// Builds the four window tables {x, y} x {fragment-pass texcoords, uv-VBO texcoords} with hrbf_window_table, uploads them as one
// device block and points a.wx / a.wy at it.  Table of an axis of n pixels: first[n] (int), count[n] (int), coord[n][kWinMax] (float).
static int build_window_tables(PrepArgs& a, float win, char** dev_out)
{
    const int dims[2] = { a.cols, a.rows };
    size_t off[2][2], total = 0;
    for (int v = 0; v < 2; ++v)
        for (int ax = 0; ax < 2; ++ax) { off[v][ax] = total; total += (size_t)dims[ax] * (2 * sizeof(int) + kWinMax * sizeof(float)); }
    std::vector<char> host(total);
    for (int v = 0; v < 2; ++v)
        for (int ax = 0; ax < 2; ++ax) {
            const int n = dims[ax];
            char* b = host.data() + off[v][ax];
            if (int rc = hrbf_window_table(n, win, v, (int*)b, (int*)(b + (size_t)n * sizeof(int)), (float*)(b + (size_t)n * 2 * sizeof(int)))) return rc;
        }
    char* dev = nullptr;
    HRBF_CUDA(cudaMalloc(&dev, total));
    HRBF_CUDA(cudaMemcpy(dev, host.data(), total, cudaMemcpyHostToDevice));
    for (int v = 0; v < 2; ++v)
        for (int ax = 0; ax < 2; ++ax) {
            const int n = dims[ax];
            char* b = dev + off[v][ax];
            const WinTab t = { (const int*)b, (const int*)(b + (size_t)n * sizeof(int)), (const float*)(b + (size_t)n * 2 * sizeof(int)) };
            (ax == 0 ? a.wx : a.wy)[v] = t;
        }
    *dev_out = dev;
    return HRBF_OK;
}

extern "C" {

int hrbf_frame_create(hrbf_frame** out, const hrbf_frame_params* p)
{
    HRBF_CHECK_ARG(out && p && p->width > 0 && p->height > 0 && p->curvWindow >= 0 && p->curvWindow <= 3 && p->depthFactor > 0);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device"); return HRBF_ERR_NO_DEVICE; }
    hrbf_frame* f = new (std::nothrow) hrbf_frame();
    HRBF_CHECK_ARG(f != nullptr);
    f->p = *p; f->a = prep_args(*p);
    if (p->curvWindow != 3) { set_error("literal-window build: preprocessingCurvEstimationWindow must be 3"); delete f; return HRBF_ERR_INVALID_ARG; }
    if (int rc = build_window_tables(f->a, 3.0f, &f->wintab)) { delete f; return rc; }
    const size_t P = (size_t)p->width * p->height;
    size_t off = 0, o[HRBF_FT_COUNT];
    auto take = [&](size_t bytes) { size_t r = off; off += (bytes + 255) & ~(size_t)255; return r; };
    for (int t = 0; t < HRBF_FT_COUNT; ++t) o[t] = take(ft_bytes(t, P));
    const size_t o_w = take(64), o_fn = take((size_t)fuse_slots_x(p->width) * fuse_slots_y(p->height) * sizeof(float4));
    if (cudaMalloc(&f->slab, off) != cudaSuccess) { set_error("cudaMalloc(%zu) failed", off); delete f; return HRBF_ERR_CUDA; }
    cudaMemset(f->slab, 0, off);
    for (int t = 0; t < HRBF_FT_COUNT; ++t) f->tex[t] = f->slab + o[t];
    f->weighting = (float*)(f->slab + o_w);
    f->fuse_normals = (float4*)(f->slab + o_fn);
    cudaMallocHost(&f->h_w, 8 * sizeof(float));
    {
        float tab[(2 * kBilR + 1) * (2 * kBilR + 1)];
        make_bilateral_table(tab);
        if (cudaMemcpyToSymbol(c_bil_space, tab, sizeof tab) != cudaSuccess) { set_error("frame_create: constant upload failed"); hrbf_frame_destroy(f); return HRBF_ERR_CUDA; }
    }
    *out = f;
    return HRBF_OK;
}
int hrbf_frame_destroy(hrbf_frame* f)
{
    if (!f) return HRBF_OK;
    if (f->h_w) cudaFreeHost(f->h_w);
    if (f->slab) cudaFree(f->slab);
    if (f->wintab) cudaFree(f->wintab);
    delete f;
    return HRBF_OK;
}
}  // extern "C"
// make_rgba = false: the RGBA texture is not refreshed (the frame pipeline reads the RGB8 upload directly after the first frame)
static int frame_upload(hrbf_frame* f, const unsigned char* rgb8, const unsigned short* depth16, int host, bool make_rgba, cudaStream_t s)
{
    const size_t P = (size_t)f->p.width * f->p.height;
    const cudaMemcpyKind k = host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    HRBF_CUDA(cudaMemcpyAsync(f->tex[HRBF_FT_RGB], rgb8, P * 3, k, s));
    HRBF_CUDA(cudaMemcpyAsync(f->tex[HRBF_FT_DEPTH_RAW], depth16, P * 2, k, s));
    if (make_rgba) {
        rgb_to_rgba_kernel<<<div_up((int)P, 256), 256, 0, s>>>((int)P, (const unsigned char*)f->tex[HRBF_FT_RGB], (uchar4*)f->tex[HRBF_FT_RGBA]);
        HRBF_KERNEL_CHECK();
    }
    return HRBF_OK;
}
extern "C" {
int hrbf_frame_upload(hrbf_frame* f, const unsigned char* rgb8, const unsigned short* depth16, int host, void* stream)
{
    HRBF_CHECK_ARG(f && rgb8 && depth16);
    return frame_upload(f, rgb8, depth16, host, true, (cudaStream_t)stream);
}
int hrbf_frame_preprocess(hrbf_frame* f, void* stream)
{
    HRBF_CHECK_ARG(f);
    cudaStream_t s = (cudaStream_t)stream;
    const int W = f->p.width, H = f->p.height;
    HRBF_LAUNCH_PDL(depth_filter_metric_kernel, dim3(div_up(W, kBilTW), div_up(H, kBilTH)), dim3(256), 0, s, f->a, (const unsigned short*)f->tex[HRBF_FT_DEPTH_RAW],
                                                                                       (float*)f->tex[HRBF_FT_DEPTH_FILTERED], (float*)f->tex[HRBF_FT_DEPTH_METRIC],
                                                                                       (float*)f->tex[HRBF_FT_DEPTH_METRIC_FILTERED]);
    HRBF_LAUNCH_PDL(vertex_normal_radius_kernel, dim3(div_up(W, 32), div_up(H, 8)), dim3(256), 0, s, f->a, (const float*)f->tex[HRBF_FT_DEPTH_METRIC], (const float*)f->tex[HRBF_FT_DEPTH_METRIC_FILTERED],
                                                                               (float4*)f->tex[HRBF_FT_VERTEX_RAW], (float4*)f->tex[HRBF_FT_VERTEX_FILTERED],
                                                                               (float4*)f->tex[HRBF_FT_NORMAL_PCA], (float*)f->tex[HRBF_FT_RADIUS]);
    // computeCurvatureGradient writes NORMAL_OPT, which updateNormalRad copies into NORMAL: written to NORMAL directly
    HRBF_LAUNCH_PDL(curvature_gradient_kernel, dim3(div_up(W, 16), div_up(H, 8)), dim3(128), 0, s, f->a, (const float4*)f->tex[HRBF_FT_VERTEX_FILTERED], (const float4*)f->tex[HRBF_FT_NORMAL_PCA],
                                                                             (float4*)f->tex[HRBF_FT_PRINCIPAL_CURV1], (float4*)f->tex[HRBF_FT_PRINCIPAL_CURV2],
                                                                             (float*)f->tex[HRBF_FT_GRADIENT_MAG], (float4*)f->tex[HRBF_FT_NORMAL]);
    return HRBF_OK;
}
}  // extern "C"
static int frame_confidence_dev(hrbf_frame* f, const float* weighting_dev, cudaStream_t s)
{
    const dim3 b(32, 8);
    confidence_kernel<<<dim3(div_up(f->p.width, 32), div_up(f->p.height, 8)), b, 0, s>>>(f->a, (const float*)f->tex[HRBF_FT_GRADIENT_MAG], weighting_dev,
                                                                                      f->p.useConfEval, f->p.confEvalEpsilon, (float*)f->tex[HRBF_FT_CONFIDENCE]);
    HRBF_KERNEL_CHECK();
    return HRBF_OK;
}
extern "C" {
int hrbf_frame_vertex_confidence(hrbf_frame* f, float weighting, void* stream)
{
    HRBF_CHECK_ARG(f);
    cudaStream_t s = (cudaStream_t)stream;
    float* h = f->h_w + (f->slot++ & 7);
    *h = weighting;
    HRBF_CUDA(cudaMemcpyAsync(f->weighting, h, sizeof(float), cudaMemcpyHostToDevice, s));
    return frame_confidence_dev(f, f->weighting, s);
}
void* hrbf_frame_texture(hrbf_frame* f, int which)
{
    if (!f || which < 0 || which >= HRBF_FT_COUNT) return nullptr;
    return f->tex[which];
}

// ======================================================================== hrbf_fillin
}  // extern "C"
struct hrbf_fillin {
    const float* inline_weighting = nullptr;   // frame pipeline: device fusion weight (confidence evaluated in place)
    const unsigned int* skip_if_dense = nullptr; float dense_thresh = 0.75f;   // frame pipeline: see FillArgs::dense_count
    int width = 0, height = 0;
    char* slab = nullptr;
    void* tex[HRBF_FILL_COUNT] = {};
};
extern "C" {
int hrbf_fillin_create(hrbf_fillin** out, int width, int height)
{
    HRBF_CHECK_ARG(out && width > 0 && height > 0);
    hrbf_fillin* f = new (std::nothrow) hrbf_fillin();
    HRBF_CHECK_ARG(f != nullptr);
    f->width = width; f->height = height;
    const size_t P = (size_t)width * height;
    size_t off = 0, o[HRBF_FILL_COUNT];
    auto take = [&](size_t bytes) { size_t r = off; off += (bytes + 255) & ~(size_t)255; return r; };
    for (int t = 0; t < HRBF_FILL_COUNT; ++t) o[t] = take(t == HRBF_FILL_IMAGE || t == HRBF_FILL_ICPWEIGHT ? P * 4 : P * 16);
    if (cudaMalloc(&f->slab, off) != cudaSuccess) { set_error("cudaMalloc(%zu) failed", off); delete f; return HRBF_ERR_CUDA; }
    cudaMemset(f->slab, 0, off);
    for (int t = 0; t < HRBF_FILL_COUNT; ++t) f->tex[t] = f->slab + o[t];
    *out = f;
    return HRBF_OK;
}
int hrbf_fillin_destroy(hrbf_fillin* f)
{
    if (!f) return HRBF_OK;
    if (f->slab) cudaFree(f->slab);
    delete f;
    return HRBF_OK;
}
int hrbf_fillin_run(hrbf_fillin* f, hrbf_indexmap* im, hrbf_frame* fr, int passthrough, float lambda, float curvThr, void* stream)
{
    HRBF_CHECK_ARG(f && im && fr && im->width == f->width && im->height == f->height && fr->p.width == f->width && fr->p.height == f->height);
    FillArgs a;
    a.eVertex = (const float4*)im->tex[HRBF_TEX_VERTEX_HRBF]; a.eNormal = (const float4*)im->tex[HRBF_TEX_NORMAL_HRBF];
    a.eK1 = (const float4*)im->tex[HRBF_TEX_CURVK1_HRBF]; a.eK2 = (const float4*)im->tex[HRBF_TEX_CURVK2_HRBF];
    a.eIcpW = (const float*)im->tex[HRBF_TEX_ICPW_HRBF]; a.eImage = (const uchar4*)im->tex[HRBF_TEX_IMAGE_HRBF];
    a.vertexFiltered = (const float4*)fr->tex[HRBF_FT_VERTEX_FILTERED]; a.normal = (const float4*)fr->tex[HRBF_FT_NORMAL];
    a.k1 = (const float4*)fr->tex[HRBF_FT_PRINCIPAL_CURV1]; a.k2 = (const float4*)fr->tex[HRBF_FT_PRINCIPAL_CURV2];
    a.confidence = (const float*)fr->tex[HRBF_FT_CONFIDENCE]; a.rgb = (const unsigned char*)fr->tex[HRBF_FT_RGB];
    a.oVertex = (float4*)f->tex[HRBF_FILL_VERTEX]; a.oNormal = (float4*)f->tex[HRBF_FILL_NORMAL]; a.oK1 = (float4*)f->tex[HRBF_FILL_CURVK1];
    a.oK2 = (float4*)f->tex[HRBF_FILL_CURVK2]; a.oIcpW = (float*)f->tex[HRBF_FILL_ICPWEIGHT]; a.oImage = (uchar4*)f->tex[HRBF_FILL_IMAGE];
    a.n = f->width * f->height; a.passthrough = passthrough; a.lambda = lambda; a.curvThr = curvThr;
    a.weighting = (f->inline_weighting && !fr->p.useConfEval) ? f->inline_weighting : nullptr;
    a.cols = fr->p.width; a.rows = fr->p.height; a.cx = fr->p.cx; a.cy = fr->p.cy;
    a.dense_count = f->skip_if_dense; a.dense_thresh = f->dense_thresh;
    HRBF_LAUNCH_PDL(fill_in_kernel, dim3(div_up(a.n, 256)), dim3(256), 0, (cudaStream_t)stream, a);
    return HRBF_OK;
}
void* hrbf_fillin_texture(hrbf_fillin* f, int which)
{
    if (!f || which < 0 || which >= HRBF_FILL_COUNT) return nullptr;
    return f->tex[which];
}
int hrbf_dense_enough(const float* vertex, int width, int height, float thresh, int* dense, void* stream)
{
    HRBF_CHECK_ARG(vertex && dense && width >= 20 && height >= 20);
    cudaStream_t s = (cudaStream_t)stream;
    int* d = nullptr;
    HRBF_CUDA(cudaMalloc(&d, sizeof(int)));
    should_fill_kernel<<<1, 256, 0, s>>>((const float4*)vertex, height, width, thresh, d);
    count_launch();
    int fill = 0;
    cudaError_t e = cudaMemcpyAsync(&fill, d, sizeof(int), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(d);
    HRBF_CUDA(e);
    *dense = fill ? 0 : 1;
    return HRBF_OK;
}
}  // extern "C"

// ======================================================================== hrbf_model
struct hrbf_model {
    ModelArgs a{};
    PrepArgs pa{};
    int useConfEval = 0; float epsilon = 1000.f;
    unsigned int capacity = 0;
    float4* vbo[2] = { nullptr, nullptr };     // ping-pong surfel arrays (GlobalModel::vbos, target / renderSource)
    int cur = 0;
    char* slab = nullptr;
    unsigned int* count[2] = {};               // device count of vbo[k]
    unsigned int* overflow = nullptr;
    unsigned int* scan_ticket = nullptr;      // "last block done" counter of clean_flags_kernel (zero between launches)
    unsigned char* flags = nullptr;
    unsigned int *block_counts = nullptr, *block_offsets = nullptr, *winner = nullptr, *best = nullptr;
    unsigned char* update_id = nullptr;
    float4* staging = nullptr;
    float* pose = nullptr;       // device ring of 8 x (R[9], t[3]) and their inverses
    float* inv_pose = nullptr;
    float* active_kf = nullptr;
    float* h_stage = nullptr;    // pinned ring 8 x 24 floats
    float* h_kf = nullptr;
    unsigned int* h_count = nullptr;   // pinned
    int slot = 0;
    unsigned int bound = 0;      // host-side upper bound of the device count (grid sizing only)
    int n_slots = 0;
    int staged_time = -1;        // time of the fuse whose staging buffer clean() must append
    float* delta = nullptr;      // updateModel: device copy of the per-sub-map corrections
    size_t delta_bytes = 0;
    char* wintab = nullptr;      // device tables of the literal window loops (pa.wx / pa.wy point into it)
};

static int model_upload_pose(hrbf_model* m, const float* pose16, cudaStream_t s, const float** pose_dev, const float** inv_dev)
{
    const int k = m->slot++ & 7;
    float* h = m->h_stage + 24 * k;
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) h[i * 3 + j] = pose16[i * 4 + j]; h[9 + i] = pose16[i * 4 + 3]; }
    float* hi = h + 12;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) hi[i * 3 + j] = pose16[j * 4 + i];
    for (int i = 0; i < 3; ++i) hi[9 + i] = -(hi[i * 3] * pose16[3] + hi[i * 3 + 1] * pose16[7] + hi[i * 3 + 2] * pose16[11]);
    HRBF_CUDA(cudaMemcpyAsync(m->pose + 12 * k, h, 12 * sizeof(float), cudaMemcpyHostToDevice, s));
    HRBF_CUDA(cudaMemcpyAsync(m->inv_pose + 12 * k, hi, 12 * sizeof(float), cudaMemcpyHostToDevice, s));
    *pose_dev = m->pose + 12 * k; *inv_dev = m->inv_pose + 12 * k;
    return HRBF_OK;
}

namespace {
int model_initialise_dev(hrbf_model* m, const float* vertexMap, const float* normalMap, const unsigned char* rgb8, const float* curv1, const float* curv2,
                         const float* gradientMag, const float* pose_dev, cudaStream_t s)
{
    InitArgs ia;
    ia.vertexRaw = (const float4*)vertexMap; ia.normal = (const float4*)normalMap; ia.curv1 = (const float4*)curv1; ia.curv2 = (const float4*)curv2;
    ia.rgb = rgb8; ia.gradientMag = gradientMag; ia.pose = pose_dev; ia.useConfEval = m->useConfEval; ia.epsilon = m->epsilon;
    const int P = m->a.cols * m->a.rows, nb = div_up(P, kScanBlock);
    init_flags_kernel<<<nb, kScanBlock, 0, s>>>(m->a, ia, m->flags, m->block_counts);
    HRBF_KERNEL_CHECK();
    scan_blocks_kernel<<<1, 1024, 0, s>>>(m->block_counts, m->block_offsets, nullptr, (unsigned int)P, m->capacity, m->count[m->cur], m->overflow);
    HRBF_KERNEL_CHECK();
    init_scatter_kernel<<<nb, kScanBlock, 0, s>>>(m->a, ia, m->flags, m->block_offsets, m->capacity, m->vbo[m->cur]);
    HRBF_KERNEL_CHECK();
    m->bound = (unsigned int)P < m->capacity ? (unsigned int)P : m->capacity;
    m->staged_time = -1;
    return HRBF_OK;
}

int model_fuse_dev(hrbf_model* m, const float* pose_dev, int time, const unsigned char* rgb8, const float* depthRaw, const float* depthFiltered,
                   const float* curv1, const float* curv2, const float* confidence, const unsigned int* indexMap, const float* vertConf,
                   const float* normRad, float depthCutoff, int indexSubmap, cudaStream_t s, const float* inline_weighting = nullptr,
                   const float* normal_slot = nullptr)
{
    ModelArgs a = m->a;
    a.maxDepth = depthCutoff;
    FuseArgs f;
    f.rgb = rgb8; f.depthRaw = depthRaw; f.depthFiltered = depthFiltered; f.curv1 = (const float4*)curv1; f.curv2 = (const float4*)curv2;
    f.confidence = confidence; f.index = indexMap; f.vertConf = (const float4*)vertConf; f.normRad = (const float4*)normRad;
    f.pose = pose_dev; f.time = time; f.indexSubmap = (float)indexSubmap; f.weighting = inline_weighting;
    f.normal_slot = (const float4*)normal_slot;
    f.staging = m->staging; f.update_id = m->update_id; f.best = m->best; f.winner = m->winner;
    const int nb = div_up(m->n_slots, 128);
    HRBF_LAUNCH_PDL(fuse_associate_kernel, dim3(nb), dim3(128), 0, s, a, m->pa, f, m->count[m->cur]);
    HRBF_LAUNCH_PDL(fuse_merge_kernel, dim3(nb), dim3(128), 0, s, a, f, m->vbo[m->cur], m->count[m->cur]);
    m->staged_time = time;
    return HRBF_OK;
}

int model_clean_dev(hrbf_model* m, const float* inv_pose_dev, int time, const unsigned int* indexMap, const float* vertConf, const float* colorTime,
                    float confThreshold, float maxDepth, cudaStream_t s)
{
    ModelArgs a = m->a;
    a.maxDepth = maxDepth; a.confThreshold = confThreshold;
    CleanArgs c;
    c.index = indexMap; c.vertConf = (const float4*)vertConf; c.colorTime = (const float4*)colorTime; c.inv_pose = inv_pose_dev;
    c.active_kf = m->active_kf; c.kf_dim = HRBF_ACTIVE_KEYFRAME_DIMENSION; c.time = time;
    c.staging = m->staging; c.update_id = m->update_id;
    c.n_slots = (m->staged_time == time) ? m->n_slots : 0;           // the newUnstableVbo of THIS frame's fuse
    const unsigned int n_bound = m->bound + (unsigned int)c.n_slots;
    int nb = (int)((n_bound + kScanBlock - 1) / kScanBlock);
    if (nb < 1) nb = 1;
    const int grid = nb < kNumSMs * 8 ? nb : kNumSMs * 8;
    const int nxt = m->cur ^ 1;
    const ScanTail tail = { m->scan_ticket, m->block_offsets, m->capacity, m->count[nxt], m->overflow };      // the last block of the flags pass scans the block counts
    HRBF_LAUNCH_PDL(clean_flags_kernel, dim3(grid), dim3(kScanBlock), 0, s, a, c, m->vbo[m->cur], m->count[m->cur], m->flags, m->block_counts, tail);
    HRBF_LAUNCH_PDL(clean_scatter_kernel, dim3(grid), dim3(kScanBlock), 0, s, c, m->vbo[m->cur], m->count[m->cur], m->flags, m->block_offsets, m->capacity, m->vbo[nxt]);
    m->cur = nxt;
    m->bound = n_bound < m->capacity ? n_bound : m->capacity;
    m->staged_time = -1;
    return HRBF_OK;
}
}  // namespace

extern "C" {

int hrbf_model_create(hrbf_model** out, int width, int height, float cx, float cy, float fx, float fy, unsigned int capacity)
{
    HRBF_CHECK_ARG(out && width > 0 && height > 0);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device"); return HRBF_ERR_NO_DEVICE; }
    hrbf_model* m = new (std::nothrow) hrbf_model();
    HRBF_CHECK_ARG(m != nullptr);
    if (capacity == 0) capacity = 4596u * 4596u;
    m->capacity = capacity;
    ModelArgs& a = m->a;
    a.cols = width; a.rows = height; a.cx = cx; a.cy = cy; a.fx = fx; a.fy = fy;
    a.icx = (float)(1.0 / (double)fx); a.icy = (float)(1.0 / (double)fy);
    a.maxDepth = 20.f; a.confThreshold = 5.f; a.radiusMultiplier = 4.f; a.curvThr = 300.f; a.pca = 1; a.cleanWindow = 2;
    hrbf_frame_params fp{};
    fp.width = width; fp.height = height; fp.cx = cx; fp.cy = cy; fp.fx = fx; fp.fy = fy; fp.depthFactor = 1.f; fp.normalPCA = 1; fp.curvWindow = 3;
    m->pa = prep_args(fp);
    if (int rc = build_window_tables(m->pa, 3.0f, &m->wintab)) { delete m; return rc; }
    m->n_slots = fuse_slots_x(width) * fuse_slots_y(height);
    const size_t P = (size_t)width * height;
    const size_t items = (size_t)capacity + (size_t)m->n_slots > P ? (size_t)capacity + (size_t)m->n_slots : P;
    const size_t nblk = (items + kScanBlock - 1) / kScanBlock + 1;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t r = off; off += (bytes + 255) & ~(size_t)255; return r; };
    const size_t o_v0 = take((size_t)capacity * 80), o_v1 = take((size_t)capacity * 80), o_flags = take(items), o_bc = take(nblk * 4), o_bo = take(nblk * 4),
                 o_win = take((size_t)capacity * 4), o_best = take((size_t)m->n_slots * 4), o_uid = take((size_t)m->n_slots), o_stage = take((size_t)m->n_slots * 80),
                 o_pose = take(8 * 12 * 4), o_inv = take(8 * 12 * 4), o_kf = take(HRBF_ACTIVE_KEYFRAME_DIMENSION * 4), o_cnt = take(64);
    if (cudaMalloc(&m->slab, off) != cudaSuccess) { set_error("cudaMalloc(%zu) failed (surfel capacity %u)", off, capacity); delete m; return HRBF_ERR_CUDA; }
    m->vbo[0] = (float4*)(m->slab + o_v0); m->vbo[1] = (float4*)(m->slab + o_v1);
    m->flags = (unsigned char*)(m->slab + o_flags); m->block_counts = (unsigned int*)(m->slab + o_bc); m->block_offsets = (unsigned int*)(m->slab + o_bo);
    m->winner = (unsigned int*)(m->slab + o_win); m->best = (unsigned int*)(m->slab + o_best); m->update_id = (unsigned char*)(m->slab + o_uid);
    m->staging = (float4*)(m->slab + o_stage); m->pose = (float*)(m->slab + o_pose); m->inv_pose = (float*)(m->slab + o_inv);
    m->active_kf = (float*)(m->slab + o_kf);
    m->count[0] = (unsigned int*)(m->slab + o_cnt); m->count[1] = m->count[0] + 1; m->overflow = m->count[0] + 2; m->scan_ticket = m->count[0] + 3;
    // only the small control buffers need defined contents (the surfel arrays are written before they are read)
    cudaMemset(m->slab + o_flags, 0, off - o_flags);
    fill_u32_kernel<<<kNumSMs * 4, 256>>>(m->winner, (size_t)capacity, kNoWinner);
    cudaMallocHost(&m->h_stage, 8 * 24 * sizeof(float));
    cudaMallocHost(&m->h_kf, HRBF_ACTIVE_KEYFRAME_DIMENSION * sizeof(float));
    cudaMallocHost(&m->h_count, 4 * sizeof(unsigned int));
    memset(m->h_kf, 0, HRBF_ACTIVE_KEYFRAME_DIMENSION * sizeof(float));
    m->h_kf[0] = 1.0f;
    cudaMemcpy(m->active_kf, m->h_kf, HRBF_ACTIVE_KEYFRAME_DIMENSION * sizeof(float), cudaMemcpyHostToDevice);
    if (cudaDeviceSynchronize() != cudaSuccess || cudaGetLastError() != cudaSuccess) { set_error("model_create: CUDA setup failed"); hrbf_model_destroy(m); return HRBF_ERR_CUDA; }
    *out = m;
    return HRBF_OK;
}
int hrbf_model_destroy(hrbf_model* m)
{
    if (!m) return HRBF_OK;
    if (m->h_stage) cudaFreeHost(m->h_stage);
    if (m->h_kf) cudaFreeHost(m->h_kf);
    if (m->h_count) cudaFreeHost(m->h_count);
    if (m->delta) cudaFree(m->delta);
    if (m->slab) cudaFree(m->slab);
    if (m->wintab) cudaFree(m->wintab);
    delete m;
    return HRBF_OK;
}
int hrbf_model_set_params(hrbf_model* m, float radiusMultiplier, float curvValidThreshold, int normalPCA, int cleanWindow, int useConfEval, float confEvalEpsilon)
{
    HRBF_CHECK_ARG(m && cleanWindow >= 1 && cleanWindow <= 8);
    m->a.radiusMultiplier = radiusMultiplier; m->a.curvThr = curvValidThreshold; m->a.pca = normalPCA; m->a.cleanWindow = cleanWindow;
    m->pa.pca = normalPCA; m->useConfEval = useConfEval; m->epsilon = confEvalEpsilon;
    return HRBF_OK;
}
int hrbf_model_set_active_keyframes(hrbf_model* m, const int* ids, int n, void* stream)
{
    HRBF_CHECK_ARG(m && (ids || n == 0) && n >= 0);
    cudaStream_t s = (cudaStream_t)stream;
    HRBF_CUDA(cudaStreamSynchronize(s));
    memset(m->h_kf, 0, HRBF_ACTIVE_KEYFRAME_DIMENSION * sizeof(float));
    for (int i = 0; i < n; ++i) { HRBF_CHECK_ARG(ids[i] >= 0 && ids[i] < HRBF_ACTIVE_KEYFRAME_DIMENSION); m->h_kf[ids[i]] = 1.0f; }
    HRBF_CUDA(cudaMemcpyAsync(m->active_kf, m->h_kf, HRBF_ACTIVE_KEYFRAME_DIMENSION * sizeof(float), cudaMemcpyHostToDevice, s));
    return HRBF_OK;
}
int hrbf_model_initialise(hrbf_model* m, const float* vertexMap, const float* normalMap, const unsigned char* rgb8, const float* curv1, const float* curv2,
                          const float* gradientMag, const float* pose16, void* stream)
{
    HRBF_CHECK_ARG(m && vertexMap && normalMap && rgb8 && curv1 && curv2 && pose16 && (gradientMag || !m->useConfEval));
    const float *pd, *pi;
    if (int rc = model_upload_pose(m, pose16, (cudaStream_t)stream, &pd, &pi)) return rc;
    return model_initialise_dev(m, vertexMap, normalMap, rgb8, curv1, curv2, gradientMag, pd, (cudaStream_t)stream);
}
int hrbf_model_fuse(hrbf_model* m, const float* pose16, int time, const unsigned char* rgb8, const float* depthRaw, const float* depthFiltered,
                    const float* curv1, const float* curv2, const float* confidence, const unsigned int* indexMap, const float* vertConf,
                    const float* colorTime, const float* normRad, float depthCutoff, float confThreshold, float weighting, int insertSubmap,
                    int indexSubmap, void* stream)
{
    (void)colorTime; (void)confThreshold; (void)weighting; (void)insertSubmap;    // uniforms data.vert declares but never reads
    HRBF_CHECK_ARG(m && pose16 && rgb8 && depthRaw && depthFiltered && curv1 && curv2 && confidence && indexMap && vertConf && normRad);
    const float *pd, *pi;
    if (int rc = model_upload_pose(m, pose16, (cudaStream_t)stream, &pd, &pi)) return rc;
    return model_fuse_dev(m, pd, time, rgb8, depthRaw, depthFiltered, curv1, curv2, confidence, indexMap, vertConf, normRad, depthCutoff, indexSubmap, (cudaStream_t)stream);
}
int hrbf_model_clean(hrbf_model* m, const float* pose16, int time, const unsigned int* indexMap, const float* vertConf, const float* colorTime,
                     const float* normRad, const float* depthMap, float confThreshold, float maxDepth, void* stream)
{
    (void)normRad; (void)depthMap;
    HRBF_CHECK_ARG(m && pose16 && indexMap && vertConf && colorTime);
    const float *pd, *pi;
    if (int rc = model_upload_pose(m, pose16, (cudaStream_t)stream, &pd, &pi)) return rc;
    return model_clean_dev(m, pi, time, indexMap, vertConf, colorTime, confThreshold, maxDepth, (cudaStream_t)stream);
}
int hrbf_model_set_model(hrbf_model* m, const float* surfels, unsigned int count, int host, void* stream)
{
    HRBF_CHECK_ARG(m && (surfels || count == 0) && count <= m->capacity);
    cudaStream_t s = (cudaStream_t)stream;
    HRBF_CUDA(cudaStreamSynchronize(s));
    if (count) HRBF_CUDA(cudaMemcpyAsync(m->vbo[m->cur], surfels, (size_t)count * 80, host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, s));
    m->h_count[2] = count;
    HRBF_CUDA(cudaMemcpyAsync(m->count[m->cur], m->h_count + 2, sizeof(unsigned int), cudaMemcpyHostToDevice, s));
    HRBF_CUDA(cudaStreamSynchronize(s));
    m->bound = count;
    m->staged_time = -1;
    return HRBF_OK;
}
int hrbf_model_update_model(hrbf_model* m, const float* delta16_host, int n_delta, void* stream)
{
    HRBF_CHECK_ARG(m && delta16_host && n_delta > 0 && n_delta <= 4096);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t bytes = (size_t)n_delta * 16 * sizeof(float);
    if (bytes > m->delta_bytes) {
        if (m->delta) { HRBF_CUDA(cudaStreamSynchronize(s)); cudaFree(m->delta); m->delta = nullptr; m->delta_bytes = 0; }
        HRBF_CUDA(cudaMalloc(&m->delta, bytes));
        m->delta_bytes = bytes;
    }
    HRBF_CUDA(cudaMemcpyAsync(m->delta, delta16_host, bytes, cudaMemcpyHostToDevice, s));
    HRBF_CUDA(cudaStreamSynchronize(s));          // pageable source: the caller may reuse its buffer on return
    if (m->bound > 0) {
        int blocks = (int)((m->bound + 255u) / 256u);
        if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
        update_model_kernel<<<blocks, 256, 0, s>>>(m->vbo[m->cur], m->count[m->cur], m->delta, n_delta);
        HRBF_KERNEL_CHECK();
    }
    return HRBF_OK;
}
const float* hrbf_model_model(hrbf_model* m) { return m ? (const float*)m->vbo[m->cur] : nullptr; }
const unsigned int* hrbf_model_count_dev(hrbf_model* m) { return m ? m->count[m->cur] : nullptr; }
int hrbf_model_last_count(hrbf_model* m, unsigned int* count, void* stream)
{
    HRBF_CHECK_ARG(m && count);
    cudaStream_t s = (cudaStream_t)stream;
    HRBF_CUDA(cudaMemcpyAsync(m->h_count, m->count[m->cur], sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
    HRBF_CUDA(cudaStreamSynchronize(s));
    *count = m->h_count[0];
    m->bound = *count;          // exact again
    return HRBF_OK;
}
// The window a shader's FLOAT-counter loop really visits along one axis (geometry.glsl:198-212, depth_curvature_gradient.frag:62-75):
//     for (float i = tx_min; i <= tx_max; i += 1 / n)     tx_min/max = clamp(texcoord -/+ win / n, 0, 1)
// For every pixel p of an axis of n texels: the first texel, the number of samples and each sample's coordinate i * n, evaluated
// with the shader's own fp32 expressions (the accumulated counter overshoots tx_max for ~40 % of the pixels: 6 samples, not 7).
// uv_vbo_coords: 0 = the texcoord of a full-screen fragment pass, (p + 0.5) / n; 1 = the uv VBO of GlobalModel::fuse
// (GlobalModel.cpp:87-96), float(p) / n + 1.0 / (2 n) summed in double.  Texel selection: GL_NEAREST on a coordinate held with 8
// fractional bits.  Host code; it builds the tables of the literal-window kernels and is tested against the
// oracle's literal loops on the CPU.
int hrbf_window_table(int n, float win, int uv_vbo_coords, int* first, int* count, float* coords /* [n][HRBF_WINDOW_MAX] */)
{
    HRBF_CHECK_ARG(n > 0 && win >= 0.f && win <= 3.5f && first && count && coords);
    const float fn = (float)n, step = 1.0f / fn;
    for (int p = 0; p < n; ++p) {
        volatile float tc = uv_vbo_coords ? (float)((double)((float)p / fn) + 1.0 / (double)(2 * fn)) : ((float)p + 0.5f) / fn;
        volatile float sw = step * win;                       // volatile: every intermediate is rounded to fp32, no contraction
        volatile float lo = tc - sw, hi = tc + sw;
        if (lo < 0.0f) lo = 0.0f;
        if (hi > 1.0f) hi = 1.0f;
        int k = 0;
        first[p] = 0;
        for (volatile float i = lo; i <= hi && k < HRBF_WINDOW_MAX; i = i + step) {
            volatile float scaled = i * fn;
            volatile float fixed = floorf(scaled * 256.0f + 0.5f);
            int t = (int)floorf(fixed / 256.0f);
            t = t < 0 ? 0 : (t >= n ? n - 1 : t);
            if (k == 0) first[p] = t;
            else if (t != first[p] + k) { set_error("window_table: samples of pixel %d are not consecutive texels", p); return HRBF_ERR_INVALID_ARG; }
            coords[(size_t)p * HRBF_WINDOW_MAX + k] = scaled;
            ++k;
        }
        count[p] = k;
        for (int q = k; q < HRBF_WINDOW_MAX; ++q) coords[(size_t)p * HRBF_WINDOW_MAX + q] = 0.0f;
    }
    return HRBF_OK;
}
// GlobalModel::downloadMap (GlobalModel.cpp:775-804): the current surfel array to the host, count records of 20 floats
int hrbf_model_download_map(hrbf_model* m, float* surfels_host, unsigned int max_count, unsigned int* count_out, void* stream)
{
    HRBF_CHECK_ARG(m && count_out && (surfels_host || max_count == 0));
    cudaStream_t s = (cudaStream_t)stream;
    if (int rc = hrbf_model_last_count(m, count_out, stream)) return rc;
    if (max_count == 0) return HRBF_OK;                      // size query
    if (*count_out > max_count) { set_error("download_map: %u surfels do not fit a buffer of %u", *count_out, max_count); return HRBF_ERR_CAPACITY; }
    if (*count_out) HRBF_CUDA(cudaMemcpyAsync(surfels_host, m->vbo[m->cur], (size_t)*count_out * 80, cudaMemcpyDeviceToHost, s));
    HRBF_CUDA(cudaStreamSynchronize(s));
    return HRBF_OK;
}
// HRBFFusion::savePly (HRBFFusion.cpp:1737-1853): header text for n vertices; returns its length (0 if buf is too small)
size_t hrbf_ply_header(unsigned int n_vertices, char* buf, size_t buf_len)
{
    if (!buf) return 0;
    const int n = snprintf(buf, buf_len,
                           "ply\nformat binary_little_endian 1.0\nelement vertex %u\nproperty float x\nproperty float y\nproperty float z"
                           "\nproperty uchar red\nproperty uchar green\nproperty uchar blue\nproperty float nx\nproperty float ny\nproperty float nz"
                           "\nproperty float curvature_max\nproperty float curvature_min\nproperty float radius\nproperty float submapIndex\nend_header\n",
                           n_vertices);
    return (n > 0 && (size_t)n < buf_len) ? (size_t)n : 0;
}
// The vertex block of savePly: surfels with confidence > confThreshold packed to 43-byte records on the device (in the idle
// ping-pong buffer), then copied to records_host.  records_host == NULL: only the vertex count is returned (size query).
int hrbf_model_export_ply(hrbf_model* m, float confThreshold, void* records_host, size_t host_bytes, unsigned int* n_vertices_out, void* stream)
{
    HRBF_CHECK_ARG(m && n_vertices_out);
    cudaStream_t s = (cudaStream_t)stream;
    *n_vertices_out = 0;
    if (m->bound == 0) return HRBF_OK;
    int nb = (int)((m->bound + kScanBlock - 1) / kScanBlock);
    const int grid = nb < kNumSMs * 8 ? nb : kNumSMs * 8;
    unsigned char* packed = (unsigned char*)m->vbo[m->cur ^ 1];        // 43 B <= 80 B per surfel: always fits
    unsigned int* total = m->count[m->cur ^ 1];                        // rewritten by the next clean before it is read
    ply_flags_kernel<<<grid, kScanBlock, 0, s>>>(m->vbo[m->cur], m->count[m->cur], confThreshold, m->flags, m->block_counts);
    HRBF_KERNEL_CHECK();
    scan_blocks_kernel<<<1, 1024, 0, s>>>(m->block_counts, m->block_offsets, m->count[m->cur], 0u, 0xffffffffu, total, m->overflow);
    HRBF_KERNEL_CHECK();
    HRBF_CUDA(cudaMemcpyAsync(m->h_count + 3, total, sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
    if (records_host) {
        ply_scatter_kernel<<<grid, kScanBlock, 0, s>>>(m->vbo[m->cur], m->count[m->cur], m->flags, m->block_offsets, packed);
        HRBF_KERNEL_CHECK();
    }
    HRBF_CUDA(cudaStreamSynchronize(s));
    const unsigned int n = m->h_count[3];
    *n_vertices_out = n;
    if (!records_host || n == 0) return HRBF_OK;
    if ((size_t)n * kPlyVertexBytes > host_bytes) { set_error("export_ply: %u vertices need %zu bytes, buffer has %zu", n, (size_t)n * kPlyVertexBytes, host_bytes); return HRBF_ERR_CAPACITY; }
    HRBF_CUDA(cudaMemcpyAsync(records_host, packed, (size_t)n * kPlyVertexBytes, cudaMemcpyDeviceToHost, s));
    HRBF_CUDA(cudaStreamSynchronize(s));
    return HRBF_OK;
}
int hrbf_model_overflowed(hrbf_model* m, int* flag, void* stream)
{
    HRBF_CHECK_ARG(m && flag);
    cudaStream_t s = (cudaStream_t)stream;
    HRBF_CUDA(cudaMemcpyAsync(m->h_count + 1, m->overflow, sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
    HRBF_CUDA(cudaStreamSynchronize(s));
    *flag = (int)m->h_count[1];
    return HRBF_OK;
}

}  // extern "C"

// ======================================================================== hrbf_fusion
struct hrbf_fusion {
    hrbf_fusion_params p{};
    hrbf_odometry* odom = nullptr;
    hrbf_indexmap* im = nullptr;
    hrbf_model* model = nullptr;
    hrbf_frame* frame = nullptr;     // the buffer of the frame being / last processed (= frames[cur] while processing)
    hrbf_frame* frames[2] = { nullptr, nullptr };   // ping-pong: frame t+1 is uploaded + preprocessed while frame t is tracked and fused
    int cur = 0;                     // buffer the next processed frame uses
    bool staged[2] = { false, false };
    cudaStream_t pre_stream = nullptr;              // staging stream (lowest priority)
    cudaStream_t so3_stream = nullptr;              // the staged SO3 pre-alignment, beside the preprocessing (lowest priority)
    cudaEvent_t ev_up = nullptr, ev_so3[2] = {};    // upload done (forks so3_stream) / SO3 pre-alignment of bank k done (joins)
    cudaEvent_t ev_img[2] = {};                     // the intensity pyramid of bank k is written (so3_stream; read by the staging stream)
    cudaEvent_t ev_track = nullptr;                 // the model-side pyramids of the frame being processed are built: its tracker starts
    bool ev_track_valid = false;
    bool ev_so3_valid[2] = { false, false };
    // tracker-input banks of the odometry (CurrBank): frame number k (1-based) uses bank k & 1, whichever frame buffer it sits in.
    // ev_curr[k]: bank k is built (the next frame's SO3 pre-alignment reads its image); ev_bank_free[k]: the frame that used it is done
    cudaEvent_t ev_staged[2] = {}, ev_free[2] = {}, ev_curr[2] = {}, ev_bank_free[2] = {};
    bool ev_free_valid[2] = { false, false }, ev_curr_valid[2] = { false, false }, ev_bank_free_valid[2] = { false, false };
    hrbf_fillin* fill = nullptr;
    int tick = 1;
    int indexSubmap = 0;
    float* dev = nullptr;            // [0..11] currPose, [12..23] lastPose, [24..35] inverse, [36] weighting, [40] shouldFill (int)
    float* traj = nullptr;           // device float[traj_cap][12]
    int traj_cap = 0, traj_n = 0;
    float* h_pose = nullptr;         // pinned 16
    bool timings = false;
    cudaEvent_t ev[5] = {};
    float last_ms[4] = { 0, 0, 0, 0 };
};

__global__ void set_identity_pose_kernel(float* p)
{
    if (threadIdx.x < 12) p[threadIdx.x] = (threadIdx.x == 0 || threadIdx.x == 4 || threadIdx.x == 8) ? 1.f : 0.f;
}

// Everything the tracker needs from the camera frame in buffer b alone (current-frame pyramids, Sobel + candidates, SO3 pre-alignment
// against the previous camera image = the other bank), on stream s: the staging stream for staged frames, else the main stream.
// the SO3 pre-alignment of the frame in buffer b (frame_number) against the previous camera frame, on stream s
static int stage_so3(hrbf_fusion* F, int b, int frame_number, cudaStream_t s)
{
    const int k = frame_number & 1;
    if (F->ev_bank_free_valid[k]) HRBF_CUDA(cudaStreamWaitEvent(s, F->ev_bank_free[k], 0));   // frame_number - 2 read this bank's result
    if (F->ev_so3_valid[k ^ 1]) HRBF_CUDA(cudaStreamWaitEvent(s, F->ev_so3[k ^ 1], 0));        // the other bank's image: written, and no longer compared with this bank's old one
    if (int rc = odom_stage_so3_dev(F->odom, k, (const unsigned char*)F->frames[b]->tex[HRBF_FT_RGB], F->p.so3 != 0, frame_number != 1, F->ev_img[k], s)) return rc;
    HRBF_CUDA(cudaEventRecord(F->ev_so3[k], s));
    F->ev_so3_valid[k] = true;
    return HRBF_OK;
}
static int stage_current(hrbf_fusion* F, int b, int frame_number, cudaStream_t s)
{
    hrbf_frame* fr = F->frames[b];
    const int k = frame_number & 1;
    OdomPrepInputs in;
    memset(&in, 0, sizeof in);
    in.vc = (const float*)fr->tex[HRBF_FT_VERTEX_FILTERED]; in.nc = (const float*)fr->tex[HRBF_FT_NORMAL];
    in.k1c = (const float*)fr->tex[HRBF_FT_PRINCIPAL_CURV1]; in.k2c = (const float*)fr->tex[HRBF_FT_PRINCIPAL_CURV2];
    in.rgba_c = (const unsigned char*)fr->tex[HRBF_FT_RGBA]; in.rgb8_c = (const unsigned char*)fr->tex[HRBF_FT_RGB];
    if (F->ev_bank_free_valid[k]) HRBF_CUDA(cudaStreamWaitEvent(s, F->ev_bank_free[k], 0));   // frame_number - 2 tracked from this bank
    HRBF_CUDA(cudaStreamWaitEvent(s, F->ev_img[k], 0));      // recorded by this frame's stage_so3 (always called first)
    if (int rc = odom_stage_current_dev(F->odom, k, in, s)) return rc;
    fr->fuse_normals_time = -1;
    if (frame_number > 1 && !F->p.rgbOnly && F->model->a.pca) {      // the PCA normals GlobalModel::fuse will ask for (data.vert)
        ModelArgs a = F->model->a;
        a.maxDepth = F->p.maxDepthProcessed;
        HRBF_LAUNCH_PDL(fuse_normals_kernel, dim3(div_up(fuse_slots_x(a.cols), kFnTX), div_up(fuse_slots_y(a.rows), kFnTY)), dim3(kFnTX * kFnTY), 0, s, a, F->model->pa, (const float*)fr->tex[HRBF_FT_DEPTH_METRIC],
                        (const float*)fr->tex[HRBF_FT_DEPTH_METRIC_FILTERED], (const float4*)fr->tex[HRBF_FT_PRINCIPAL_CURV1], (const float4*)fr->tex[HRBF_FT_PRINCIPAL_CURV2],
                        frame_number, fr->fuse_normals);
        fr->fuse_normals_time = frame_number;
    }
    HRBF_CUDA(cudaEventRecord(F->ev_curr[k], s));
    F->ev_curr_valid[k] = true;
    return HRBF_OK;
}

// Everything of a frame that needs nothing but the uploaded camera frame, on stream s (the staging stream, or the caller's stream for
// an unstaged frame); the SO3 pre-alignment -- a chain of <= 10 small dependent reductions -- runs beside it on its own stream.
static int stage_all(hrbf_fusion* F, int b, int frame_number, cudaStream_t s)
{
    HRBF_CUDA(cudaEventRecord(F->ev_up, s));
    HRBF_CUDA(cudaStreamWaitEvent(F->so3_stream, F->ev_up, 0));
    if (int rc = stage_so3(F, b, frame_number, F->so3_stream)) return rc;
    if (int rc = hrbf_frame_preprocess(F->frames[b], s)) return rc;
    if (int rc = stage_current(F, b, frame_number, s)) return rc;
    HRBF_CUDA(cudaStreamWaitEvent(s, F->ev_so3[frame_number & 1], 0));
    return HRBF_OK;
}

// preprocessed = true: frames[cur] was uploaded and preprocessed by hrbf_fusion_stage_frame (on the staging stream)
static int fusion_frame(hrbf_fusion* F, float weightMultiplier, cudaStream_t s, bool preprocessed = false)
{
    const hrbf_fusion_params& p = F->p;
    hrbf_frame* fr = F->frames[F->cur];
    F->frame = fr;
    hrbf_indexmap* im = F->im;
    hrbf_model* M = F->model;
    float* currPose = F->dev, *lastPose = F->dev + 12, *invPose = F->dev + 24, *weighting = F->dev + 36;
    unsigned int* denseCount = (unsigned int*)(F->dev + 44);      // two counters, by frame parity
    auto FT = [&](int t) { return fr->tex[t]; };
    auto IT = [&](int t) { return im->tex[t]; };
    auto LT = [&](int t) { return F->fill->tex[t]; };
    auto mark = [&](int k) { if (F->timings) cudaEventRecord(F->ev[k], s); };
    auto splat = [&](int out_mask) { return indexmap_splat(im, invPose, (const float*)M->vbo[M->cur], M->count[M->cur], M->bound, p.maxDepthProcessed, s, out_mask); };

    mark(0);
    if (!preprocessed) {
        if (int rc = stage_all(F, F->cur, F->tick, s)) return rc;
    }
    odom_select_bank(F->odom, F->tick & 1);
    mark(1);
    // a staged frame is needed from here on: by the map initialisation (first frame), else by the tracker -- the model-side pyramids
    // before it read the previous prediction only, so the staging streams get the whole frame period
    auto wait_staged = [&]() { return preprocessed ? cudaStreamWaitEvent(s, F->ev_staged[F->cur], 0) : cudaSuccess; };
    if (F->tick == 1) {
        HRBF_CUDA(wait_staged());
        set_identity_pose_kernel<<<1, 32, 0, s>>>(currPose);
        HRBF_KERNEL_CHECK();
        if (int rc = model_initialise_dev(M, (const float*)FT(HRBF_FT_VERTEX_RAW), (const float*)FT(HRBF_FT_NORMAL), (const unsigned char*)FT(HRBF_FT_RGB),
                                          (const float*)FT(HRBF_FT_PRINCIPAL_CURV1), (const float*)FT(HRBF_FT_PRINCIPAL_CURV2), (const float*)FT(HRBF_FT_GRADIENT_MAG),
                                          currPose, s)) return rc;
        // (initFirstRGB, HRBFFusion.cpp:1058: this frame's image pyramid is in its bank, where the next frame's SO3 pre-alignment reads it)
        // VertexConfidence is not run on the first frame (HRBFFusion.cpp:1028-1126): CONFIDENCE keeps its initial zeros
        mark(2); mark(3);
    } else {
        // ---- Registration (HRBFFusion.cpp:1063-1109) ----
        {   // the reference's 7 init* calls (HRBFFusion.cpp:1073-1099) as one launch
            OdomPrepInputs in;
            in.vm = (const float*)IT(HRBF_TEX_VERTEX_HRBF); in.nm = (const float*)IT(HRBF_TEX_NORMAL_HRBF);
            in.vm_alt = (const float*)LT(HRBF_FILL_VERTEX); in.nm_alt = (const float*)LT(HRBF_FILL_NORMAL);
            in.k1m = (const float*)IT(HRBF_TEX_CURVK1_HRBF); in.k2m = (const float*)IT(HRBF_TEX_CURVK2_HRBF);
            in.k1m_alt = (const float*)LT(HRBF_FILL_CURVK1); in.k2m_alt = (const float*)LT(HRBF_FILL_CURVK2);
            in.w = (const float*)IT(HRBF_TEX_ICPW_HRBF); in.w_alt = (const float*)LT(HRBF_FILL_ICPWEIGHT);
            in.rgba_m = (const unsigned char*)IT(HRBF_TEX_IMAGE_HRBF); in.rgba_m_alt = (const unsigned char*)LT(HRBF_FILL_IMAGE);
            in.vc = (const float*)FT(HRBF_FT_VERTEX_FILTERED); in.nc = (const float*)FT(HRBF_FT_NORMAL);
            in.k1c = (const float*)FT(HRBF_FT_PRINCIPAL_CURV1); in.k2c = (const float*)FT(HRBF_FT_PRINCIPAL_CURV2);
            in.rgba_c = (const unsigned char*)FT(HRBF_FT_RGBA);
            // denseEnough (HRBFFusion.cpp:1069-1070): the previous prediction counted its samples, no separate reduction kernel
            in.sel = nullptr; in.dense_count = denseCount + ((F->tick - 1) & 1); in.dense_count_reset = denseCount + (F->tick & 1);
            in.dense_thresh = p.denseEnoughThresh; in.pose_dev = currPose;
            in.rgb8_c = (const unsigned char*)FT(HRBF_FT_RGB);
            if (int rc = odom_prep_all_dev(F->odom, in, s)) return rc;
            // the next frame's staged kernels wait for this point: they fit beside the (latency-bound) tracker, whereas beside the
            // pyramid kernel they would only delay the tracker's start
            HRBF_CUDA(cudaEventRecord(F->ev_track, s));
            F->ev_track_valid = true;
            HRBF_CUDA(wait_staged());
        }
        {
            // the persistent tracker also writes lastPose, the inverse pose, the fusion weight and the trajectory row
            const OdomFrameEpilogue ep = { lastPose, invPose, weighting, weightMultiplier, F->traj_n < F->traj_cap ? F->traj + 12 * (size_t)F->traj_n : nullptr };
            const int rc = odom_track_frame_dev(F->odom, currPose, ep, p.rgbOnly != 0, p.icpWeight, p.pyramid != 0, p.fastOdom != 0, p.so3 != 0, p.weightedICP != 0, s);
            if (rc < 0) return rc;
            if (rc == 1) {      // kernel-graph tracker: the separate calls
                HRBF_CUDA(cudaMemcpyAsync(lastPose, currPose, 12 * sizeof(float), cudaMemcpyDeviceToDevice, s));
                if (int rc2 = hrbf_odometry_track_async(F->odom, lastPose, currPose, p.rgbOnly, p.icpWeight, p.pyramid, p.fastOdom, p.so3, p.weightedICP, s)) return rc2;
                velocity_weighting_kernel<<<1, 32, 0, s>>>(currPose, lastPose, weightMultiplier, weighting);
                HRBF_KERNEL_CHECK();
                pose_inverse_kernel<<<1, 32, 0, s>>>(currPose, invPose);
                HRBF_KERNEL_CHECK();
                if (ep.traj_out) HRBF_CUDA(cudaMemcpyAsync(ep.traj_out, currPose, 12 * sizeof(float), cudaMemcpyDeviceToDevice, s));
            }
        }
        // VertexConfidence (HRBFFusion.cpp:1124): evaluated in place by fuse and fill-in; the texture only with preprocessingUseConfEval
        const float* inline_w = p.frame.useConfEval ? nullptr : weighting;
        F->fill->inline_weighting = inline_w;
        if (!inline_w) { if (int rc = frame_confidence_dev(fr, weighting, s)) return rc; }
        mark(2);
        // ---- Integration (HRBFFusion.cpp:1192-1227) ----
        if (!p.rgbOnly) {
            if (int rc = splat(2)) return rc;                  // fuse reads index, vertConf, normRad (data.vert)
            if (int rc = model_fuse_dev(M, currPose, F->tick, (const unsigned char*)FT(HRBF_FT_RGB), (const float*)FT(HRBF_FT_DEPTH_METRIC),
                                        (const float*)FT(HRBF_FT_DEPTH_METRIC_FILTERED), (const float*)FT(HRBF_FT_PRINCIPAL_CURV1), (const float*)FT(HRBF_FT_PRINCIPAL_CURV2),
                                        (const float*)FT(HRBF_FT_CONFIDENCE), (const unsigned int*)IT(HRBF_TEX_INDEX), (const float*)IT(HRBF_TEX_VERTCONF),
                                        (const float*)IT(HRBF_TEX_NORMRAD), p.maxDepthProcessed, F->indexSubmap, s, inline_w,
                                        fr->fuse_normals_time == F->tick ? (const float*)fr->fuse_normals : nullptr)) return rc;
            if (int rc = splat(1)) return rc;                  // clean reads index, vertConf, colorTime (copy_unstable.vert)
            if (int rc = model_clean_dev(M, invPose, F->tick, (const unsigned int*)IT(HRBF_TEX_INDEX), (const float*)IT(HRBF_TEX_VERTCONF),
                                         (const float*)IT(HRBF_TEX_COLORTIME), p.confidenceThreshold, p.maxDepthProcessed, s)) return rc;
        }
        mark(3);
    }
    if (F->tick == 1) { pose_inverse_kernel<<<1, 32, 0, s>>>(currPose, invPose); HRBF_KERNEL_CHECK(); }
    // ---- Prediction (HRBFFusion.cpp:1244-1260) ----
    if (int rc = splat(7)) return rc;
    im->dense_count_next = denseCount + (F->tick & 1);
    F->fill->skip_if_dense = denseCount + (F->tick & 1); F->fill->dense_thresh = p.denseEnoughThresh;
    if (int rc = hrbf_indexmap_predict_hrbf(im, 0, p.predWindow, p.predMinNeighbors, p.predMaxNeighbors, p.predConfThreshold, p.icpWeightLambda, s)) return rc;
    if (int rc = hrbf_fillin_run(F->fill, im, fr, 0, p.icpWeightLambda, p.curvValidThreshold, s)) return rc;
    mark(4);
    if (F->tick == 1 && F->traj_n < F->traj_cap) HRBF_CUDA(cudaMemcpyAsync(F->traj + 12 * (size_t)F->traj_n, currPose, 12 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    ++F->traj_n;
    // the buffer and the bank are free for the frame after next once everything above has run
    HRBF_CUDA(cudaEventRecord(F->ev_bank_free[F->tick & 1], s));
    F->ev_bank_free_valid[F->tick & 1] = true;
    ++F->tick;
    HRBF_CUDA(cudaEventRecord(F->ev_free[F->cur], s));
    F->ev_free_valid[F->cur] = true;
    return HRBF_OK;
}

extern "C" {

void hrbf_fusion_default_params(hrbf_fusion_params* p, int width, int height, float cx, float cy, float fx, float fy)
{
    if (!p) return;
    memset(p, 0, sizeof *p);
    p->frame.width = width; p->frame.height = height; p->frame.cx = cx; p->frame.cy = cy; p->frame.fx = fx; p->frame.fy = fy;
    p->frame.depthFactor = 1.0f / 5000.0f; p->frame.depthCutoff = 3.5f; p->frame.radiusMultiplier = 4.0f; p->frame.normalPCA = 1;
    p->frame.curvWindow = 3; p->frame.bilateral = 1; p->frame.useConfEval = 0; p->frame.confEvalEpsilon = 1000.0f;
    p->confidenceThreshold = 5.0f; p->maxDepthProcessed = 20.0f; p->icpWeight = 10.0f;
    p->rgbOnly = 0; p->pyramid = 1; p->fastOdom = 0; p->so3 = 1; p->weightedICP = 1;
    p->predWindow = 3; p->predMinNeighbors = 6; p->predMaxNeighbors = 10; p->predConfThreshold = 3.0f;
    p->icpWeightLambda = 10.0f; p->curvValidThreshold = 300.0f; p->denseEnoughThresh = 0.75f; p->cleanWindow = 2;
    p->capacity = 0;
}

int hrbf_fusion_create(hrbf_fusion** out, const hrbf_fusion_params* p)
{
    HRBF_CHECK_ARG(out && p);
    hrbf_fusion* F = new (std::nothrow) hrbf_fusion();
    HRBF_CHECK_ARG(F != nullptr);
    F->p = *p;
    const hrbf_frame_params& fp = p->frame;
    int rc = hrbf_frame_create(&F->frames[0], &fp);
    if (!rc) rc = hrbf_frame_create(&F->frames[1], &fp);
    F->frame = F->frames[0];
    if (!rc) rc = hrbf_fillin_create(&F->fill, fp.width, fp.height);
    if (!rc) rc = hrbf_indexmap_create(&F->im, fp.width, fp.height, fp.cx, fp.cy, fp.fx, fp.fy);
    if (!rc) rc = hrbf_model_create(&F->model, fp.width, fp.height, fp.cx, fp.cy, fp.fx, fp.fy, p->capacity);
    if (!rc) rc = hrbf_model_set_params(F->model, fp.radiusMultiplier, p->curvValidThreshold, fp.normalPCA, p->cleanWindow, fp.useConfEval, fp.confEvalEpsilon);
    if (!rc) rc = hrbf_odometry_create(&F->odom, fp.width, fp.height, fp.cx, fp.cy, fp.fx, fp.fy, 0.10f, sinf(20.f * 3.14159265f / 180.f));
    if (!rc) rc = hrbf_odometry_set_params(F->odom, p->curvValidThreshold, 0, 2, 0);
    if (!rc) rc = hrbf_odometry_set_tracker_threads(F->odom, p->trackerThreads ? p->trackerThreads : 384);
    if (!rc) rc = odom_enable_banks(F->odom);
    if (!rc) {
        F->traj_cap = 1 << 16;
        if (cudaMalloc(&F->dev, 64 * sizeof(float)) != cudaSuccess || cudaMalloc(&F->traj, (size_t)F->traj_cap * 12 * sizeof(float)) != cudaSuccess ||
            cudaMallocHost(&F->h_pose, 16 * sizeof(float)) != cudaSuccess) { set_error("fusion_create: allocation failed"); rc = HRBF_ERR_CUDA; }
        else {
            cudaMemset(F->dev, 0, 64 * sizeof(float));
            for (auto& e : F->ev) cudaEventCreate(&e);
            for (int k = 0; k < 2; ++k) { cudaEventCreateWithFlags(&F->ev_staged[k], cudaEventDisableTiming); cudaEventCreateWithFlags(&F->ev_free[k], cudaEventDisableTiming); cudaEventCreateWithFlags(&F->ev_curr[k], cudaEventDisableTiming); cudaEventCreateWithFlags(&F->ev_bank_free[k], cudaEventDisableTiming); cudaEventCreateWithFlags(&F->ev_so3[k], cudaEventDisableTiming); cudaEventCreateWithFlags(&F->ev_img[k], cudaEventDisableTiming); }
            cudaEventCreateWithFlags(&F->ev_up, cudaEventDisableTiming);
            cudaEventCreateWithFlags(&F->ev_track, cudaEventDisableTiming);
            int lo = 0, hi = 0;
            cudaDeviceGetStreamPriorityRange(&lo, &hi);
            if (cudaStreamCreateWithPriority(&F->pre_stream, cudaStreamNonBlocking, lo) != cudaSuccess ||
                cudaStreamCreateWithPriority(&F->so3_stream, cudaStreamNonBlocking, lo) != cudaSuccess) { set_error("fusion_create: stream creation failed"); rc = HRBF_ERR_CUDA; }
        }
    }
    if (rc) { hrbf_fusion_destroy(F); return rc; }
    *out = F;
    return HRBF_OK;
}
int hrbf_fusion_destroy(hrbf_fusion* F)
{
    if (!F) return HRBF_OK;
    hrbf_odometry_destroy(F->odom); hrbf_model_destroy(F->model); hrbf_indexmap_destroy(F->im); hrbf_fillin_destroy(F->fill); hrbf_frame_destroy(F->frames[0]); hrbf_frame_destroy(F->frames[1]);
    if (F->pre_stream) cudaStreamDestroy(F->pre_stream);
    if (F->so3_stream) cudaStreamDestroy(F->so3_stream);
    if (F->ev_up) cudaEventDestroy(F->ev_up);
    if (F->ev_track) cudaEventDestroy(F->ev_track);
    for (int k = 0; k < 2; ++k) { if (F->ev_staged[k]) cudaEventDestroy(F->ev_staged[k]); if (F->ev_free[k]) cudaEventDestroy(F->ev_free[k]); if (F->ev_curr[k]) cudaEventDestroy(F->ev_curr[k]); if (F->ev_bank_free[k]) cudaEventDestroy(F->ev_bank_free[k]); if (F->ev_so3[k]) cudaEventDestroy(F->ev_so3[k]); if (F->ev_img[k]) cudaEventDestroy(F->ev_img[k]); }
    if (F->dev) cudaFree(F->dev);
    if (F->traj) cudaFree(F->traj);
    if (F->h_pose) cudaFreeHost(F->h_pose);
    for (auto& e : F->ev) if (e) cudaEventDestroy(e);
    delete F;
    return HRBF_OK;
}
int hrbf_fusion_get_pose(hrbf_fusion* F, float* pose16, void* stream)
{
    HRBF_CHECK_ARG(F && pose16);
    cudaStream_t s = (cudaStream_t)stream;
    HRBF_CUDA(cudaMemcpyAsync(F->h_pose, F->dev, 12 * sizeof(float), cudaMemcpyDeviceToHost, s));
    HRBF_CUDA(cudaStreamSynchronize(s));
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) pose16[i * 4 + j] = F->h_pose[i * 3 + j]; pose16[i * 4 + 3] = F->h_pose[9 + i]; }
    pose16[12] = pose16[13] = pose16[14] = 0.f; pose16[15] = 1.f;
    return HRBF_OK;
}
int hrbf_fusion_process_frame_dev(hrbf_fusion* F, const unsigned char* rgb8, const unsigned short* depth16, long long timestamp, float weightMultiplier, void* stream)
{
    (void)timestamp;
    HRBF_CHECK_ARG(F && rgb8 && depth16);
    if (F->staged[F->cur]) { set_error("process_frame_dev: a staged frame is pending, call hrbf_fusion_process_staged"); return HRBF_ERR_INVALID_ARG; }
    if (int rc = frame_upload(F->frames[F->cur], rgb8, depth16, 0, F->tick == 1, (cudaStream_t)stream)) return rc;
    return fusion_frame(F, weightMultiplier, (cudaStream_t)stream);
}
int hrbf_fusion_process_frame(hrbf_fusion* F, const unsigned char* rgb8, const unsigned short* depth16, long long timestamp, float weightMultiplier,
                              float* pose16_out, void* stream)
{
    (void)timestamp;
    HRBF_CHECK_ARG(F && rgb8 && depth16);
    if (F->staged[F->cur]) { set_error("process_frame: a staged frame is pending, call hrbf_fusion_process_staged"); return HRBF_ERR_INVALID_ARG; }
    if (int rc = frame_upload(F->frames[F->cur], rgb8, depth16, 1, F->tick == 1, (cudaStream_t)stream)) return rc;
    if (int rc = fusion_frame(F, weightMultiplier, (cudaStream_t)stream)) return rc;
    float tmp[16];
    if (int rc = hrbf_fusion_get_pose(F, pose16_out ? pose16_out : tmp, stream)) return rc;
    if (F->timings) for (int k = 0; k < 4; ++k) cudaEventElapsedTime(&F->last_ms[k], F->ev[k], F->ev[k + 1]);
    return HRBF_OK;
}
// Pipelined form of processFrame for log replay: stage_frame uploads and preprocesses the NEXT frame(s) on an internal
// low-priority stream while the main stream still tracks / fuses the current one; process_staged consumes the oldest staged frame.
int hrbf_fusion_stage_frame(hrbf_fusion* F, const unsigned char* rgb8, const unsigned short* depth16, int host, void* stream)
{
    (void)stream;
    HRBF_CHECK_ARG(F && rgb8 && depth16);
    int b;
    if (!F->staged[F->cur]) b = F->cur;
    else if (!F->staged[F->cur ^ 1]) b = F->cur ^ 1;
    else { set_error("stage_frame: two frames are already staged"); return HRBF_ERR_CAPACITY; }
    // the first frame ever builds the RGBA texture (initFirstRGB); staged frame index = tick (+1 if it is the one after next)
    const bool first = F->tick + (b != F->cur ? 1 : 0) == 1;
    if (F->ev_free_valid[b]) HRBF_CUDA(cudaStreamWaitEvent(F->pre_stream, F->ev_free[b], 0));
    if (F->ev_track_valid) HRBF_CUDA(cudaStreamWaitEvent(F->pre_stream, F->ev_track, 0));
    if (int rc = frame_upload(F->frames[b], rgb8, depth16, host, first, F->pre_stream)) return rc;
    const int frame_number = F->tick + (b != F->cur ? 1 : 0);
    if (int rc = stage_all(F, b, frame_number, F->pre_stream)) return rc;
    HRBF_CUDA(cudaEventRecord(F->ev_staged[b], F->pre_stream));
    F->staged[b] = true;
    return HRBF_OK;
}
int hrbf_fusion_process_staged(hrbf_fusion* F, long long timestamp, float weightMultiplier, float* pose16_out, void* stream)
{
    (void)timestamp;
    HRBF_CHECK_ARG(F);
    if (!F->staged[F->cur]) { set_error("process_staged: no staged frame"); return HRBF_ERR_INVALID_ARG; }
    cudaStream_t s = (cudaStream_t)stream;
    // (the staged frame is waited for inside fusion_frame, where it is first read: after the model-side pyramids)
    if (int rc = fusion_frame(F, weightMultiplier, s, true)) return rc;
    F->staged[F->cur] = false;
    F->cur ^= 1;
    if (pose16_out) {
        if (int rc = hrbf_fusion_get_pose(F, pose16_out, stream)) return rc;
        if (F->timings) for (int k = 0; k < 4; ++k) cudaEventElapsedTime(&F->last_ms[k], F->ev[k], F->ev[k + 1]);
    }
    return HRBF_OK;
}
int hrbf_fusion_tick(const hrbf_fusion* F) { return F ? F->tick : 0; }
const float* hrbf_fusion_trajectory_dev(hrbf_fusion* F, int* n)
{
    if (!F) return nullptr;
    if (n) *n = F->traj_n < F->traj_cap ? F->traj_n : F->traj_cap;
    return F->traj;
}
hrbf_odometry* hrbf_fusion_odometry(hrbf_fusion* F) { return F ? F->odom : nullptr; }
hrbf_indexmap* hrbf_fusion_indexmap(hrbf_fusion* F) { return F ? F->im : nullptr; }
hrbf_model* hrbf_fusion_model(hrbf_fusion* F) { return F ? F->model : nullptr; }
hrbf_frame* hrbf_fusion_frame(hrbf_fusion* F) { return F ? F->frame : nullptr; }
hrbf_fillin* hrbf_fusion_fillin(hrbf_fusion* F) { return F ? F->fill : nullptr; }
int hrbf_fusion_enable_timings(hrbf_fusion* F, int on) { HRBF_CHECK_ARG(F); F->timings = on != 0; return HRBF_OK; }
int hrbf_fusion_last_timings(hrbf_fusion* F, float ms4[4])
{
    HRBF_CHECK_ARG(F && ms4);
    if (F->timings && cudaEventQuery(F->ev[4]) == cudaSuccess)       // the last enqueued frame has finished: its spans (else the last ones read)
        for (int k = 0; k < 4; ++k) cudaEventElapsedTime(&F->last_ms[k], F->ev[k], F->ev[k + 1]);
    (void)cudaGetLastError();
    for (int k = 0; k < 4; ++k) ms4[k] = F->last_ms[k];
    return HRBF_OK;
}

}  // extern "C"
