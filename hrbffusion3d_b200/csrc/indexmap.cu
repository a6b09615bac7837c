// indexmap.cu -- host side of SURVEY.md section 8 rows 6-7: GL-free restatement of the reference's
// IndexMap class (Core/src/IndexMap.{h,cpp}) behind the C ABI of include/hrbf_b200.h.
// Every GPUTexture of the reference is a dense device buffer here (RGBA32F -> float4[h][w]).
#include "indexmap_kernels.cuh"
#include <new>

using namespace hrbf;

#include "hrbf_internal.h"

static size_t tex_bytes(int which, size_t P)
{
    switch (which) {
    case HRBF_TEX_INDEX: return P * 4;
    case HRBF_TEX_IMAGE_HRBF: case HRBF_TEX_OLD_IMAGE_HRBF: return P * 4;
    case HRBF_TEX_TIME_HRBF: case HRBF_TEX_OLD_TIME_HRBF: return P * 2;
    case HRBF_TEX_ICPW_HRBF: case HRBF_TEX_OLD_ICPW_HRBF: return P * 4;
    default: return P * 16;
    }
}

extern "C" {

int hrbf_indexmap_create(hrbf_indexmap** out, int width, int height, float cx, float cy, float fx, float fy)
{
    HRBF_CHECK_ARG(out && width > 0 && height > 0);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device"); return HRBF_ERR_NO_DEVICE; }
    hrbf_indexmap* m = new (std::nothrow) hrbf_indexmap();
    HRBF_CHECK_ARG(m != nullptr);
    m->width = width; m->height = height; m->cx = cx; m->cy = cy; m->fx = fx; m->fy = fy;
    const size_t P = (size_t)width * height;
    size_t off = 0, o_tex[HRBF_TEX_COUNT];
    auto take = [&](size_t bytes) { size_t r = off; off += (bytes + 255) & ~(size_t)255; return r; };
    for (int t = 0; t < HRBF_TEX_COUNT; ++t) o_tex[t] = take(tex_bytes(t, P));
    const size_t o_keys = take(P * 8), o_kf = take(HRBF_ACTIVE_KEYFRAME_DIMENSION * 4), o_pose = take(8 * 12 * 4), o_cnt = take(8 * 4), o_lut = take(7 * 128 * 8);
    if (cudaMalloc(&m->slab, off) != cudaSuccess) { set_error("cudaMalloc(%zu) failed", off); delete m; return HRBF_ERR_CUDA; }
    cudaMemset(m->slab, 0, off);
    for (int t = 0; t < HRBF_TEX_COUNT; ++t) m->tex[t] = m->slab + o_tex[t];
    m->keys = (unsigned long long*)(m->slab + o_keys);
    m->active_kf = (float*)(m->slab + o_kf);
    m->inv_pose = (float*)(m->slab + o_pose);
    m->count_slot = (unsigned int*)(m->slab + o_cnt);
    m->row_lut = (unsigned long long*)(m->slab + o_lut);
    {
        unsigned long long lut[7 * 128];
        make_pred_row_lut(lut);
        cudaMemcpy(m->row_lut, lut, sizeof lut, cudaMemcpyHostToDevice);
    }
    cudaMallocHost(&m->h_stage, 8 * 16 * sizeof(float));
    cudaMallocHost(&m->h_kf, HRBF_ACTIVE_KEYFRAME_DIMENSION * sizeof(float));
    fill_keys_kernel<<<div_up((int)P, 256), 256>>>(m->keys, (int)P);
    const PredTable t = make_pred_table();
    cudaMemcpyToSymbol(c_pred, &t, sizeof t);
    // sub-map 0 active by default (HRBFFusion.cpp pushes key-frame 0 at start-up)
    memset(m->h_kf, 0, HRBF_ACTIVE_KEYFRAME_DIMENSION * sizeof(float));
    m->h_kf[0] = 1.0f;
    cudaMemcpy(m->active_kf, m->h_kf, HRBF_ACTIVE_KEYFRAME_DIMENSION * sizeof(float), cudaMemcpyHostToDevice);
    if (cudaDeviceSynchronize() != cudaSuccess || cudaGetLastError() != cudaSuccess) { set_error("indexmap_create: CUDA setup failed"); hrbf_indexmap_destroy(m); return HRBF_ERR_CUDA; }
    *out = m;
    return HRBF_OK;
}

int hrbf_indexmap_destroy(hrbf_indexmap* m)
{
    if (!m) return HRBF_OK;
    if (m->h_stage) cudaFreeHost(m->h_stage);
    if (m->h_kf) cudaFreeHost(m->h_kf);
    if (m->slab) cudaFree(m->slab);
    delete m;
    return HRBF_OK;
}

int hrbf_indexmap_set_active_keyframes(hrbf_indexmap* m, const int* ids, int n, void* stream)
{
    HRBF_CHECK_ARG(m && (ids || n == 0) && n >= 0);
    cudaStream_t s = (cudaStream_t)stream;
    HRBF_CUDA(cudaStreamSynchronize(s));     // the pinned mask may still be in flight
    memset(m->h_kf, 0, HRBF_ACTIVE_KEYFRAME_DIMENSION * sizeof(float));
    for (int i = 0; i < n; ++i) {
        HRBF_CHECK_ARG(ids[i] >= 0 && ids[i] < HRBF_ACTIVE_KEYFRAME_DIMENSION);
        m->h_kf[ids[i]] = 1.0f;
    }
    HRBF_CUDA(cudaMemcpyAsync(m->active_kf, m->h_kf, HRBF_ACTIVE_KEYFRAME_DIMENSION * sizeof(float), cudaMemcpyHostToDevice, s));
    return HRBF_OK;
}

using hrbf::indexmap_splat;

int hrbf_indexmap_predict_indices(hrbf_indexmap* m, const float* pose16, int time, int maxTime, const float* surfels_dev,
                                  unsigned int count, float depthCutoff, int insertSubmap, int indexSubmap, void* stream)
{
    (void)time; (void)maxTime; (void)insertSubmap; (void)indexSubmap;     // uniforms the shader declares but never reads
    HRBF_CHECK_ARG(m && pose16 && (surfels_dev || count == 0));
    cudaStream_t s = (cudaStream_t)stream;
    const int k = m->ring.acquire();
    float* h = m->h_stage + 16 * k;
    // pose.inverse() of a rigid transform: R^T, -R^T t (float)
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) h[i * 3 + j] = pose16[j * 4 + i];
    for (int i = 0; i < 3; ++i) h[9 + i] = -(h[i * 3] * pose16[3] + h[i * 3 + 1] * pose16[7] + h[i * 3 + 2] * pose16[11]);
    memcpy(h + 12, &count, 4);
    HRBF_CUDA(cudaMemcpyAsync(m->inv_pose + 12 * k, h, 12 * sizeof(float), cudaMemcpyHostToDevice, s));
    HRBF_CUDA(cudaMemcpyAsync(m->count_slot + k, h + 12, 4, cudaMemcpyHostToDevice, s));
    const int rc = indexmap_splat(m, m->inv_pose + 12 * k, surfels_dev, m->count_slot + k, count, depthCutoff, s);
    m->ring.release(k, s);      // after the kernels that read the slot's device twin
    return rc;
}

int hrbf_indexmap_predict_indices_dev(hrbf_indexmap* m, const float* inv_pose_dev, const float* surfels_dev, const unsigned int* count_dev,
                                      unsigned int count_bound, float depthCutoff, void* stream)
{
    HRBF_CHECK_ARG(m && inv_pose_dev && surfels_dev && count_dev);
    return indexmap_splat(m, inv_pose_dev, surfels_dev, count_dev, count_bound, depthCutoff, (cudaStream_t)stream);
}

int hrbf_indexmap_predict_hrbf(hrbf_indexmap* m, int predictionType, int win, int minNeighbors, int maxNeighbors, float confThreshold,
                               float icpWeightLambda, void* stream)
{
    HRBF_CHECK_ARG(m && (predictionType == 0 || predictionType == 1));
    HRBF_CHECK_ARG(win >= 0 && win <= 3 && maxNeighbors >= 0 && maxNeighbors <= 16 && minNeighbors >= 0);
    const int base = predictionType == 0 ? HRBF_TEX_IMAGE_HRBF : HRBF_TEX_OLD_IMAGE_HRBF;
    PredictArgs a;
    a.vertConf = (const float4*)m->tex[HRBF_TEX_VERTCONF]; a.colorTime = (const float4*)m->tex[HRBF_TEX_COLORTIME];
    a.normRad = (const float4*)m->tex[HRBF_TEX_NORMRAD]; a.curvMax = (const float4*)m->tex[HRBF_TEX_CURVMAX]; a.curvMin = (const float4*)m->tex[HRBF_TEX_CURVMIN];
    a.image = (uchar4*)m->tex[base + 0]; a.vertex = (float4*)m->tex[base + 1]; a.normal = (float4*)m->tex[base + 2];
    a.ocurvMax = (float4*)m->tex[base + 3]; a.ocurvMin = (float4*)m->tex[base + 4]; a.time = (unsigned short*)m->tex[base + 5]; a.icpw = (float*)m->tex[base + 6];
    a.cols = m->width; a.rows = m->height; a.cx = m->cx; a.cy = m->cy;
    a.icx = (float)(1.0 / (double)m->fx); a.icy = (float)(1.0 / (double)m->fy);      // IndexMap.cpp:449-452
    a.win = win; a.minN = minNeighbors; a.maxN = maxNeighbors; a.confThr = confThreshold; a.lambda = icpWeightLambda;
    a.dense_count = predictionType == 0 ? m->dense_count_next : nullptr;
    a.row_lut = m->row_lut;
    const dim3 grid(div_up(m->width, kPredTileW), div_up(m->height, kPredTileH));
    HRBF_LAUNCH_PDL(predict_hrbf_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, a);
    return HRBF_OK;
}

void* hrbf_indexmap_texture(hrbf_indexmap* m, int which)
{
    if (!m || which < 0 || which >= HRBF_TEX_COUNT) return nullptr;
    return m->tex[which];
}

}  // extern "C"

int hrbf::indexmap_splat(hrbf_indexmap* m, const float* inv_pose_dev, const float* surfels, const unsigned int* count_dev, unsigned int bound,
                         float depthCutoff, cudaStream_t s, int out_mask)
{
    SplatArgs a;
    a.inv_pose = inv_pose_dev;
    a.fx = m->fx; a.fy = m->fy; a.cx = m->cx; a.cy = m->cy; a.cols = m->width; a.rows = m->height; a.maxDepth = depthCutoff;
    a.active_kf = m->active_kf; a.kf_dim = HRBF_ACTIVE_KEYFRAME_DIMENSION;
    const int P = m->width * m->height;
    if (bound > 0) {
        int blocks = (int)((bound + 255u) / 256u);
        if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
        HRBF_LAUNCH_PDL(splat_keys_kernel, dim3(blocks), dim3(256), 0, s, (const float4*)surfels, count_dev, a, m->keys);
    }
    HRBF_LAUNCH_PDL(splat_gather_kernel, dim3(div_up(P, 256)), dim3(256), 0, s, (const float4*)surfels, a, m->keys, (unsigned int*)m->tex[HRBF_TEX_INDEX],
                                                      (float4*)m->tex[HRBF_TEX_VERTCONF], (float4*)m->tex[HRBF_TEX_COLORTIME], (float4*)m->tex[HRBF_TEX_NORMRAD],
                                                      (float4*)m->tex[HRBF_TEX_CURVMAX], (float4*)m->tex[HRBF_TEX_CURVMIN], out_mask);
    return HRBF_OK;
}
