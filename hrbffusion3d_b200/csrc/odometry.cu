// odometry.cu -- host side of the tracking path: GL-free restatement of the reference's
// RGBDOdometry class (Core/src/Utils/RGBDOdometry.{h,cpp}) and of the cudafuncs.cuh host
// functions, behind the C ABI of include/hrbf_b200.h.
//
// B200 design: all pyramid maps live in one dense HBM slab per object; each init* call is ONE
// fused kernel; getIncrementalTransformation replays ONE CUDA graph that contains the SO3
// pre-alignment, every ICP/RGB reduction of the three levels and the fp64 solves -- zero host
// round trips inside the Gauss-Newton loop (the reference makes ~67 per frame).
#include "track_persistent.cuh"
#include "icp_tile.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <map>
#include <new>
#include <stdarg.h>
#include <vector>

namespace hrbf {

std::atomic<unsigned long long> g_launches{0};
static thread_local char t_err[512] = "";
void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof t_err, fmt, ap);
    va_end(ap);
}

static inline dim3 grid2d(int cols, int rows, dim3 b) { return dim3(div_up(cols, b.x), div_up(rows, b.y)); }
static inline int reduce_blocks(int n) { int b = div_up(n, kReduceThreads * 2); return b < 1 ? 1 : (b > kMaxReduceBlocks ? kMaxReduceBlocks : b); }


}  // namespace hrbf

using namespace hrbf;

#include <mutex>
#include "hrbf_internal.h"

namespace hrbf {

static PyrOut pyr_out(hrbf_odometry* o, int which)
{
    PyrOut r;
    for (int l = 0; l < 3; ++l) {
        r.p[l] = o->maps[which][l]; r.pitch[l] = o->cols(l);
        const bool curr = which == M_VC || which == M_K1C, model = which == M_VG || which == M_K1G;
        r.pk0[l] = curr ? o->pk[0][l] : model ? o->pk[2][l] : nullptr;
        r.pk1[l] = curr ? o->pk[1][l] : model ? o->pk[3][l] : nullptr;
    }
    return r;
}

static int upload_pose(hrbf_odometry* o, const float* pose16, cudaStream_t s, float** dev_out)
{
    // row-major 4x4 -> R[9], t[3]; staged through a small pinned ring so the copy is truly async
    const int k = o->model_pose_ring.acquire();
    float* h = o->h_model_pose + 12 * k;
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) h[i * 3 + j] = pose16[i * 4 + j]; h[9 + i] = pose16[i * 4 + 3]; }
    HRBF_CUDA(cudaMemcpyAsync(o->pose_scratch, h, 12 * sizeof(float), cudaMemcpyHostToDevice, s));
    o->model_pose_ring.release(k, s);      // only the host slot is per-use here: the copy is its last reader
    *dev_out = o->pose_scratch;
    return HRBF_OK;
}

static IcpArgs icp_args(const hrbf_odometry* o, int l, bool use_weight)
{
    const int rows = o->rows(l), cols = o->cols(l), div = 1 << l;
    IcpArgs ia;
    ia.vc = o->maps[M_VC][l]; ia.nc = o->maps[M_NC][l]; ia.k1c = o->maps[M_K1C][l]; ia.k2c = o->maps[M_K2C][l]; ia.cpitch = cols;
    ia.vg = o->maps[M_VG][l]; ia.ng = o->maps[M_NG][l]; ia.k1g = o->maps[M_K1G][l]; ia.k2g = o->maps[M_K2G][l]; ia.gpitch = cols;
    ia.w = o->maps[M_W][l]; ia.wpitch = cols;
    ia.rows = rows; ia.cols = cols;
    ia.fx = o->intr.fx / div; ia.fy = o->intr.fy / div; ia.cx = o->intr.cx / div; ia.cy = o->intr.cy / div;
    ia.dist_thres = o->distThres; ia.angle_thres = o->angleThres;
    ia.use_search = o->useSearch; ia.radius = o->searchRadius; ia.use_weight = use_weight; ia.corres = nullptr;
    ia.pc0 = o->pk[0][l]; ia.pc1 = o->pk[1][l]; ia.pg0 = o->pk[2][l]; ia.pg1 = o->pk[3][l];
    return ia;
}
static RgbResArgs rgbres_args(const hrbf_odometry* o, int l)
{
    RgbResArgs ra;
    ra.minScale = (float)(pow((double)o->minGrad[l], 2.0) / pow((double)o->sobelScale, 2.0)); ra.maxDepthDelta = o->maxDepthDeltaRGB;
    ra.dIdx = o->dIdx[l]; ra.dIdy = o->dIdy[l]; ra.lastDepth = o->lastDepth[l]; ra.nextDepth = o->nextDepth[l];
    ra.lastImage = o->lastImage[l]; ra.nextImage = o->nextImage[l]; ra.corres = o->corresImg[l]; ra.rows = o->rows(l); ra.cols = o->cols(l);
    return ra;
}
static RgbStepArgs rgbstep_args(const hrbf_odometry* o, int l)
{
    const int div = 1 << l;
    RgbStepArgs sa;
    sa.corres = o->corresImg[l]; sa.cloud3 = o->cloud[l]; sa.dIdx = o->dIdx[l]; sa.dIdy = o->dIdy[l];
    sa.fx = o->intr.fx / div; sa.fy = o->intr.fy / div; sa.sobelScale = o->sobelScale; sa.use_grad_weight = o->rgbGradWeight;
    sa.rows = o->rows(l); sa.cols = o->cols(l);
    return sa;
}


// ---- tensor maps of the packed ICP records (icp_tile.cuh).  cuTensorMapEncodeTiled is fetched through the runtime, so the library
// does not link against libcuda.  geometry 0 = stand-alone tile kernel, 1 = persistent tracker (one tile per CTA)
static IcpTileGeom tile_geom_of(const hrbf_odometry* o, int which, int l)
{
    return which == 0 ? icp_tile_geom(o->rows(l), o->cols(l), o->num_sms) : track_tile_geom(o->rows(l), o->cols(l), o->num_sms);
}
static IcpTileMaps tile_maps_of(const hrbf_odometry* o, int which, int l)
{
    const CUtensorMap* m = (const CUtensorMap*)o->tmaps_host + (which * 3 + l) * 5;
    IcpTileMaps r;
    r.pc0 = m[0]; r.pc1 = m[1]; r.pg0 = m[2]; r.pg1 = m[3]; r.w = m[4];
    return r;
}
static int build_tile_tensor_maps(hrbf_odometry* o)
{
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
        (void)cudaGetLastError();
        return HRBF_OK;      // no TMA tensor maps on this driver: the gather kernels serve every path
    }
    auto encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    CUtensorMap* host = new (std::nothrow) CUtensorMap[2 * 3 * 5];
    if (!host) return HRBF_OK;
    auto fail = [&]() { delete[] host; return (int)HRBF_OK; };
    for (int which = 0; which < 2; ++which)
        for (int l = 0; l < 3; ++l) {
            const IcpTileGeom g = tile_geom_of(o, which, l);
            const cuuint64_t rows = (cuuint64_t)o->rows(l), cols = (cuuint64_t)o->cols(l);
            for (int k = 0; k < 5; ++k) {
                const bool weight = k == 4, curr = k < 2;
                void* base = weight ? (void*)o->maps[M_W][l] : (void*)o->pk[k][l];
                // 16-byte pixels as 2 x 64-bit elements (a box is at most 256 elements wide); the weight map as floats
                const cuuint64_t dims[2] = { weight ? cols : 2 * cols, rows };
                const cuuint64_t strides[1] = { cols * (weight ? sizeof(float) : sizeof(float4)) };
                const cuuint32_t bx = (cuuint32_t)(curr ? g.cbx : g.mbx), by = (cuuint32_t)(curr ? g.th : g.mh);
                const cuuint32_t box[2] = { weight ? bx : 2 * bx, by };
                const cuuint32_t estr[2] = { 1, 1 };
                if (box[0] > 256 || box[1] > 256 || strides[0] % 16 != 0) return fail();
                const CUresult r = encode(&host[(which * 3 + l) * 5 + k], weight ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, base, dims, strides,
                                          box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) return fail();
            }
        }
    o->tmaps_host = host;
    o->bank[o->cur_bank].tmaps = host;
    return HRBF_OK;
}

// The TMA-staged tile form of the ICP reduction (icp_tile.cuh) on the object's packed pyramids; pdl: programmatic stream serialization
static cudaError_t launch_icp_tile(hrbf_odometry* o, const IcpArgs& ia, int mode, int level, int next_level, cudaStream_t s, bool pdl)
{
    if (o->tmaps_host == nullptr || ia.cols % 4 != 0 || ia.pc0 == nullptr) {      // no tensor maps (driver), or weight-map rows not 16-byte aligned
        icp_reduce_kernel<false><<<reduce_blocks(ia.rows * ia.cols), kReduceThreads, 0, s>>>(ia, o->work, mode, level, next_level);
        return cudaGetLastError();
    }
    const IcpTileGeom g = tile_geom_of(o, 0, level);
    const IcpTileMaps maps = tile_maps_of(o, 0, level);
    const size_t dyn = icp_tile_smem_bytes(g);
    {
        static std::mutex mu;
        static size_t dyn_set = 0;
        std::lock_guard<std::mutex> lock(mu);
        if (dyn > dyn_set) {
            cudaError_t e = cudaFuncSetAttribute(icp_tile_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
            if (e != cudaSuccess) return e;
            dyn_set = dyn;
        }
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(g.ctas); cfg.blockDim = dim3(g.threads); cfg.dynamicSmemBytes = dyn; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, icp_tile_reduce_kernel, ia, g, maps, o->work, mode, level, next_level);
}

// Enqueue the whole tracking loop on `s` (used under stream capture).  Returns kernel count.
static int enqueue_track(hrbf_odometry* o, cudaStream_t s, bool rgbOnly, float icpWeight, bool pyramid, bool fastOdom,
                         bool so3, bool use_weight, bool host_io)
{
    const bool icp = !rgbOnly && icpWeight > 0;
    const bool rgb = rgbOnly || icpWeight < 100;
    int iters[3] = { fastOdom ? 3 : 10, pyramid ? 5 : 0, pyramid ? 4 : 0 };
    int first_level = 0;
    for (int l = 2; l >= 0; --l) if (iters[l] > 0) { first_level = l; break; }
    int n = 0;
    const dim3 b2(32, 8);
    ReduceWork* wk = o->work;

    if (host_io) cudaMemcpyAsync(o->pose_scratch + 12, o->h_pose, 12 * sizeof(float), cudaMemcpyHostToDevice, s);
    track_begin_kernel<<<1, 32, 0, s>>>(wk, o->pose_scratch + 12, icp, rgb, rgbOnly, so3, icpWeight, first_level); ++n;
    if (rgb)
        for (int l = 0; l < 3; ++l) { sobel_kernel<<<grid2d(o->cols(l), o->rows(l), b2), b2, 0, s>>>(o->rows(l), o->cols(l), o->nextImage[l], o->dIdx[l], o->dIdy[l]); ++n; }
    if (so3) {
        for (int it = 0; it < 10; ++it) {
            so3_reduce_kernel<<<reduce_blocks(o->rows(2) * o->cols(2)), kReduceThreads, 0, s>>>(o->lastNextImage[2], o->nextImage[2], o->rows(2), o->cols(2), wk, 1); ++n;
        }
        track_after_so3_kernel<<<1, 32, 0, s>>>(wk, first_level); ++n;
    }
    for (int l = 2; l >= 0; --l) {
        if (iters[l] == 0) continue;
        const int rows = o->rows(l), cols = o->cols(l), div = 1 << l;
        const float fx = o->intr.fx / div, fy = o->intr.fy / div, cx = o->intr.cx / div, cy = o->intr.cy / div;
        if (rgb) { project_cloud_kernel<<<grid2d(cols, rows, b2), b2, 0, s>>>(rows, cols, o->lastDepth[l], o->cloud[l], 1.0f / fx, 1.0f / fy, cx, cy); ++n; }
        int next_lower = -1;
        for (int q = l - 1; q >= 0; --q) if (iters[q] > 0) { next_lower = q; break; }
        const IcpArgs ia = icp_args(o, l, use_weight);
        const RgbResArgs ra = rgbres_args(o, l);
        const RgbStepArgs sa = rgbstep_args(o, l);
        const int nb = reduce_blocks(rows * cols);
        for (int j = 0; j < iters[l]; ++j) {
            const int next_level = (j + 1 < iters[l]) ? l : next_lower;
            if (rgb) { rgb_residual_kernel<<<nb, 256, 0, s>>>(ra, wk, 1, l, j == 0, next_lower); ++n; }
            if (icp) {
                if (o->useSearch) icp_reduce_kernel<true><<<nb, kReduceThreads, 0, s>>>(ia, wk, rgb ? 0 : 1, l, next_level);
                else icp_reduce_kernel<false><<<nb, kReduceThreads, 0, s>>>(ia, wk, rgb ? 0 : 1, l, next_level);
                ++n;
            }
            if (rgb) { rgb_step_kernel<<<nb, kReduceThreads, 0, s>>>(sa, -2.0f, wk, 1, l, next_level); ++n; }
        }
    }
    track_end_kernel<<<1, 32, 0, s>>>(wk, o->pose_scratch + 24); ++n;
    if (host_io) {
        cudaMemcpyAsync(o->h_pose + 12, o->pose_scratch + 24, 12 * sizeof(float), cudaMemcpyDeviceToHost, s);
        cudaMemcpyAsync(o->h_state, &wk->st, sizeof(TrackState), cudaMemcpyDeviceToHost, s);
    }
    return n;
}

// The whole tracking loop as one cooperative persistent kernel (track_persistent.cuh).
static int launch_track_persistent(hrbf_odometry* o, cudaStream_t s, bool rgbOnly, float icpWeight, bool pyramid, bool fastOdom,
                                   bool so3, bool use_weight, const float* prev_pose_dev, float* pose_out_dev, const int* iters_override = nullptr,
                                   const OdomFrameEpilogue* ep = nullptr)
{
    TrackParams p;
    int iters[3] = { fastOdom ? 3 : 10, pyramid ? 5 : 0, pyramid ? 4 : 0 };
    if (iters_override) for (int l = 0; l < 3; ++l) iters[l] = iters_override[l];
    for (int l = 0; l < 3; ++l) {
        p.lvl[l].icp = icp_args(o, l, use_weight);
        p.lvl[l].res = rgbres_args(o, l);
        p.lvl[l].step = rgbstep_args(o, l);
        p.lvl[l].cand = o->cand[l];
        p.lvl[l].iters = iters[l];
    }
    // RGB slots: one per pixel of a CTA's range and thread
    int max_slots = 1;
    for (int l = 0; l < 3; ++l)
        if (iters[l] > 0) { const int s_l = div_up(div_up(o->rows(l) * o->cols(l), o->num_sms), o->track_threads); if (s_l > max_slots) max_slots = s_l; }
    p.max_slots = max_slots;
    // shape 0: 512 threads (the whole register file: nothing co-resides), 1: 256 threads (two CTAs, or a CTA and other kernels, per SM),
    // 2: 384 threads (three quarters of the register file; the staging stream's kernels fit beside it)
    const int half = o->track_threads == 256 ? 1 : o->track_threads == 384 ? 2 : 0;
    const void* kernel = half == 1 ? (const void*)track_persistent_kernel<256> : half == 2 ? (const void*)track_persistent_kernel<384> : (const void*)track_persistent_kernel<512>;
    // resident ICP tiles (icp_tile.cuh): a level keeps its tile in shared memory when slots + tile fit beside the kernel's static
    // shared memory (the 256-thread shape shares the SM with another CTA: half the budget)
    size_t dyn = track_slots_bytes(max_slots, o->track_threads);
    {
        static std::mutex mu;
        static size_t budget[3] = { 0, 0, 0 };
        std::lock_guard<std::mutex> lock(mu);
        if (budget[half] == 0) {
            cudaFuncAttributes fa;
            HRBF_CUDA(cudaFuncGetAttributes(&fa, kernel));
            int dev = 0, optin = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
            size_t total = (size_t)optin;
            if (half == 1) total = total / 2 - 1024;      // two CTAs per SM, 1 KB reserved each
            if (half == 2) total = total * 3 / 4 - 1024;  // a quarter of the SM is left to co-resident kernels
            budget[half] = total > fa.sharedSizeBytes + 1024 ? total - fa.sharedSizeBytes - 512 : 1;
        }
        size_t tile_bytes = 0;
        for (int l = 0; l < 3; ++l) {
            p.tile[l] = tile_geom_of(o, 1, l);
            if (o->tmaps_host) p.tmaps[l] = tile_maps_of(o, 1, l);
            else memset(&p.tmaps[l], 0, sizeof p.tmaps[l]);
            const size_t need = icp_tile_smem_bytes(p.tile[l]);
            p.resident[l] = (iters[l] > 0 && o->tile_resident && o->tmaps_host && !o->useSearch && o->cols(l) % 4 == 0 && dyn + need <= budget[half]) ? 1 : 0;
            if (p.resident[l] && need > tile_bytes) tile_bytes = need;
        }
        dyn += tile_bytes;
    }
    {   // the attribute belongs to the kernel, not to this object: only ever raise it
        static std::mutex mu;
        static size_t dyn_set[3] = { 0, 0, 0 };
        std::lock_guard<std::mutex> lock(mu);
        if (dyn > dyn_set[half]) {
            HRBF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
            dyn_set[half] = dyn;
        }
    }
    p.so3_last = o->lastNextImage[2]; p.so3_next = o->nextImage[2];
    {   // a staged frame (odom_stage_current_dev) brings its Sobel images, candidate masks and SO3 pre-alignment along
        const CurrBank& cb = o->bank[o->cur_bank];
        p.cand_ready = (o->banked && cb.cand_ready) ? 1 : 0;
        p.so3_pre = (o->banked && so3 && cb.so3_ready) ? cb.so3 : nullptr;
    }
    p.icp = (!rgbOnly && icpWeight > 0) ? 1 : 0;
    p.rgb = (rgbOnly || icpWeight < 100) ? 1 : 0;
    p.rgbOnly = rgbOnly; p.so3 = so3; p.icpWeight = icpWeight;
    p.prev_pose = prev_pose_dev; p.pose_out = pose_out_dev;
    p.last_pose_out = ep ? ep->last_pose_out : nullptr; p.inv_pose_out = ep ? ep->inv_pose_out : nullptr;
    p.weighting_out = ep ? ep->weighting_out : nullptr; p.weight_multiplier = ep ? ep->weight_multiplier : 1.f;
    p.traj_out = ep ? ep->traj_out : nullptr;
    p.st_global = &o->work->st;
    p.ll_f = o->tp_ll_f; p.ll_i = o->tp_ll_i;
    o->tp_epoch = (o->tp_epoch + 1) & 0xfffffu;
    if (o->tp_epoch == 0) o->tp_epoch = 1;
    p.epoch = o->tp_epoch;
    p.dbg = o->tp_dbg;
    {   // cooperative (all CTAs co-resident: they spin on each other's words) + programmatic stream serialization (see pdl_wait)
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(o->num_sms); cfg.blockDim = dim3(o->track_threads); cfg.dynamicSmemBytes = dyn; cfg.stream = s;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeCooperative; attr[0].val.cooperative = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = o->tp_no_pdl ? 1 : 2;
        void* kargs[1] = { (void*)&p };
        cudaError_t e = cudaLaunchKernelExC(&cfg, kernel, kargs);
        if (e != cudaSuccess && !o->tp_no_pdl) {      // a driver that refuses the combination: cooperative only, from now on
            (void)cudaGetLastError();
            o->tp_no_pdl = true;
            cfg.numAttrs = 1;
            e = cudaLaunchKernelExC(&cfg, kernel, kargs);
        }
        HRBF_CUDA(e);
    }
    count_launch();
    return HRBF_OK;
}

// bring the packed records up to date after a builder that only wrote the SoA maps
static int repack_if_dirty(hrbf_odometry* o, cudaStream_t s)
{
    for (int side = 0; side < 2; ++side) {
        bool& dirty = side == 0 ? o->pack_dirty_curr : o->pack_dirty_model;
        if (!dirty) continue;
        for (int l = 0; l < 3; ++l) {
            const int n = o->rows(l) * o->cols(l);
            if (side == 0) pack_maps_kernel<<<div_up(n, 256), 256, 0, s>>>(n, o->maps[M_VC][l], o->maps[M_NC][l], o->maps[M_K1C][l], o->maps[M_K2C][l], o->pk[0][l], o->pk[1][l]);
            else pack_maps_kernel<<<div_up(n, 256), 256, 0, s>>>(n, o->maps[M_VG][l], o->maps[M_NG][l], o->maps[M_K1G][l], o->maps[M_K2G][l], o->pk[2][l], o->pk[3][l]);
            HRBF_KERNEL_CHECK();
        }
        dirty = false;
    }
    return HRBF_OK;
}

static int get_graph(hrbf_odometry* o, bool rgbOnly, float icpWeight, bool pyramid, bool fastOdom, bool so3, bool use_weight,
                     bool host_io, cudaGraphExec_t* exec, int* nk)
{
    uint32_t wbits;
    memcpy(&wbits, &icpWeight, 4);
    const uint64_t key = ((uint64_t)wbits << 32) | (uint64_t)(rgbOnly | pyramid << 1 | fastOdom << 2 | so3 << 3 | use_weight << 4 | host_io << 5 |
                                                             (o->useSearch ? 1 : 0) << 6 | (o->rgbGradWeight ? 1 : 0) << 7 | (uint64_t)(o->searchRadius & 0xff) << 8 | (uint64_t)(o->so3_parity & 1) << 16 | (uint64_t)(o->cur_bank & 1) << 17);
    auto it = o->graphs.find(key);
    if (it == o->graphs.end()) {
        cudaGraph_t g = nullptr;
        HRBF_CUDA(cudaStreamBeginCapture(o->cap_stream, cudaStreamCaptureModeThreadLocal));
        (void)cudaGetLastError();
        const int n = enqueue_track(o, o->cap_stream, rgbOnly, icpWeight, pyramid, fastOdom, so3, use_weight, host_io);
        // a launch that failed during capture (bad configuration, ...) must not leave a half-built graph in the cache
        const cudaError_t launch_err = cudaGetLastError();
        const cudaError_t end_err = cudaStreamEndCapture(o->cap_stream, &g);
        if (launch_err != cudaSuccess || end_err != cudaSuccess || g == nullptr) {
            if (g) cudaGraphDestroy(g);
            set_error("capture of the tracking graph failed: %s", cudaGetErrorString(launch_err != cudaSuccess ? launch_err : end_err));
            (void)cudaGetLastError();
            return HRBF_ERR_CUDA;
        }
        cudaGraphExec_t e = nullptr;
        const cudaError_t inst_err = cudaGraphInstantiate(&e, g, 0);
        cudaGraphDestroy(g);
        if (inst_err != cudaSuccess) { set_error("cudaGraphInstantiate -> %s", cudaGetErrorString(inst_err)); return HRBF_ERR_CUDA; }
        it = o->graphs.emplace(key, std::make_pair(e, n)).first;
    }
    *exec = it->second.first;
    *nk = it->second.second;
    return HRBF_OK;
}

}  // namespace hrbf

extern "C" {

const char* hrbf_last_error(void) { return t_err; }
const char* hrbf_version(void) { return "hrbf_b200 0.1 (sm_100a)"; }
unsigned long long hrbf_launch_count(void) { return g_launches.load(); }
size_t hrbf_reduce_workspace_bytes(void) { return sizeof(ReduceWork); }
int hrbf_copy_device(void* dst, const void* src, size_t bytes, void* stream)
{
    HRBF_CHECK_ARG(dst && src);
    HRBF_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return HRBF_OK;
}

// ------------------------------------------------------------------ row 5 ---
#define STEP_EL(step) ((int)((step) / sizeof(float)))

int hrbf_copy_maps(const float* v, const float* n, float* vmap, size_t vstep, float* nmap, size_t nstep, int rows, int cols, void* stream)
{
    HRBF_CHECK_ARG(v && n && vmap && nmap && rows > 0 && cols > 0 && vstep >= cols * sizeof(float) && nstep >= cols * sizeof(float));
    const dim3 b(32, 8);
    copy_maps_kernel<<<grid2d(cols, rows, b), b, 0, (cudaStream_t)stream>>>(rows, cols, (const float4*)v, (const float4*)n, vmap, STEP_EL(vstep), nmap, STEP_EL(nstep));
    HRBF_KERNEL_CHECK();
    return HRBF_OK;
}
int hrbf_copy_curvature_map(const float* c, float* cmap, size_t cstep, int rows, int cols, float thr, void* stream)
{
    HRBF_CHECK_ARG(c && cmap && rows > 0 && cols > 0 && cstep >= cols * sizeof(float));
    const dim3 b(32, 8);
    copy_curv_kernel<<<grid2d(cols, rows, b), b, 0, (cudaStream_t)stream>>>(rows, cols, (const float4*)c, cmap, STEP_EL(cstep), thr);
    HRBF_KERNEL_CHECK();
    return HRBF_OK;
}
int hrbf_copy_icpweight_map(const float* w, float* dst, size_t wstep, int rows, int cols, void* stream)
{
    HRBF_CHECK_ARG(w && dst && rows > 0 && cols > 0 && wstep >= cols * sizeof(float));
    const dim3 b(32, 8);
    copy_weight_kernel<<<grid2d(cols, rows, b), b, 0, (cudaStream_t)stream>>>(rows, cols, w, dst, STEP_EL(wstep));
    HRBF_KERNEL_CHECK();
    return HRBF_OK;
}
}  // extern "C"
template <int MODE>
static int resize_any(const float* in, size_t is, float* out, size_t os, int in_rows, int in_cols, void* stream)
{
    HRBF_CHECK_ARG(in && out && in_rows > 1 && in_cols > 1);
    const int drows = in_rows / 2, dcols = in_cols / 2;
    HRBF_CHECK_ARG(is >= in_cols * sizeof(float) && os >= dcols * sizeof(float));
    const dim3 b(32, 8);
    resize_kernel<MODE><<<grid2d(dcols, drows, b), b, 0, (cudaStream_t)stream>>>(drows, dcols, in_rows, in, STEP_EL(is), out, STEP_EL(os));
    HRBF_KERNEL_CHECK();
    return HRBF_OK;
}
extern "C" {
int hrbf_resize_vmap(const float* in, size_t is, float* out, size_t os, int r, int c, void* s) { return resize_any<0>(in, is, out, os, r, c, s); }
int hrbf_resize_nmap(const float* in, size_t is, float* out, size_t os, int r, int c, void* s) { return resize_any<1>(in, is, out, os, r, c, s); }
int hrbf_resize_cmap(const float* in, size_t is, float* out, size_t os, int r, int c, void* s) { return resize_any<2>(in, is, out, os, r, c, s); }
int hrbf_resize_icpweight_map(const float* in, size_t is, float* out, size_t os, int r, int c, void* s) { return resize_any<3>(in, is, out, os, r, c, s); }

}  // extern "C"
template <int MODE>
static int transform_any(const float* as, size_t ass, const float* bs, size_t bss, const float* R, const float* t,
                         float* ad, size_t ads, float* bd, size_t bds, int rows, int cols, void* stream)
{
    HRBF_CHECK_ARG(as && bs && R && t && ad && bd && rows > 0 && cols > 0);
    Mat33 Rm; Vec3 tv;
    memcpy(Rm.m, R, sizeof Rm.m);
    tv.x = t[0]; tv.y = t[1]; tv.z = t[2];
    const dim3 b(32, 8);
    transform_kernel<MODE><<<grid2d(cols, rows, b), b, 0, (cudaStream_t)stream>>>(rows, cols, as, STEP_EL(ass), bs, STEP_EL(bss), Rm, tv, ad, STEP_EL(ads), bd, STEP_EL(bds));
    HRBF_KERNEL_CHECK();
    return HRBF_OK;
}
extern "C" {
int hrbf_transform_maps(const float* vs, size_t vss, const float* ns, size_t nss, const float* R, const float* t,
                        float* vd, size_t vds, float* nd, size_t nds, int rows, int cols, void* stream)
{ return transform_any<0>(vs, vss, ns, nss, R, t, vd, vds, nd, nds, rows, cols, stream); }
int hrbf_transform_curv_maps(const float* k1s, size_t k1ss, const float* k2s, size_t k2ss, const float* R, const float* t,
                             float* k1d, size_t k1ds, float* k2d, size_t k2ds, int rows, int cols, void* stream)
{ return transform_any<1>(k1s, k1ss, k2s, k2ss, R, t, k1d, k1ds, k2d, k2ds, rows, cols, stream); }

// --------------------------------------------------------------- rows 1-3 ---
static int init_work_state(ReduceWork* wk, cudaStream_t s, const TrackState& h)
{
    HRBF_CUDA(cudaMemcpyAsync(&wk->st, &h, sizeof(TrackState), cudaMemcpyHostToDevice, s));
    return HRBF_OK;
}
static void unpack_host_se3(const double* sums, float* A, float* b, float* residual)
{
    int shift = 0;
    for (int i = 0; i < 6; ++i)
        for (int j = i; j < 7; ++j) {
            const float value = (float)sums[shift++];
            if (j == 6) b[i] = value; else A[j * 6 + i] = A[i * 6 + j] = value;
        }
    if (residual) { residual[0] = (float)sums[27]; residual[1] = (float)sums[28]; }
}

int hrbf_icp_step(const float* Rcurr, const float* tcurr, const float* vc, const float* nc, const float* k1c, const float* k2c, size_t cstep,
                  const float* Rprev_inv, const float* tprev, hrbf_camera intr,
                  const float* vg, const float* ng, const float* k1g, const float* k2g, size_t gstep,
                  const float* w, size_t wstep, int rows, int cols, const hrbf_icp_options* opts,
                  int* corres, void* work, float* A, float* b, float* residual, double* sums29, void* stream)
{
    HRBF_CHECK_ARG(Rcurr && tcurr && vc && nc && k1c && k2c && Rprev_inv && tprev && vg && ng && k1g && k2g && opts && work && A && b && residual);
    HRBF_CHECK_ARG(rows > 0 && cols > 0 && cstep >= cols * sizeof(float) && gstep >= cols * sizeof(float));
    HRBF_CHECK_ARG(!opts->use_weight || (w && wstep >= cols * sizeof(float)));
    cudaStream_t s = (cudaStream_t)stream;
    ReduceWork* wk = (ReduceWork*)work;
    TrackState h;
    memset(&h, 0, sizeof h);
    memcpy(h.Rcurr, Rcurr, 36); memcpy(h.tcurr, tcurr, 12); memcpy(h.Rprev_inv, Rprev_inv, 36); memcpy(h.tprev, tprev, 12);
    h.done_level = -1; h.icp = 1;
    if (int rc = init_work_state(wk, s, h)) return rc;
    IcpArgs ia;
    ia.vc = vc; ia.nc = nc; ia.k1c = k1c; ia.k2c = k2c; ia.cpitch = STEP_EL(cstep);
    ia.vg = vg; ia.ng = ng; ia.k1g = k1g; ia.k2g = k2g; ia.gpitch = STEP_EL(gstep);
    ia.w = w; ia.wpitch = STEP_EL(wstep);
    ia.rows = rows; ia.cols = cols; ia.fx = intr.fx; ia.fy = intr.fy; ia.cx = intr.cx; ia.cy = intr.cy;
    ia.dist_thres = opts->dist_thres; ia.angle_thres = opts->angle_thres;
    ia.use_search = opts->use_search; ia.radius = opts->search_radius; ia.use_weight = opts->use_weight; ia.corres = (int2*)corres;
    ia.pc0 = ia.pc1 = ia.pg0 = ia.pg1 = nullptr;          // caller-owned SoA maps
    const int nb = reduce_blocks(rows * cols);
    if (opts->use_search) icp_reduce_kernel<true><<<nb, kReduceThreads, 0, s>>>(ia, wk, 0, 0, -1);
    else icp_reduce_kernel<false><<<nb, kReduceThreads, 0, s>>>(ia, wk, 0, 0, -1);
    HRBF_KERNEL_CHECK();
    double sums[32];
    HRBF_CUDA(cudaMemcpyAsync(sums, wk->st.icp_sums, sizeof sums, cudaMemcpyDeviceToHost, s));
    HRBF_CUDA(cudaStreamSynchronize(s));
    unpack_host_se3(sums, A, b, residual);
    if (sums29) memcpy(sums29, sums, 29 * sizeof(double));
    return HRBF_OK;
}

int hrbf_compute_rgb_residual(float minScale, const short* dIdx, const short* dIdy, const float* lastDepth, const float* nextDepth,
                              const unsigned char* lastImage, const unsigned char* nextImage, hrbf_dataterm* corresImg,
                              float maxDepthDelta, const float* kt, const float* krkinv, int rows, int cols, void* work,
                              int* sigmaSum, int* count, void* stream)
{
    HRBF_CHECK_ARG(dIdx && dIdy && lastDepth && nextDepth && lastImage && nextImage && corresImg && kt && krkinv && work && sigmaSum && count && rows > 0 && cols > 0);
    cudaStream_t s = (cudaStream_t)stream;
    ReduceWork* wk = (ReduceWork*)work;
    TrackState h;
    memset(&h, 0, sizeof h);
    memcpy(h.krkinv, krkinv, 36); memcpy(h.kt, kt, 12);
    h.done_level = -1;
    if (int rc = init_work_state(wk, s, h)) return rc;
    RgbResArgs ra;
    ra.minScale = minScale; ra.maxDepthDelta = maxDepthDelta; ra.dIdx = dIdx; ra.dIdy = dIdy; ra.lastDepth = lastDepth; ra.nextDepth = nextDepth;
    ra.lastImage = lastImage; ra.nextImage = nextImage; ra.corres = corresImg; ra.rows = rows; ra.cols = cols;
    rgb_residual_kernel<<<reduce_blocks(rows * cols), 256, 0, s>>>(ra, wk, 0, 0, 1, -1);
    HRBF_KERNEL_CHECK();
    int out[2];
    HRBF_CUDA(cudaMemcpyAsync(out, &wk->st.rgb_count, sizeof out, cudaMemcpyDeviceToHost, s));
    HRBF_CUDA(cudaStreamSynchronize(s));
    *count = out[0]; *sigmaSum = out[1];
    return HRBF_OK;
}

int hrbf_rgb_step(const hrbf_dataterm* corresImg, float sigma, const float* cloud3, float fx, float fy, const short* dIdx, const short* dIdy,
                  int use_gradient_weight, float sobelScale, int rows, int cols, void* work, float* A, float* b, double* sums29, void* stream)
{
    HRBF_CHECK_ARG(corresImg && cloud3 && dIdx && dIdy && work && A && b && rows > 0 && cols > 0);
    cudaStream_t s = (cudaStream_t)stream;
    ReduceWork* wk = (ReduceWork*)work;
    TrackState h;
    memset(&h, 0, sizeof h);
    h.done_level = -1; h.rgb = 1;
    if (int rc = init_work_state(wk, s, h)) return rc;
    RgbStepArgs sa;
    sa.corres = corresImg; sa.cloud3 = cloud3; sa.dIdx = dIdx; sa.dIdy = dIdy; sa.fx = fx; sa.fy = fy; sa.sobelScale = sobelScale;
    sa.use_grad_weight = use_gradient_weight; sa.rows = rows; sa.cols = cols;
    rgb_step_kernel<<<reduce_blocks(rows * cols), kReduceThreads, 0, s>>>(sa, sigma, wk, 0, 0, -1);
    HRBF_KERNEL_CHECK();
    double sums[32];
    HRBF_CUDA(cudaMemcpyAsync(sums, wk->st.rgb_sums, sizeof sums, cudaMemcpyDeviceToHost, s));
    HRBF_CUDA(cudaStreamSynchronize(s));
    unpack_host_se3(sums, A, b, nullptr);
    if (sums29) memcpy(sums29, sums, 29 * sizeof(double));
    return HRBF_OK;
}

int hrbf_so3_step(const unsigned char* lastImage, const unsigned char* nextImage, const float* imageBasis, const float* kinv, const float* krlr,
                  int rows, int cols, void* work, float* A, float* b, float* residual, double* sums11, void* stream)
{
    HRBF_CHECK_ARG(lastImage && nextImage && imageBasis && kinv && krlr && work && A && b && residual && rows > 2 && cols > 2);
    cudaStream_t s = (cudaStream_t)stream;
    ReduceWork* wk = (ReduceWork*)work;
    TrackState h;
    memset(&h, 0, sizeof h);
    memcpy(h.so3_basis, imageBasis, 36); memcpy(h.so3_kinv, kinv, 36); memcpy(h.so3_krlr, krlr, 36);
    h.done_level = -1;
    if (int rc = init_work_state(wk, s, h)) return rc;
    so3_reduce_kernel<<<reduce_blocks(rows * cols), kReduceThreads, 0, s>>>(lastImage, nextImage, rows, cols, wk, 0);
    HRBF_KERNEL_CHECK();
    double sums[16];
    HRBF_CUDA(cudaMemcpyAsync(sums, wk->st.so3_sums, sizeof sums, cudaMemcpyDeviceToHost, s));
    HRBF_CUDA(cudaStreamSynchronize(s));
    int shift = 0;
    for (int i = 0; i < 3; ++i)
        for (int j = i; j < 4; ++j) {
            const float value = (float)sums[shift++];
            if (j == 3) b[i] = value; else A[j * 3 + i] = A[i * 3 + j] = value;
        }
    residual[0] = (float)sums[9]; residual[1] = (float)sums[10];
    if (sums11) memcpy(sums11, sums, 11 * sizeof(double));
    return HRBF_OK;
}

// ------------------------------------------------------------------ row 4 ---
int hrbf_odometry_create(hrbf_odometry** out, int width, int height, float cx, float cy, float fx, float fy, float distThresh, float angleThresh)
{
    HRBF_CHECK_ARG(out && width >= 32 && height >= 16 && width % 8 == 0 && height % 4 == 0);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device"); return HRBF_ERR_NO_DEVICE; }
    hrbf_odometry* o = new (std::nothrow) hrbf_odometry();
    HRBF_CHECK_ARG(o != nullptr);
    o->width = width; o->height = height;
    o->intr.fx = fx; o->intr.fy = fy; o->intr.cx = cx; o->intr.cy = cy;
    o->distThres = distThresh; o->angleThres = angleThresh;
    o->sobelScale = (float)(1.0 / pow(2.0, 3));

    // one slab, 256-B aligned sub-buffers
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t r = off; off += (bytes + 255) & ~(size_t)255; return r; };
    size_t o_maps[M_COUNT][3], o_dt[3], o_ld[3], o_nd[3], o_li[3], o_ni[3], o_lni[3], o_dx[3], o_dy[3], o_cl[3], o_ci[3], o_cd[3], o_pk[4][3];
    for (int l = 0; l < 3; ++l) {
        const size_t P = (size_t)o->rows(l) * o->cols(l);
        for (int m = 0; m < M_W; ++m) o_maps[m][l] = take(4 * P * sizeof(float));
        o_maps[M_W][l] = take(P * sizeof(float));
        o_dt[l] = take(P * 4); o_ld[l] = take(P * 4); o_nd[l] = take(P * 4);
        o_li[l] = take(P); o_ni[l] = take(P); o_lni[l] = take(P);
        o_dx[l] = take(P * 2); o_dy[l] = take(P * 2); o_cl[l] = take(P * 12); o_ci[l] = take(P * sizeof(hrbf_dataterm)); o_cd[l] = take(P);
        for (int k = 0; k < 4; ++k) o_pk[k][l] = take(P * sizeof(float4));
    }
    const size_t o_vd = take((size_t)width * height * 4), o_work = take(sizeof(ReduceWork)), o_pose = take(64 * sizeof(float));
    cudaDeviceGetAttribute(&o->num_sms, cudaDevAttrMultiProcessorCount, 0);
    { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&o->num_sms, cudaDevAttrMultiProcessorCount, dev); }
    if (o->num_sms <= 0) o->num_sms = kNumSMs;
    const size_t o_tpp = take((size_t)2 * o->num_sms * 64 * sizeof(unsigned long long)), o_tpi = take((size_t)2 * o->num_sms * kIntStride * sizeof(unsigned long long));
    if (cudaMalloc(&o->slab, off) != cudaSuccess) { set_error("cudaMalloc(%zu) failed", off); delete o; return HRBF_ERR_CUDA; }
    cudaMemset(o->slab, 0, off);
    for (int l = 0; l < 3; ++l) {
        for (int m = 0; m < M_COUNT; ++m) o->maps[m][l] = (float*)(o->slab + o_maps[m][l]);
        o->depth_tmp[l] = (float*)(o->slab + o_dt[l]); o->lastDepth[l] = (float*)(o->slab + o_ld[l]); o->nextDepth[l] = (float*)(o->slab + o_nd[l]);
        o->lastImage[l] = (unsigned char*)(o->slab + o_li[l]); o->nextImage[l] = (unsigned char*)(o->slab + o_ni[l]); o->lastNextImage[l] = (unsigned char*)(o->slab + o_lni[l]);
        o->dIdx[l] = (short*)(o->slab + o_dx[l]); o->dIdy[l] = (short*)(o->slab + o_dy[l]);
        o->cloud[l] = (float*)(o->slab + o_cl[l]); o->corresImg[l] = (hrbf_dataterm*)(o->slab + o_ci[l]); o->cand[l] = (unsigned char*)(o->slab + o_cd[l]);
        for (int k = 0; k < 4; ++k) o->pk[k][l] = (float4*)(o->slab + o_pk[k][l]);
    }
    o->vdepth_tmp = (float*)(o->slab + o_vd);
    o->work = (ReduceWork*)(o->slab + o_work);
    o->pose_scratch = (float*)(o->slab + o_pose);
    for (int l = 0; l < 3; ++l) {
        CurrBank& b0 = o->bank[0];
        for (int m = 0; m < 4; ++m) b0.maps[m][l] = o->maps[M_VC + m][l];
        b0.pk[0][l] = o->pk[0][l]; b0.pk[1][l] = o->pk[1][l];
        b0.nextImage[l] = o->nextImage[l]; b0.nextDepth[l] = o->nextDepth[l]; b0.dIdx[l] = o->dIdx[l]; b0.dIdy[l] = o->dIdy[l]; b0.cand[l] = o->cand[l];
    }
    o->tp_ll_f = (unsigned long long*)(o->slab + o_tpp); o->tp_ll_i = (unsigned long long*)(o->slab + o_tpi);      // zeroed with the slab: tag 0 never matches
    if (int rc = build_tile_tensor_maps(o)) { hrbf_odometry_destroy(o); return rc; }
    cudaMallocHost(&o->h_pose, 24 * sizeof(float));
    cudaMallocHost(&o->h_state, sizeof(TrackState));
    cudaMallocHost(&o->h_model_pose, 8 * 12 * sizeof(float));
    cudaStreamCreateWithFlags(&o->cap_stream, cudaStreamNonBlocking);
    // camera lives in the device state for the fp64 K matrices
    TrackState h;
    memset(&h, 0, sizeof h);
    h.fx = fx; h.fy = fy; h.cx = cx; h.cy = cy; h.done_level = -1;
    cudaMemcpy(&o->work->st, &h, sizeof h, cudaMemcpyHostToDevice);
    if (cudaGetLastError() != cudaSuccess) { set_error("odometry_create: CUDA setup failed"); hrbf_odometry_destroy(o); return HRBF_ERR_CUDA; }
    *out = o;
    return HRBF_OK;
}

int hrbf_odometry_destroy(hrbf_odometry* o)
{
    if (!o) return HRBF_OK;
    for (auto& kv : o->graphs) cudaGraphExecDestroy(kv.second.first);
    if (o->cap_stream) cudaStreamDestroy(o->cap_stream);
    if (o->h_pose) cudaFreeHost(o->h_pose);
    if (o->h_state) cudaFreeHost(o->h_state);
    if (o->h_model_pose) cudaFreeHost(o->h_model_pose);
    o->model_pose_ring.destroy();
    if (o->slab) cudaFree(o->slab);
    if (o->bank1_slab) cudaFree(o->bank1_slab);
    delete[] (CUtensorMap*)o->bank[0].tmaps;
    delete[] (CUtensorMap*)o->bank[1].tmaps;
    if (o->tp_dbg) cudaFree(o->tp_dbg);
    delete o;
    return HRBF_OK;
}

int hrbf_odometry_debug_stamps(hrbf_odometry* o, long long* out_host, int n)
{   // development aid: (slot << 56 | globaltimer ns) stamps written by CTA 0 of the last persistent tracking call
    HRBF_CHECK_ARG(o && out_host && n > 0 && n <= 512);
    if (!o->tp_dbg) { HRBF_CUDA(cudaMalloc(&o->tp_dbg, 512 * sizeof(long long))); HRBF_CUDA(cudaMemset(o->tp_dbg, 0, 512 * sizeof(long long))); return HRBF_OK; }
    HRBF_CUDA(cudaMemcpy(out_host, o->tp_dbg, n * sizeof(long long), cudaMemcpyDeviceToHost));
    HRBF_CUDA(cudaMemset(o->tp_dbg, 0, 512 * sizeof(long long)));
    return HRBF_OK;
}
int hrbf_odometry_set_tracker(hrbf_odometry* o, int use_kernel_graph)
{
    HRBF_CHECK_ARG(o);
    o->use_graph = use_kernel_graph != 0;
    return HRBF_OK;
}
int hrbf_odometry_set_tracker_threads(hrbf_odometry* o, int threads)
{
    HRBF_CHECK_ARG(o);
    if (threads == 0) threads = kTrackThreadsDefault;
    if (threads != 256 && threads != 384 && threads != 512) { set_error("set_tracker_threads: 256, 384 or 512 (0 = default 512)"); return HRBF_ERR_INVALID_ARG; }
    o->track_threads = threads;
    return HRBF_OK;
}
int hrbf_odometry_set_tracker_tiles(hrbf_odometry* o, int resident)
{
    HRBF_CHECK_ARG(o);
    o->tile_resident = resident != 0;
    return HRBF_OK;
}
int hrbf_odometry_set_params(hrbf_odometry* o, float curvThr, int useSearch, int searchRadius, int rgbGradWeight)
{
    HRBF_CHECK_ARG(o && searchRadius >= 0 && searchRadius <= 2);
    o->curvThr = curvThr; o->useSearch = useSearch; o->searchRadius = searchRadius; o->rgbGradWeight = rgbGradWeight;
    return HRBF_OK;
}

int hrbf_odometry_init_icp_depth(hrbf_odometry* o, const float* depth, float cutoff, float factor, void* stream)
{
    HRBF_CHECK_ARG(o && depth);
    cudaStream_t s = (cudaStream_t)stream;
    const dim3 b(32, 8);
    HRBF_CUDA(cudaMemcpyAsync(o->depth_tmp[0], depth, (size_t)o->width * o->height * 4, cudaMemcpyDeviceToDevice, s));
    for (int i = 1; i < 3; ++i) {
        pyrdown_depth_kernel<<<grid2d(o->cols(i), o->rows(i), b), b, 0, s>>>(o->rows(i - 1), o->cols(i - 1), o->depth_tmp[i - 1], o->depth_tmp[i]);
        HRBF_KERNEL_CHECK();
    }
    for (int i = 0; i < 3; ++i) {
        const int div = 1 << i;
        create_vmap_kernel<<<grid2d(o->cols(i), o->rows(i), b), b, 0, s>>>(o->rows(i), o->cols(i), o->depth_tmp[i], o->maps[M_VC][i], o->cols(i),
                                                                         1.f / (o->intr.fx / div), 1.f / (o->intr.fy / div), o->intr.cx / div, o->intr.cy / div, cutoff, factor);
        HRBF_KERNEL_CHECK();
        create_nmap_kernel<<<grid2d(o->cols(i), o->rows(i), b), b, 0, s>>>(o->rows(i), o->cols(i), o->maps[M_VC][i], o->cols(i), o->maps[M_NC][i], o->cols(i));
        HRBF_KERNEL_CHECK();
    }
    o->pack_dirty_curr = true;
    return HRBF_OK;
}

static inline dim3 pyr_grid(const hrbf_odometry* o) { return dim3(div_up(o->width, 32), div_up(o->height, 8)); }

int hrbf_odometry_init_icp(hrbf_odometry* o, const float* v, const float* n, float depthCutoff, void* stream)
{
    (void)depthCutoff;
    HRBF_CHECK_ARG(o && v && n);
    pyr_pair_kernel<PYR_VN><<<pyr_grid(o), 256, 0, (cudaStream_t)stream>>>((const float4*)v, (const float4*)n, o->height, o->width, 0.f, nullptr,
                                                                         pyr_out(o, M_VC), pyr_out(o, M_NC), o->vdepth_tmp, o->maxDepthRGB, nullptr, nullptr, nullptr);
    HRBF_KERNEL_CHECK();
    return HRBF_OK;
}
}  // extern "C"
namespace hrbf {
// device-pose / device-select forms used by the fused frame pipeline (fusion.cu)
int odom_init_icp_model_dev(hrbf_odometry* o, const float* v, const float* n, const float* v_alt, const float* n_alt, const int* sel,
                            const float* pose_dev, cudaStream_t s)
{
    pyr_pair_kernel<PYR_VN><<<pyr_grid(o), 256, 0, s>>>((const float4*)v, (const float4*)n, o->height, o->width, 0.f, pose_dev,
                                                      pyr_out(o, M_VG), pyr_out(o, M_NG), o->vdepth_tmp, o->maxDepthRGB,
                                                      (const float4*)v_alt, (const float4*)n_alt, sel);
    HRBF_KERNEL_CHECK();
    return HRBF_OK;
}
int odom_init_curvature_model_dev(hrbf_odometry* o, const float* k1, const float* k2, const float* k1_alt, const float* k2_alt, const int* sel,
                                  const float* pose_dev, cudaStream_t s)
{
    pyr_pair_kernel<PYR_K><<<pyr_grid(o), 256, 0, s>>>((const float4*)k1, (const float4*)k2, o->height, o->width, o->curvThr, pose_dev,
                                                     pyr_out(o, M_K1G), pyr_out(o, M_K2G), nullptr, 0.f, (const float4*)k1_alt, (const float4*)k2_alt, sel);
    HRBF_KERNEL_CHECK();
    return HRBF_OK;
}
// Tracking call of the frame pipeline: pose_inout holds the previous pose on entry and the new one on exit; the epilogue
// outputs (HRBFFusion.cpp:1109-1123) are written by the tracking kernel itself.  Returns 1 if the configured tracker cannot do
// that (kernel-graph tracker): the caller then issues the separate calls.
int odom_track_frame_dev(hrbf_odometry* o, float* pose_inout, const OdomFrameEpilogue& ep, bool rgbOnly, float icpWeight, bool pyramid,
                         bool fastOdom, bool so3, bool use_weight, cudaStream_t s)
{
    if (o->use_graph) return 1;
    if (int rc = repack_if_dirty(o, s)) return rc;
    if (int rc = launch_track_persistent(o, s, rgbOnly, icpWeight, pyramid, fastOdom, so3, use_weight, pose_inout, pose_inout, nullptr, &ep)) return rc;
    if (so3 && !o->banked) {      // RGBDOdometry.cpp:1239-1245 (swap_so3_images); with banks the other bank IS the last camera image
        for (int i = 0; i < 3; ++i) std::swap(o->lastNextImage[i], o->nextImage[i]);
        o->so3_parity ^= 1;
    }
    return HRBF_OK;
}

// ---- current-frame banks (hrbf_internal.h: CurrBank) ----
int odom_enable_banks(hrbf_odometry* o)
{
    if (o->banked) return HRBF_OK;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t r = off; off += (bytes + 255) & ~(size_t)255; return r; };
    size_t o_maps[4][3], o_pk[2][3], o_ni[3], o_nd[3], o_dx[3], o_dy[3], o_cd[3];
    for (int l = 0; l < 3; ++l) {
        const size_t P = (size_t)o->rows(l) * o->cols(l);
        for (int m = 0; m < 4; ++m) o_maps[m][l] = take(4 * P * sizeof(float));
        for (int k = 0; k < 2; ++k) o_pk[k][l] = take(P * sizeof(float4));
        o_ni[l] = take(P); o_nd[l] = take(P * 4); o_dx[l] = take(P * 2); o_dy[l] = take(P * 2); o_cd[l] = take(P);
    }
    const size_t o_so3 = take(2 * sizeof(So3Pre));
    if (cudaMalloc(&o->bank1_slab, off) != cudaSuccess) { set_error("cudaMalloc(%zu) failed", off); return HRBF_ERR_CUDA; }
    HRBF_CUDA(cudaMemset(o->bank1_slab, 0, off));
    CurrBank& b1 = o->bank[1];
    for (int l = 0; l < 3; ++l) {
        for (int m = 0; m < 4; ++m) b1.maps[m][l] = (float*)(o->bank1_slab + o_maps[m][l]);
        for (int k = 0; k < 2; ++k) b1.pk[k][l] = (float4*)(o->bank1_slab + o_pk[k][l]);
        b1.nextImage[l] = (unsigned char*)(o->bank1_slab + o_ni[l]); b1.nextDepth[l] = (float*)(o->bank1_slab + o_nd[l]);
        b1.dIdx[l] = (short*)(o->bank1_slab + o_dx[l]); b1.dIdy[l] = (short*)(o->bank1_slab + o_dy[l]); b1.cand[l] = (unsigned char*)(o->bank1_slab + o_cd[l]);
    }
    o->bank[0].so3 = (So3Pre*)(o->bank1_slab + o_so3); b1.so3 = o->bank[0].so3 + 1;
    o->banked = true;
    if (o->tmaps_host) {      // tensor maps over bank 1's records
        odom_select_bank(o, 1);
        o->tmaps_host = nullptr;
        const int rc = build_tile_tensor_maps(o);
        if (rc || o->tmaps_host == nullptr) {      // all or nothing: without maps for both banks the gather kernels serve every path
            delete[] (CUtensorMap*)o->bank[0].tmaps; o->bank[0].tmaps = nullptr; b1.tmaps = nullptr;
        }
        odom_select_bank(o, 0);
    }
    return HRBF_OK;
}
void odom_select_bank(hrbf_odometry* o, int b)
{
    const CurrBank& cb = o->bank[b];
    for (int l = 0; l < 3; ++l) {
        for (int m = 0; m < 4; ++m) o->maps[M_VC + m][l] = cb.maps[m][l];
        o->pk[0][l] = cb.pk[0][l]; o->pk[1][l] = cb.pk[1][l];
        o->nextImage[l] = cb.nextImage[l]; o->nextDepth[l] = cb.nextDepth[l];
        o->dIdx[l] = cb.dIdx[l]; o->dIdy[l] = cb.dIdy[l]; o->cand[l] = cb.cand[l];
        if (o->banked) o->lastNextImage[l] = o->bank[b ^ 1].nextImage[l];
    }
    o->tmaps_host = cb.tmaps;
    o->cur_bank = b;
}
static int launch_prep_all(hrbf_odometry* o, const OdomPrepInputs& in, const int* jobs, int njobs, cudaStream_t s);
int odom_stage_so3_dev(hrbf_odometry* o, int b, const unsigned char* rgb8, bool so3, bool has_previous, cudaEvent_t image_done, cudaStream_t s)
{
    if (!o->banked) { set_error("odom_stage_so3_dev: banks not enabled"); return HRBF_ERR_INVALID_ARG; }
    CurrBank& cb = o->bank[b];
    cb.so3_ready = false;
    cb.image_ready = false;
    if (so3) {
        RgbdJob j;
        memset(&j, 0, sizeof j);
        j.rgb8 = rgb8;
        for (int l = 0; l < 3; ++l) j.img[l] = cb.nextImage[l];
        HRBF_LAUNCH_PDL(so3_image_kernel, dim3(div_up(o->cols(2), 8), div_up(o->rows(2), 8)), dim3(256), 0, s, j, o->height, o->width);
        cb.image_ready = true;
    }
    if (image_done) HRBF_CUDA(cudaEventRecord(image_done, s));
    if (so3 && has_previous) {
        HRBF_LAUNCH_PDL(so3_prealign_kernel, dim3(1), dim3(kSo3Threads), 0, s, (const unsigned char*)o->bank[b ^ 1].nextImage[2], (const unsigned char*)cb.nextImage[2],
                        o->rows(2), o->cols(2), o->intr.fx, o->intr.fy, o->intr.cx, o->intr.cy, cb.so3);
        cb.so3_ready = true;
    }
    return HRBF_OK;
}
int odom_stage_current_dev(hrbf_odometry* o, int b, const OdomPrepInputs& in, cudaStream_t s)
{
    if (!o->banked) { set_error("odom_stage_current_dev: banks not enabled"); return HRBF_ERR_INVALID_ARG; }
    const int before = o->cur_bank;
    odom_select_bank(o, b);          // the argument builders read the active pointers; kernel arguments are captured at launch
    static const int jobs[3] = { 1, 3, 5 };      // RGB-D pyramids of the camera frame, its vertex / normal maps, its curvature maps
    if (int rc = launch_prep_all(o, in, jobs, 3, s)) { odom_select_bank(o, before); return rc; }
    SobelCandArgs sc;
    for (int l = 0; l < 3; ++l) { sc.r[l] = rgbres_args(o, l); sc.cand[l] = o->cand[l]; }
    HRBF_LAUNCH_PDL(sobel_cand_kernel, dim3(div_up(o->width, 32), div_up(o->height, 8), 3), dim3(256), 0, s, sc);
    o->bank[b].cand_ready = true;
    o->pack_dirty_curr = false;
    odom_select_bank(o, before);
    return HRBF_OK;
}

int odom_prep_all_dev(hrbf_odometry* o, const OdomPrepInputs& in, cudaStream_t s)
{
    // grid z order: the (heavier) RGB-D pyramid jobs first; with banks the camera-frame jobs were run by odom_stage_current_dev
    static const int all[7] = { 0, 1, 2, 3, 4, 5, 6 }, model[4] = { 0, 2, 4, 6 };
    return o->banked ? launch_prep_all(o, in, model, 4, s) : launch_prep_all(o, in, all, 7, s);
}
static int launch_prep_all(hrbf_odometry* o, const OdomPrepInputs& in, const int* jobs, int njobs, cudaStream_t s)
{
    PrepAllArgs A;
    for (int k = 0; k < 7; ++k) A.jobs[k] = k < njobs ? jobs[k] : 0;
    A.cur_depth_only = (o->banked && o->bank[o->cur_bank].image_ready) ? 1 : 0;
    A.rows = o->height; A.cols = o->width; A.sel = in.sel; A.pose = in.pose_dev;
    A.dense_count = in.dense_count; A.dense_count_reset = in.dense_count_reset; A.dense_thresh = in.dense_thresh; A.curv_thr = o->curvThr; A.depth_cutoff = o->maxDepthRGB;
    A.vm = (const float4*)in.vm; A.nm = (const float4*)in.nm; A.vm_alt = (const float4*)in.vm_alt; A.nm_alt = (const float4*)in.nm_alt;
    A.vc = (const float4*)in.vc; A.nc = (const float4*)in.nc;
    A.k1m = (const float4*)in.k1m; A.k2m = (const float4*)in.k2m; A.k1m_alt = (const float4*)in.k1m_alt; A.k2m_alt = (const float4*)in.k2m_alt;
    A.k1c = (const float4*)in.k1c; A.k2c = (const float4*)in.k2c;
    A.w = in.w; A.w_alt = in.w_alt;
    A.o_vg = pyr_out(o, M_VG); A.o_ng = pyr_out(o, M_NG); A.o_vc = pyr_out(o, M_VC); A.o_nc = pyr_out(o, M_NC);
    A.o_k1g = pyr_out(o, M_K1G); A.o_k2g = pyr_out(o, M_K2G); A.o_k1c = pyr_out(o, M_K1C); A.o_k2c = pyr_out(o, M_K2C);
    for (int l = 0; l < 3; ++l) { A.o_w[l] = o->maps[M_W][l]; A.w_pitch[l] = o->cols(l); }
    A.rgbd[0].rgba = (const uchar4*)in.rgba_m; A.rgbd[0].rgba_alt = (const uchar4*)in.rgba_m_alt;
    A.rgbd[0].vertex = (const float4*)in.vm; A.rgbd[0].vertex_alt = (const float4*)in.vm_alt;
    A.rgbd[0].rgb8 = nullptr; A.rgbd[1].rgb8 = in.rgb8_c;
    A.rgbd[1].rgba = (const uchar4*)in.rgba_c; A.rgbd[1].rgba_alt = nullptr;
    A.rgbd[1].vertex = (const float4*)in.vc; A.rgbd[1].vertex_alt = nullptr;
    for (int l = 0; l < 3; ++l) {
        A.rgbd[0].img[l] = o->lastImage[l]; A.rgbd[0].depth[l] = o->lastDepth[l];
        A.rgbd[1].img[l] = o->nextImage[l]; A.rgbd[1].depth[l] = o->nextDepth[l];
    }
    A.pyr_bx = div_up(o->width, 32); A.pyr_by = div_up(o->height, 8);
    A.rgbd_bx = div_up(o->cols(2), 8); A.rgbd_by = div_up(o->rows(2), 8);
    const dim3 grid(A.pyr_bx > A.rgbd_bx ? A.pyr_bx : A.rgbd_bx, A.pyr_by > A.rgbd_by ? A.pyr_by : A.rgbd_by, njobs);
    HRBF_LAUNCH_PDL(prep_all_kernel, dim3(grid), dim3(256), 0, s, A);
    return HRBF_OK;
}
int odom_init_icp_weight_dev(hrbf_odometry* o, const float* w, const float* w_alt, const int* sel, cudaStream_t s)
{
    pyr_weight_kernel<<<pyr_grid(o), 256, 0, s>>>(w, o->height, o->width, o->maps[M_W][0], o->cols(0), o->maps[M_W][1], o->cols(1), o->maps[M_W][2], o->cols(2), w_alt, sel);
    HRBF_KERNEL_CHECK();
    return HRBF_OK;
}
}  // namespace hrbf
extern "C" {
int hrbf_odometry_init_icp_model(hrbf_odometry* o, const float* v, const float* n, float depthCutoff, const float* pose16, void* stream)
{
    (void)depthCutoff;
    HRBF_CHECK_ARG(o && v && n && pose16);
    float* dpose = nullptr;
    if (int rc = upload_pose(o, pose16, (cudaStream_t)stream, &dpose)) return rc;
    return odom_init_icp_model_dev(o, v, n, nullptr, nullptr, nullptr, dpose, (cudaStream_t)stream);
}
int hrbf_odometry_init_curvature(hrbf_odometry* o, const float* k1, const float* k2, void* stream)
{
    HRBF_CHECK_ARG(o && k1 && k2);
    pyr_pair_kernel<PYR_K><<<pyr_grid(o), 256, 0, (cudaStream_t)stream>>>((const float4*)k1, (const float4*)k2, o->height, o->width, o->curvThr, nullptr,
                                                                        pyr_out(o, M_K1C), pyr_out(o, M_K2C), nullptr, 0.f, nullptr, nullptr, nullptr);
    HRBF_KERNEL_CHECK();
    return HRBF_OK;
}
int hrbf_odometry_init_curvature_model(hrbf_odometry* o, const float* k1, const float* k2, const float* pose16, void* stream)
{
    HRBF_CHECK_ARG(o && k1 && k2 && pose16);
    float* dpose = nullptr;
    if (int rc = upload_pose(o, pose16, (cudaStream_t)stream, &dpose)) return rc;
    return odom_init_curvature_model_dev(o, k1, k2, nullptr, nullptr, nullptr, dpose, (cudaStream_t)stream);
}
int hrbf_odometry_init_icp_weight(hrbf_odometry* o, const float* w, void* stream)
{
    HRBF_CHECK_ARG(o && w);
    return odom_init_icp_weight_dev(o, w, nullptr, nullptr, (cudaStream_t)stream);
}
int hrbf_odometry_fill_neutral_curvature(hrbf_odometry* o, void* stream)
{
    HRBF_CHECK_ARG(o);
    cudaStream_t s = (cudaStream_t)stream;
    for (int l = 0; l < 3; ++l) {
        const size_t P = (size_t)o->rows(l) * o->cols(l);
        for (int m : { M_K1G, M_K2G, M_K1C, M_K2C }) HRBF_CUDA(cudaMemsetAsync(o->maps[m][l], 0, 4 * P * sizeof(float), s));
        std::vector<float> ones(P, 1.0f);
        HRBF_CUDA(cudaMemcpyAsync(o->maps[M_W][l], ones.data(), P * sizeof(float), cudaMemcpyHostToDevice, s));
        HRBF_CUDA(cudaStreamSynchronize(s));
    }
    o->pack_dirty_curr = o->pack_dirty_model = true;
    return HRBF_OK;
}

static int populate_rgbd(hrbf_odometry* o, const unsigned char* rgba, float** depths, unsigned char** images, cudaStream_t s,
                         const unsigned char* rgba_alt = nullptr, const int* sel = nullptr)
{
    const dim3 b(32, 8);
    HRBF_CUDA(cudaMemcpyAsync(depths[0], o->vdepth_tmp, (size_t)o->width * o->height * 4, cudaMemcpyDeviceToDevice, s));
    for (int i = 0; i + 1 < 3; ++i) {
        pyrdown_gauss_f32_kernel<<<grid2d(o->cols(i + 1), o->rows(i + 1), b), b, 0, s>>>(o->rows(i), o->cols(i), depths[i], depths[i + 1]);
        HRBF_KERNEL_CHECK();
    }
    const int n = o->width * o->height;
    rgba_to_intensity_kernel<<<div_up(n, 256), 256, 0, s>>>(n, (const uchar4*)rgba, images[0], (const uchar4*)rgba_alt, sel);
    HRBF_KERNEL_CHECK();
    for (int i = 0; i + 1 < 3; ++i) {
        pyrdown_gauss_u8_kernel<<<grid2d(o->cols(i + 1), o->rows(i + 1), b), b, 0, s>>>(o->rows(i), o->cols(i), images[i], images[i + 1]);
        HRBF_KERNEL_CHECK();
    }
    return HRBF_OK;
}
int hrbf_odometry_init_rgb(hrbf_odometry* o, const unsigned char* rgba, void* stream)
{ HRBF_CHECK_ARG(o && rgba); return populate_rgbd(o, rgba, o->nextDepth, o->nextImage, (cudaStream_t)stream); }
int hrbf_odometry_init_rgb_model(hrbf_odometry* o, const unsigned char* rgba, void* stream)
{ HRBF_CHECK_ARG(o && rgba); return populate_rgbd(o, rgba, o->lastDepth, o->lastImage, (cudaStream_t)stream); }
}  // extern "C"
namespace hrbf {
int odom_init_rgb_model_dev(hrbf_odometry* o, const unsigned char* rgba, const unsigned char* rgba_alt, const int* sel, cudaStream_t s)
{ return populate_rgbd(o, rgba, o->lastDepth, o->lastImage, s, rgba_alt, sel); }
}
extern "C" {
int hrbf_odometry_init_first_rgb(hrbf_odometry* o, const unsigned char* rgba, void* stream)
{
    HRBF_CHECK_ARG(o && rgba);
    cudaStream_t s = (cudaStream_t)stream;
    const dim3 b(32, 8);
    const int n = o->width * o->height;
    rgba_to_intensity_kernel<<<div_up(n, 256), 256, 0, s>>>(n, (const uchar4*)rgba, o->lastNextImage[0], nullptr, nullptr);
    HRBF_KERNEL_CHECK();
    for (int i = 0; i + 1 < 3; ++i) {
        pyrdown_gauss_u8_kernel<<<grid2d(o->cols(i + 1), o->rows(i + 1), b), b, 0, s>>>(o->rows(i), o->cols(i), o->lastNextImage[i], o->lastNextImage[i + 1]);
        HRBF_KERNEL_CHECK();
    }
    return HRBF_OK;
}

static void swap_so3_images(hrbf_odometry* o)
{   // RGBDOdometry.cpp:1239-1245 : pointer swap.  Captured graphs bake buffer addresses, so the
    // graph key carries the swap parity (two alternating graph sets).
    for (int i = 0; i < 3; ++i) std::swap(o->lastNextImage[i], o->nextImage[i]);
    o->so3_parity ^= 1;
}

int hrbf_odometry_get_incremental_transformation(hrbf_odometry* o, float* trans, float* rot, int rgbOnly, float icpWeight, int pyramid,
                                                 int fastOdom, int so3, int if_curvature_info, int index_frame, hrbf_track_stats* stats, void* stream)
{
    (void)index_frame;
    HRBF_CHECK_ARG(o && trans && rot);
    cudaStream_t s = (cudaStream_t)stream;
    if (int rc = repack_if_dirty(o, s)) return rc;
    memcpy(o->h_pose, rot, 36);
    memcpy(o->h_pose + 9, trans, 12);
    int nk = 1;
    if (o->use_graph) {
        cudaGraphExec_t exec = nullptr;
        if (int rc = get_graph(o, rgbOnly != 0, icpWeight, pyramid != 0, fastOdom != 0, so3 != 0, if_curvature_info != 0, true, &exec, &nk)) return rc;
        HRBF_CUDA(cudaGraphLaunch(exec, s));
        count_launch(nk);
    } else {
        HRBF_CUDA(cudaMemcpyAsync(o->pose_scratch + 12, o->h_pose, 12 * sizeof(float), cudaMemcpyHostToDevice, s));
        if (int rc = launch_track_persistent(o, s, rgbOnly != 0, icpWeight, pyramid != 0, fastOdom != 0, so3 != 0, if_curvature_info != 0,
                                             o->pose_scratch + 12, o->pose_scratch + 24)) return rc;
        HRBF_CUDA(cudaMemcpyAsync(o->h_pose + 12, o->pose_scratch + 24, 12 * sizeof(float), cudaMemcpyDeviceToHost, s));
        HRBF_CUDA(cudaMemcpyAsync(o->h_state, &o->work->st, sizeof(TrackState), cudaMemcpyDeviceToHost, s));
    }
    HRBF_CUDA(cudaStreamSynchronize(s));
    if (so3) swap_so3_images(o);
    memcpy(rot, o->h_pose + 12, 36);
    memcpy(trans, o->h_pose + 21, 12);
    if (stats) {
        const TrackState& h = *o->h_state;
        stats->lastICPError = h.lastICPError; stats->lastICPCount = h.lastICPCount;
        stats->lastRGBError = h.lastRGBError; stats->lastRGBCount = h.lastRGBCount;
        stats->lastSO3Error = h.lastSO3Error; stats->lastSO3Count = h.lastSO3Count;
        memcpy(stats->lastA, h.lastA, sizeof h.lastA); memcpy(stats->lastb, h.lastb, sizeof h.lastb);
        stats->icp_iterations_run = h.icp_iterations_run;
        stats->kernel_launches = nk;
    }
    return HRBF_OK;
}

int hrbf_odometry_track_async(hrbf_odometry* o, const float* prev_pose_dev, float* pose_out_dev, int rgbOnly, float icpWeight, int pyramid,
                              int fastOdom, int so3, int if_curvature_info, void* stream)
{
    HRBF_CHECK_ARG(o && prev_pose_dev && pose_out_dev);
    cudaStream_t s = (cudaStream_t)stream;
    if (int rc = repack_if_dirty(o, s)) return rc;
    HRBF_CUDA(cudaMemcpyAsync(o->pose_scratch + 12, prev_pose_dev, 12 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    if (o->use_graph) {
        int nk = 0;
        cudaGraphExec_t exec = nullptr;
        if (int rc = get_graph(o, rgbOnly != 0, icpWeight, pyramid != 0, fastOdom != 0, so3 != 0, if_curvature_info != 0, false, &exec, &nk)) return rc;
        HRBF_CUDA(cudaGraphLaunch(exec, s));
        count_launch(nk);
    } else {
        if (int rc = launch_track_persistent(o, s, rgbOnly != 0, icpWeight, pyramid != 0, fastOdom != 0, so3 != 0, if_curvature_info != 0,
                                             o->pose_scratch + 12, o->pose_scratch + 24)) return rc;
    }
    if (so3) swap_so3_images(o);
    HRBF_CUDA(cudaMemcpyAsync(pose_out_dev, o->pose_scratch + 24, 12 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    return HRBF_OK;
}


int hrbf_odometry_icp_step(hrbf_odometry* o, int level, const float* Rcurr, const float* tcurr, const float* Rprev_inv, const float* tprev,
                           int use_weight, int tiled, float* A, float* b, float* residual, double* sums29, void* stream)
{
    HRBF_CHECK_ARG(o && level >= 0 && level <= 2 && Rcurr && tcurr && Rprev_inv && tprev && A && b && residual);
    cudaStream_t s = (cudaStream_t)stream;
    if (int rc = repack_if_dirty(o, s)) return rc;
    TrackState h;
    HRBF_CUDA(cudaMemcpyAsync(&h, &o->work->st, sizeof h, cudaMemcpyDeviceToHost, s));
    HRBF_CUDA(cudaStreamSynchronize(s));
    memcpy(h.Rcurr, Rcurr, 36); memcpy(h.tcurr, tcurr, 12); memcpy(h.Rprev_inv, Rprev_inv, 36); memcpy(h.tprev, tprev, 12);
    h.done_level = -1; h.icp = 1; h.ticket = 0u;
    if (int rc = init_work_state(o->work, s, h)) return rc;
    const IcpArgs ia = icp_args(o, level, use_weight != 0);
    if (tiled && !o->useSearch) HRBF_CUDA(launch_icp_tile(o, ia, 0, level, -1, s, false));
    else if (o->useSearch) icp_reduce_kernel<true><<<reduce_blocks(ia.rows * ia.cols), kReduceThreads, 0, s>>>(ia, o->work, 0, level, -1);
    else icp_reduce_kernel<false><<<reduce_blocks(ia.rows * ia.cols), kReduceThreads, 0, s>>>(ia, o->work, 0, level, -1);
    HRBF_KERNEL_CHECK();
    double sums[32];
    HRBF_CUDA(cudaMemcpyAsync(sums, o->work->st.icp_sums, sizeof sums, cudaMemcpyDeviceToHost, s));
    HRBF_CUDA(cudaStreamSynchronize(s));
    unpack_host_se3(sums, A, b, residual);
    if (sums29) memcpy(sums29, sums, 29 * sizeof(double));
    return HRBF_OK;
}

int hrbf_odometry_time_kernel(hrbf_odometry* o, int which, int level, int with_update, int reps, float* avg_us, void* stream)
{
    HRBF_CHECK_ARG(o && avg_us && which >= 0 && which <= 7 && level >= 0 && level <= 2 && reps > 0 && reps <= 2000);
    cudaStream_t s = (cudaStream_t)stream;
    if (int rc = repack_if_dirty(o, s)) return rc;
    cudaEvent_t e0, e1;
    HRBF_CUDA(cudaEventCreate(&e0));
    HRBF_CUDA(cudaEventCreate(&e1));
    const IcpArgs ia = icp_args(o, level, true);
    const RgbResArgs ra = rgbres_args(o, level);
    const RgbStepArgs sa = rgbstep_args(o, level);
    const int nb = reduce_blocks(o->rows(level) * o->cols(level));
    if (which == 4) {
        // ONE launch of the persistent tracker doing `reps` ICP-only Gauss-Newton iterations at `level` (reduction + cross-CTA
        // exchange + fp64 solve each): the production form of the reduction.  Returns the time per iteration.
        int it[3] = { 0, 0, 0 };
        it[level] = reps;
        for (int r = 0; r < 2; ++r) {
            if (r == 1) HRBF_CUDA(cudaEventRecord(e0, s));
            // start pose = the one the last tracking call started from (TrackState begins with Rprev[9], tprev[3]): consistent with the loaded maps
            if (int rc = launch_track_persistent(o, s, false, 100.0f, true, false, false, true, (const float*)&o->work->st, o->pose_scratch + 24, it)) return rc;
        }
        HRBF_CUDA(cudaEventRecord(e1, s));
        HRBF_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        HRBF_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        *avg_us = ms * 1000.f / (float)reps;
        return HRBF_OK;
    }
    if (which == 6 || which == 7) {
        // COLD: before every launch a buffer larger than L2 is overwritten (so the maps come from HBM), and every launch is timed
        // alone between its own pair of events; avg_us = mean over `reps` launches.  6: tile kernel, 7: the per-pixel-gather kernel.
        const size_t flush_bytes = (size_t)256 << 20;
        void* flush = nullptr;
        HRBF_CUDA(cudaMalloc(&flush, flush_bytes));
        double total = 0.0;
        for (int r = 0; r < reps + 2; ++r) {
            HRBF_CUDA(cudaMemsetAsync(flush, r & 0xff, flush_bytes, s));
            HRBF_CUDA(cudaEventRecord(e0, s));
            if (which == 6) HRBF_CUDA(launch_icp_tile(o, ia, with_update ? 1 : 0, level, -1, s, false));
            else icp_reduce_kernel<false><<<nb, kReduceThreads, 0, s>>>(ia, o->work, with_update ? 1 : 0, level, -1);
            HRBF_CUDA(cudaEventRecord(e1, s));
            HRBF_CUDA(cudaEventSynchronize(e1));
            float ms = 0.f;
            HRBF_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            if (r >= 2) total += ms;
            count_launch();
        }
        cudaFree(flush);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        HRBF_CUDA(cudaGetLastError());
        *avg_us = (float)(total * 1000.0 / reps);
        return HRBF_OK;
    }
    // `reps` back-to-back launches captured into ONE graph (so the host's launch rate is not what is measured), replayed once
    // untimed and once between two events on `s`
    auto launch_once = [&](cudaStream_t q) {
        switch (which) {
        case 5: (void)launch_icp_tile(o, ia, with_update ? 1 : 0, level, -1, q, true); break;
        case 0: icp_reduce_kernel<false><<<nb, kReduceThreads, 0, q>>>(ia, o->work, with_update ? 1 : 0, level, -1); break;
        case 1: rgb_residual_kernel<<<nb, 256, 0, q>>>(ra, o->work, 1, level, 0, -1); break;
        case 2: rgb_step_kernel<<<nb, kReduceThreads, 0, q>>>(sa, -2.0f, o->work, with_update ? 1 : 0, level, -1); break;
        default: so3_reduce_kernel<<<reduce_blocks(o->rows(2) * o->cols(2)), kReduceThreads, 0, q>>>(o->lastNextImage[2], o->nextImage[2], o->rows(2), o->cols(2), o->work, 0); break;
        }
    };
    HRBF_CUDA(cudaStreamSynchronize(s));
    cudaGraph_t g = nullptr;
    cudaGraphExec_t ge = nullptr;
    HRBF_CUDA(cudaStreamBeginCapture(o->cap_stream, cudaStreamCaptureModeThreadLocal));
    for (int r = 0; r < reps; ++r) launch_once(o->cap_stream);
    HRBF_CUDA(cudaStreamEndCapture(o->cap_stream, &g));
    HRBF_CUDA(cudaGraphInstantiate(&ge, g, 0));
    cudaGraphDestroy(g);
    HRBF_CUDA(cudaGraphLaunch(ge, s));
    HRBF_CUDA(cudaEventRecord(e0, s));
    HRBF_CUDA(cudaGraphLaunch(ge, s));
    HRBF_CUDA(cudaEventRecord(e1, s));
    count_launch(2 * reps);
    HRBF_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    HRBF_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaGraphExecDestroy(ge);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    HRBF_CUDA(cudaGetLastError());
    *avg_us = ms * 1000.f / (float)reps;
    return HRBF_OK;
}

const float* hrbf_odometry_map(const hrbf_odometry* o, int which, int level, size_t* step)
{
    if (!o || which < 0 || which >= M_COUNT || level < 0 || level > 2) return nullptr;
    if (step) *step = (size_t)o->cols(level) * sizeof(float);
    return o->maps[which][level];
}
const unsigned char* hrbf_odometry_image(const hrbf_odometry* o, int which, int level)
{
    if (!o || level < 0 || level > 2) return nullptr;
    if (which == 3) return (o->banked && level == 2) ? o->bank[o->cur_bank].nextImage[2] : nullptr;      // what the staged SO3 pre-alignment read as "next"
    return which == 0 ? o->lastImage[level] : which == 1 ? o->nextImage[level] : o->lastNextImage[level];
}
const short* hrbf_odometry_gradient(const hrbf_odometry* o, int axis, int level)
{
    if (!o || level < 0 || level > 2 || axis < 0 || axis > 1) return nullptr;
    return axis == 0 ? o->dIdx[level] : o->dIdy[level];
}
const unsigned char* hrbf_odometry_candidates(const hrbf_odometry* o, int level)
{
    if (!o || level < 0 || level > 2) return nullptr;
    return o->cand[level];
}
const float* hrbf_odometry_depth(const hrbf_odometry* o, int which, int level)
{
    if (!o || level < 0 || level > 2) return nullptr;
    return which == 0 ? o->lastDepth[level] : o->nextDepth[level];
}

}  // extern "C"

#include "cudafuncs_api.inl"
