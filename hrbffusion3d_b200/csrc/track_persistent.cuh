// track_persistent.cuh -- the whole coarse-to-fine tracking loop of RGBDOdometry::getIncrementalTransformation
// (Core/src/Utils/RGBDOdometry.cpp:796-1249) as ONE persistent cooperative kernel: one CTA per SM, the Gauss-Newton
// state in shared memory, tagged-word exchanges (no grid barrier) between the reduction phases.
//
// Every CTA reduces the same per-CTA partial sums in the same fixed order and runs the same fp64 solve, so all
// CTAs hold bit-identical copies of the state and only ONE exchange is needed per reduction (none to
// broadcast the new pose).  Compared with one kernel per reduction (icp_reduce_kernel & co., kept for the
// step-function ABI) this removes ~70 dependent kernel boundaries per frame, the ticket atomics and the
// __threadfence of the last-block-done scheme.
#pragma once
#define GN_STAMP(k) do { if (dbg && dbg_n_ptr && blockIdx.x == gridDim.x / 2 && threadIdx.x == 0) { long long t_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_)); dbg[*dbg_n_ptr < 500 ? (*dbg_n_ptr)++ : 499] = ((long long)(k) << 56) | (t_ & 0x00ffffffffffffffll); } } while (0)
#include "icp_tile.cuh"

namespace hrbf {

// Threads per CTA (x 148 CTAs, one per SM) are a template parameter of the tracker, chosen per odometry object at run time
// (hrbf_odometry_set_tracker_threads): 512 x 128 registers fill an SM's register file -- the lowest latency for ONE sequence;
// 256 leave half of every SM to the kernels of other sequences' pipelines (several fusion objects on their own streams).
constexpr int kTrackThreadsDefault = 512;
// result of the SO3 pre-alignment when it runs ahead of the tracking call (so3_prealign_kernel)
struct So3Pre { double resultR[9]; float lastSO3Error, lastSO3Count; };
#ifndef HRBF_TRACK_POLL_SLEEP
#define HRBF_TRACK_POLL_SLEEP 100      // ns between polling rounds of the shared-SM shape
#endif

struct TrackLevel {
    IcpArgs icp;
    RgbResArgs res;
    RgbStepArgs step;
    unsigned char* cand;     // per pixel: 1 = passes the pose-independent tests of computeRgbResidual (see rgb_static_candidate)
    int iters;
};
struct TrackParams {
    TrackLevel lvl[3];
    const unsigned char *so3_last, *so3_next;      // level-2 images of the SO3 pre-alignment
    int icp, rgb, rgbOnly, so3;
    float icpWeight;
    const float* prev_pose;                        // device R[9], t[3]
    float* pose_out;                               // device R[9], t[3] (may alias prev_pose: written by CTA 0 after the last exchange)
    // optional frame-pipeline epilogue (HRBFFusion.cpp:1109-1123, 1195): null = not written
    float* last_pose_out;                          // copy of the pose tracking started from
    float* inv_pose_out;                           // rigid inverse of the new pose (the splats' t_inv)
    float* weighting_out; float weight_multiplier; // fusion weight from the inter-frame motion
    float* traj_out;                               // this frame's row of the trajectory
    TrackState* st_global;                         // camera in, statistics out
    int cand_ready;                                // the Sobel images and candidate masks were built by sobel_cand_kernel (staged frame)
    const So3Pre* so3_pre;                         // non-null: the SO3 pre-alignment was run by so3_prealign_kernel (staged frame), this is its result
    int max_slots;                                 // dynamic shared memory holds max_slots x kTrackThreads RgbSlots ...
    IcpTileGeom tile[3];                           // ... followed by the resident ICP tile of the level being worked on (icp_tile.cuh):
    IcpTileMaps tmaps[3];                          // tensor maps of the level's packed records for that geometry
    int resident[3];                               // level l keeps its tile of packed records + model window in shared memory for all its iterations
    unsigned long long* ll_f;                      // [2][gridDim.x][64] (float, tag) words: the per-CTA partial sums
    unsigned long long* ll_i;                      // [2][gridDim.x][kIntStride] (int, tag) words: {count, sum diff^2} of computeRgbResidual in words 0-1
    unsigned int epoch;                            // launch counter (20 bits, never 0): tags of older launches never match
    long long* dbg;                                // optional: %globaltimer stamps of CTA 0 (profiling builds)
};

// Cross-CTA exchange without a grid barrier: every value travels as ONE 64-bit word {payload, tag} (the scheme NCCL's
// LL protocol uses over NVLink, here through L2).  A reader simply re-loads a word until its tag is the one of the
// current exchange: no atomics, no fences, no separate barrier round trip -- the data IS the flag.  Words are
// double-buffered by exchange parity: a CTA can only be one exchange ahead of the slowest one (it needs everybody's
// words of exchange n to get to n+1), so the words of exchange n are never overwritten before they were all read.
__device__ __forceinline__ void ll_store(unsigned long long* p, unsigned int payload, unsigned int tag)
{
    const unsigned long long w = ((unsigned long long)tag << 32) | payload;
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ unsigned long long ll_load(const unsigned long long* p)
{
    unsigned long long w;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    return w;
}

// 32 per-thread sums -> this CTA's partial, published by warp 0 as 32 LL words dst[0..31]
template <int kTrackThreads>
__device__ __forceinline__ void block_partial32(float (&acc)[32], float (*s_w)[32], unsigned long long* dst, unsigned int tag)
{
    constexpr int kTrackWarps = kTrackThreads / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float r = warp_reduce32_transpose(acc);
    s_w[warp][lane] = r;
    __syncthreads();
    if (threadIdx.x < 32) {
        float p = 0.f;
#pragma unroll
        for (int w = 0; w < kTrackWarps; ++w) p += s_w[w][threadIdx.x];
        ll_store(dst + threadIdx.x, __float_as_uint(p), tag);
    }
    __syncthreads();
}

// every CTA: grid totals of the 2 x 32 partial sums, fp64, fixed order -> out[64] (shared).
// Thread t owns value (t & 63) of CTA slice (t >> 6): ALL its loads are issued at once (one L2 round trip), words that
// are not there yet are re-polled; the additions run in a fixed order -> bit-identical totals in every CTA.
// lo / hi: whether the first / second 32 values were published in this exchange (ICP / RGB).
template <int kTrackThreads>
__device__ __forceinline__ void all_reduce_partials(const unsigned long long* part /* [gridDim.x][64] */, unsigned int tag, bool lo, bool hi,
                                                    double (*s_d)[64], double* out /* [64] */)
{
    constexpr int kSlices = kTrackThreads / 64, U = kTrackThreads == 256 ? 20 : (160 + kSlices - 1) / kSlices;      // up to 160 CTAs per batch (80 for 256 threads)
    static_assert(kTrackThreads % 64 == 0, "slice layout");
    const int v = threadIdx.x & 63, q = threadIdx.x >> 6;
    const bool active = v < 32 ? lo : hi;
    double a = 0.0;
    if (active) {
        for (unsigned int b0 = q; b0 < gridDim.x; b0 += U * kSlices) {
            unsigned long long w[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned int b = b0 + u * kSlices;
                w[u] = (b < gridDim.x) ? ll_load(part + (size_t)b * 64 + v) : ((unsigned long long)tag << 32);
            }
            bool all;
            do {
                all = true;
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if ((unsigned int)(w[u] >> 32) != tag) { w[u] = ll_load(part + (size_t)(b0 + u * kSlices) * 64 + v); all = false; }
                // sharing the SM with other sequences' kernels (256-thread shape): do not burn their issue slots and L2 bandwidth while waiting
                if (kTrackThreads < kTrackThreadsDefault && !all) __nanosleep(HRBF_TRACK_POLL_SLEEP);
            } while (!all);
#pragma unroll
            for (int u = 0; u < U; ++u) a += (double)__uint_as_float((unsigned int)w[u]);
        }
    }
    s_d[q][v] = a;
    __syncthreads();
    if (threadIdx.x < 64) {
        double r = 0.0;
#pragma unroll
        for (int w = 0; w < kSlices; ++w) r += s_d[w][threadIdx.x];
        out[threadIdx.x] = r;
    }
    __syncthreads();
}

// every CTA: grid totals of the {count, sum} integer pair -> s_i[0..1] (shared).
// A CTA's pair sits alone in a 1-KB stride (kIntStride words): the address -> L2-slice hash uses bits 8 and 10-27, so the
// 148 polled pairs land on different slices.  (Packed into 2.4 KB they shared ~6 of 184 slices and the ~44 K polling loads
// of one round queued there for ~4 us -- the round trip this exchange replaces took 1.5.)  One 16-byte load fetches both
// words; each half is still written by ONE 8-byte store, so a half whose tag matches is complete.
constexpr int kIntStride = 128;      // 64-bit words
__device__ __forceinline__ void all_reduce_int2(const unsigned long long* ip /* [gridDim.x][kIntStride] */, unsigned int tag, int* s_i)
{
    if (threadIdx.x < 2) s_i[threadIdx.x] = 0;
    __syncthreads();
    int c2 = 0, g2 = 0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
        unsigned long long w0, w1;
        const unsigned long long* q = ip + (size_t)b * kIntStride;
        for (;;) {
            asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(q) : "memory");
            if ((unsigned int)(w0 >> 32) == tag && (unsigned int)(w1 >> 32) == tag) break;
            __nanosleep(40);
        }
        c2 += (int)(unsigned int)w0; g2 += (int)(unsigned int)w1;
    }
    c2 = __reduce_add_sync(0xffffffffu, c2);
    g2 = __reduce_add_sync(0xffffffffu, g2);
    if ((threadIdx.x & 31) == 0 && (c2 | g2)) { atomicAdd(&s_i[0], c2); atomicAdd(&s_i[1], g2); }
    __syncthreads();
}

__device__ __forceinline__ long long gtimer() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define TP_STAMP(slot) do { if (p.dbg && blockIdx.x == gridDim.x / 2 && threadIdx.x == 0) p.dbg[dbg_n < 500 ? dbg_n++ : 499] = ((long long)(slot) << 56) | (gtimer() & 0x00ffffffffffffffll); } while (0)

// One correspondence of computeRgbResidual kept on chip between the residual and the step pass (12 B instead of the
// reference's 16-B DataTerm in HBM + the 12-B cloud point: the point is rebuilt from d0).  uv = v0 << 16 | u0, -1 = none.
struct RgbSlot { float diff, d0; int uv; };

constexpr int kRgbInFlight = 5;      // pixels per thread whose loads are in flight together (640x480 level 0: 5 slots per thread)

// computeRgbResidual over this CTA's pixel range [begin, end): pixel begin + tid + m * kTrackThreads -> slot m of this thread.
// Same arithmetic as rgb_residual_pixel, staged so that the (dependent) loads of kRgbInFlight pixels overlap:
// candidate byte + next depth  ->  projection, gather of last depth / last image  ->  residual.
template <int kTrackThreads>
__device__ __forceinline__ void rgb_residual_pass(const RgbResArgs& a, const unsigned char* cand, const float* s_k, int begin, int end,
                                                  RgbSlot* s_slots, int& cnt, int& sig)
{
    const int cols = a.cols, rows = a.rows, tid = threadIdx.x;
    int m0 = 0;
    for (int k0 = begin + tid; k0 < end; k0 += kRgbInFlight * kTrackThreads, m0 += kRgbInFlight) {
        bool c[kRgbInFlight];
        float d1[kRgbInFlight], td1[kRgbInFlight], d0[kRgbInFlight];
        int uv[kRgbInFlight];
        unsigned char li[kRgbInFlight], ni[kRgbInFlight];
#pragma unroll
        for (int u = 0; u < kRgbInFlight; ++u) {
            const int k = k0 + u * kTrackThreads;
            const bool in = k < end;
            c[u] = in && cand[in ? k : begin] != 0;           // plain load: written by this thread in the prep phase
            d1[u] = __ldg(a.nextDepth + (in ? k : begin));
        }
#pragma unroll
        for (int u = 0; u < kRgbInFlight; ++u) {
            const int k = k0 + u * kTrackThreads;
            uv[u] = -1; d0[u] = 0.f; td1[u] = 0.f; li[u] = 0; ni[u] = 0;
            if (c[u]) {
                const int y = k / cols, x = k - y * cols;
                td1[u] = rgb_warp_row(s_k, 2, x, y, d1[u]);
                const int u0 = __float2int_rn(rgb_warp_row(s_k, 0, x, y, d1[u]) / td1[u]);
                const int v0 = __float2int_rn(rgb_warp_row(s_k, 1, x, y, d1[u]) / td1[u]);
                if (u0 >= 0 && v0 >= 0 && u0 < cols && v0 < rows) {
                    uv[u] = (v0 << 16) | u0;
                    d0[u] = __ldg(a.lastDepth + (size_t)v0 * cols + u0);
                    li[u] = __ldg(a.lastImage + (size_t)v0 * cols + u0);
                    ni[u] = __ldg(a.nextImage + k);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kRgbInFlight; ++u) {
            const int k = k0 + u * kTrackThreads;
            if (k < end) {
                RgbSlot sl;
                sl.uv = -1; sl.diff = 0.f; sl.d0 = 0.f;
                if (uv[u] != -1 && d0[u] > 0 && fabsf(td1[u] - d0[u]) <= a.maxDepthDelta && li[u] != 0) {
                    sl.diff = (float)ni[u] - (float)li[u];
                    sl.d0 = d0[u];
                    sl.uv = uv[u];
                    cnt += 1;
                    sig += (int)(sl.diff * sl.diff);
                }
                s_slots[(m0 + u) * kTrackThreads + tid] = sl;
            }
        }
    }
}

// rgbStep (reduce.cu:718-811) from the slots: the cloud point of projectToPointCloud (cudafuncs.cu:995-1013) is rebuilt
// from d0 with the same expression, the gradients are read at the pixel itself (DataTerm.one == the pixel).
template <int kTrackThreads>
__device__ __forceinline__ void rgb_step_pass(const RgbStepArgs& a, float sigma, const RgbSlot* s_slots, int begin, int end,
                                              float ifx, float ify, float cx, float cy, float (&acc)[32])
{
    const int tid = threadIdx.x;
    int m0 = 0;
    for (int k0 = begin + tid; k0 < end; k0 += kRgbInFlight * kTrackThreads, m0 += kRgbInFlight) {
        RgbSlot sl[kRgbInFlight];
        short ix[kRgbInFlight], iy[kRgbInFlight];
#pragma unroll
        for (int u = 0; u < kRgbInFlight; ++u) {
            const int k = k0 + u * kTrackThreads;
            sl[u].uv = -1; sl[u].diff = 0.f; sl[u].d0 = 1.f; ix[u] = 0; iy[u] = 0;
            if (k < end) {
                sl[u] = s_slots[(m0 + u) * kTrackThreads + tid];
                if (sl[u].uv != -1) { ix[u] = a.dIdx[k]; iy[u] = a.dIdy[k]; }      // plain loads: written by this thread in the prep phase
            }
        }
#pragma unroll
        for (int u = 0; u < kRgbInFlight; ++u) {
            if (k0 + u * kTrackThreads >= end) continue;
            float row[7] = { 0, 0, 0, 0, 0, 0, 0 };
            float rgb_weight = 1.f;
            const bool valid = sl[u].uv != -1;
            if (valid) {
                const float diff = sl[u].diff;
                float w = sigma + fabsf(diff);
                w = w > 1.19209290E-07F ? 1.0f / w : 1.0f;
                if (sigma == -1.f) w = 1.f;
                row[6] = -w * diff;
                const int u0 = sl[u].uv & 0xffff, v0 = sl[u].uv >> 16;
                const float pz = sl[u].d0;
                const float px = (u0 - cx) * pz * ifx, py = (v0 - cy) * pz * ify;
                const float invz = 1.0f / pz;
                const float gx = w * a.sobelScale * (float)ix[u], gy = w * a.sobelScale * (float)iy[u];
                const float v0_ = gx * a.fx * invz, v1 = gy * a.fy * invz;
                const float v2 = -(v0_ * px + v1 * py) * invz;
                row[0] = v0_; row[1] = v1; row[2] = v2;
                row[3] = -pz * v1 + py * v2;
                row[4] = pz * v0_ - px * v2;
                row[5] = -py * v0_ + px * v1;
                if (a.use_grad_weight) {
                    const float gm = sqrtf(gx * gx + gy * gy);
                    rgb_weight = expf(-0.5f * (10.f / gm) * (10.f / gm));
                }
            }
            accumulate_row7(acc, row, rgb_weight, valid);
        }
    }
}

// dynamic shared memory of the persistent tracker: the RGB slots, max_slots x kTrackThreads (rounded up to 128 B), then the resident ICP tile
inline size_t track_slots_bytes(int max_slots, int threads) { return ((size_t)max_slots * threads * sizeof(RgbSlot) + 127) & ~(size_t)127; }
// one tile per CTA: 4 x (ctas / 4) tiles when the CTA count allows, a narrower window than the stand-alone kernel's (the bounding box is
// taken under the pose the level starts with; 2 pixels of margin on the low side, the rest of the slack on the high side)
inline IcpTileGeom track_tile_geom(int rows, int cols, int ctas)
{
    IcpTileGeom g;
    g.ncol = (ctas % 4 == 0) ? 4 : (ctas % 2 == 0) ? 2 : 1;
    g.nrow = ctas / g.ncol;
    g.tw = (div_up(cols, g.ncol) + 3) & ~3;
    g.th = div_up(rows, g.nrow);
    g.mw = g.tw + 8;
    g.mh = g.th + 6;
    icp_tile_boxes(g);
    g.ctas = ctas; g.threads = 0;
    return g;
}

template <int kTrackThreads>
__global__ void __maxnreg__(128) track_persistent_kernel(const __grid_constant__ TrackParams p)
{
    constexpr int kTrackWarps = kTrackThreads / 32;
    pdl_wait();
    extern __shared__ __align__(128) unsigned char s_dyn[];
    RgbSlot* s_slots = reinterpret_cast<RgbSlot*>(s_dyn);
    unsigned char* s_tile = s_dyn + (((size_t)p.max_slots * kTrackThreads * sizeof(RgbSlot) + 127) & ~(size_t)127);
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ int s_box[4];
    __shared__ IcpTileView s_tv;
    uint32_t par0 = 0, par1 = 0;                         // phase parities of the two mbarriers (uniform over the CTA)
    if (threadIdx.x == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); mbar_fence_init(); }
    int dbg_n = 0;
    __shared__ TrackState S;
    __shared__ float s_w[kTrackWarps][32];
    __shared__ double s_d[kTrackThreads / 64][64];
    __shared__ double s_tot[64];
    __shared__ int s_i[2];
    unsigned int phase = 0, iphase = 0;                  // exchange counters (float partials / integer pair), identical in every CTA
    const unsigned int tag_base = p.epoch << 12;
    const int tid = threadIdx.x, gtid = blockIdx.x * blockDim.x + threadIdx.x, gstride = gridDim.x * blockDim.x;

    // ---- state initialisation (track_begin_kernel), identical in every CTA ----
    int first_level = 0;
    for (int l = 2; l >= 0; --l) if (p.lvl[l].iters > 0) { first_level = l; break; }
    if (tid == 0) {
        S.fx = p.st_global->fx; S.fy = p.st_global->fy; S.cx = p.st_global->cx; S.cy = p.st_global->cy;
        for (int k = 0; k < 9; ++k) { S.Rprev[k] = p.prev_pose[k]; S.Rcurr[k] = p.prev_pose[k]; }
        for (int k = 0; k < 3; ++k) { S.tprev[k] = p.prev_pose[9 + k]; S.tcurr[k] = p.prev_pose[9 + k]; }
        inv3f(S.Rprev, S.Rprev_inv);
        for (int k = 0; k < 16; ++k) S.resultRt[k] = (k % 5 == 0) ? 1.0 : 0.0;
        for (int k = 0; k < 9; ++k) { S.resultR[k] = S.lastResultR[k] = (k % 4 == 0) ? 1.0 : 0.0; S.R_lr[k] = (k % 4 == 0) ? 1.f : 0.f; }
        S.so3_lastError = FLT_MAX / 2; S.so3_lastCount = FLT_MAX / 2; S.so3_done = 0;
        S.icp = p.icp; S.rgb = p.rgb; S.rgbOnly = p.rgbOnly; S.so3 = p.so3; S.icpWeight = p.icpWeight;
        S.done_level = -1; S.rgb_count = 0; S.rgb_sigma = 0; S.sigmaVal = 0.f;
        S.lastICPError = 0; S.lastICPCount = 0; S.lastRGBError = FLT_MAX; S.lastRGBCount = 0; S.lastSO3Error = 0; S.lastSO3Count = 0;
        S.icp_iterations_run = 0; S.ticket = 0u;
        for (int k = 0; k < 32; ++k) { S.icp_sums[k] = 0; S.rgb_sums[k] = 0; }
        for (int k = 0; k < 36; ++k) S.lastA[k] = 0;
        for (int k = 0; k < 6; ++k) S.lastb[k] = 0;
        const double I3[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
        const double I4[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
        if (p.so3 && p.so3_pre != nullptr) {
#pragma unroll
            for (int k = 0; k < 9; ++k) S.resultR[k] = p.so3_pre->resultR[k];
            S.lastSO3Error = p.so3_pre->lastSO3Error; S.lastSO3Count = p.so3_pre->lastSO3Count;
        } else if (p.so3) update_so3_mats(&S, I3);
        else if (p.rgb) update_krk(&S, I4, first_level);
    }

    // ---- RGB branch prep: Sobel of the next image + the pose-independent part of computeRgbResidual, all levels.
    // Pixel k of a level belongs to CTA range [begin, end) and, inside it, to thread (k - begin) % kTrackThreads: the SAME
    // mapping as the residual / step passes below, so every thread only ever re-reads what it wrote itself (no barrier).
    if (p.rgb && !p.cand_ready) {
        for (int l = 0; l < 3; ++l) {
            const RgbResArgs& r = p.lvl[l].res;
            const int N = r.rows * r.cols;
            const int begin = (int)(((long long)N * blockIdx.x) / gridDim.x), end = (int)(((long long)N * (blockIdx.x + 1)) / gridDim.x);
            for (int k = begin + tid; k < end; k += kTrackThreads) {
                const int y = k / r.cols, x = k - y * r.cols;
                sobel_pixel(r.rows, r.cols, r.nextImage, const_cast<short*>(r.dIdx), const_cast<short*>(r.dIdy), x, y);
                p.lvl[l].cand[k] = rgb_static_candidate(r, k, r.dIdx[k], r.dIdy[k]) ? 1 : 0;     // plain loads of this thread's own stores
            }
        }
    }
    __syncthreads();

    // ---- SO3 pre-alignment (RGBDOdometry.cpp:827-914), level 2 ----
    if (p.so3) {
        const int rows = p.lvl[2].res.rows, cols = p.lvl[2].res.cols, N = rows * cols;
        for (int it = 0; it < 10 && p.so3_pre == nullptr; ++it) {
            if (S.so3_done) break;
            float acc[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) acc[k] = 0.f;
            // so3_pixel reads s_m = basis[9], kinv[9], krlr[9]: contiguous in TrackState
            for (int k = gtid; k < N; k += gstride) so3_pixel(p.so3_last, p.so3_next, rows, cols, S.so3_basis, k, acc);
            unsigned long long* part = p.ll_f + (size_t)(phase & 1) * gridDim.x * 64;
            const unsigned int tag = tag_base | (phase + 1);
            block_partial32<kTrackThreads>(acc, s_w, part + (size_t)blockIdx.x * 64, tag);
            all_reduce_partials<kTrackThreads>(part, tag, true, false, s_d, s_tot);
            if (tid < 16) S.so3_sums[tid] = s_tot[tid];
            __syncthreads();
            if (tid == 0) so3_update(&S);
            __syncthreads();
            ++phase;
        }
        if (tid == 0) {      // track_after_so3_kernel
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) S.resultRt[a * 4 + b] = S.resultR[a * 3 + b];
            if (S.rgb) {
                double Rt[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) Rt[k] = S.resultRt[k];
                update_krk(&S, Rt, first_level);
            }
        }
        __syncthreads();
    }

    // ---- coarse-to-fine Gauss-Newton (RGBDOdometry.cpp:946-1206) ----
    for (int l = 2; l >= 0; --l) {
        const TrackLevel& L = p.lvl[l];
        if (L.iters == 0) continue;
        int next_lower = -1;
        for (int q = l - 1; q >= 0; --q) if (p.lvl[q].iters > 0) { next_lower = q; break; }
        const int N = L.icp.rows * L.icp.cols;
        const int begin = (int)(((long long)N * blockIdx.x) / gridDim.x), end = (int)(((long long)N * (blockIdx.x + 1)) / gridDim.x);
        const float ifx = 1.0f / L.step.fx, ify = 1.0f / L.step.fy, lcx = L.icp.cx, lcy = L.icp.cy;     // projectToPointCloud's 1/fx, 1/fy, cx, cy of this level
        // ---- resident ICP tile: this CTA's tile of current-frame records and the model window its associations fall into are
        // staged ONCE per level by the TMA unit (icp_tile.cuh) and serve all iterations of the level from shared memory; the window
        // is the associations' bounding box under the pose the level starts with, plus a margin for the pose to move; an
        // association that leaves it is served from global memory with identical arithmetic ----
        IcpTileView& tv = s_tv;                          // shared: the view costs no registers across the iteration loop
        const bool tile_res = p.resident[l] && p.icp && !L.icp.use_search;
        if (tile_res) {
            const IcpTileGeom& g = p.tile[l];
            float4* s_c0 = reinterpret_cast<float4*>(s_tile);
            float4* s_c1 = reinterpret_cast<float4*>(s_tile + icp_tile_curr_bytes(g));
            float4* s_g0 = reinterpret_cast<float4*>(s_tile + 2 * icp_tile_curr_bytes(g));
            float4* s_g1 = reinterpret_cast<float4*>(s_tile + 2 * icp_tile_curr_bytes(g) + icp_tile_model_bytes(g));
            float* s_gw = reinterpret_cast<float*>(s_tile + 2 * icp_tile_curr_bytes(g) + 2 * icp_tile_model_bytes(g));
            __syncthreads();      // nobody still reads the previous level's view or tile
            if (tid == 0) {
                const int tcx = (int)blockIdx.x % g.ncol, try_ = (int)blockIdx.x / g.ncol;
                tv.x0 = min(tcx * g.tw, L.icp.cols); tv.w = min(g.tw, L.icp.cols - tv.x0);
                tv.y0 = (int)(((long long)L.icp.rows * try_) / g.nrow); tv.h = (int)(((long long)L.icp.rows * (try_ + 1)) / g.nrow) - tv.y0;
                if (try_ >= g.nrow) tv.w = tv.h = 0;
                tv.c0 = s_c0; tv.c1 = s_c1; tv.g0 = s_g0; tv.g1 = s_g1; tv.gw = s_gw; tv.cbx = g.cbx; tv.cbs = g.cbs; tv.mbx = g.mbx; tv.mbs = g.mbs; tv.wbs = g.wbs;
                tv.mx0 = tv.my0 = tv.mwa = tv.mha = 0;
                s_box[0] = s_box[1] = 1 << 30; s_box[2] = s_box[3] = -1;
                if (tv.w > 0 && tv.h > 0) icp_tile_issue_curr(p.tmaps[l], g, s_c0, s_c1, tv.x0, tv.y0, &s_bar[0]);
            }
            __syncthreads();
            if (tv.w > 0 && tv.h > 0) {
                mbar_wait(&s_bar[0], par0); par0 ^= 1u;
                float Rc[9], tc[3], Rpi[9], tp[3];
#pragma unroll
                for (int k = 0; k < 9; ++k) { Rc[k] = S.Rcurr[k]; Rpi[k] = S.Rprev_inv[k]; }
#pragma unroll
                for (int k = 0; k < 3; ++k) { tc[k] = S.tcurr[k]; tp[k] = S.tprev[k]; }
                icp_tile_bbox<kTrackThreads>(L.icp, tv, Rc, tc, Rpi, tp, s_box);
                __syncthreads();
                if (tid == 0) {
                    int mx0, my0; bool any;
                    icp_tile_window(s_box, g.mw, g.mh, 2, mx0, my0, any);
                    if (any) {
                        tv.mx0 = mx0; tv.my0 = my0; tv.mwa = g.mw; tv.mha = g.mh;
                        icp_tile_issue_model(p.tmaps[l], g, L.icp.use_weight != 0, s_g0, s_g1, s_gw, mx0, my0, &s_bar[1]);
                    }
                }
                __syncthreads();
                if (tv.mwa > 0) { mbar_wait(&s_bar[1], par1); par1 ^= 1u; }
            }
        }
        for (int j = 0; j < L.iters; ++j) {
            if (S.done_level == l) break;
            const int next_level = (j + 1 < L.iters) ? l : next_lower;
            unsigned long long* part = p.ll_f + (size_t)(phase & 1) * gridDim.x * 64;
            const unsigned int tag = tag_base | (phase + 1);
            // ---- phase A: everything that only needs the current pose.  The photometric residuals go first: their integer
            // pair is published before the (longer) ICP pass, so the exchange is complete by the time anybody asks for it ----
            TP_STAMP(1);
            unsigned long long* ip = p.ll_i + (size_t)(iphase & 1) * gridDim.x * kIntStride;
            const unsigned int itag = tag_base | (iphase + 1);
            if (p.rgb) {
                // computeRgbResidual (reduce.cu:986-1060): the correspondence of each of this thread's pixels stays in its
                // shared-memory slot for the step pass below (the reference's corresImg round trip through HBM is gone)
                int cnt = 0, sig = 0;
                rgb_residual_pass<kTrackThreads>(L.res, L.cand, S.krkinv, begin, end, s_slots, cnt, sig);      // krkinv[9], kt[3] contiguous
                TP_STAMP(7);
                cnt = __reduce_add_sync(0xffffffffu, cnt);
                sig = __reduce_add_sync(0xffffffffu, sig);
                if (tid < 2) s_i[tid] = 0;
                __syncthreads();
                if ((tid & 31) == 0 && (cnt | sig)) { atomicAdd(&s_i[0], cnt); atomicAdd(&s_i[1], sig); }
                __syncthreads();
                if (tid < 2) ll_store(ip + (size_t)blockIdx.x * kIntStride + tid, (unsigned int)s_i[tid], itag);
                TP_STAMP(8);
            }
            if (p.icp) {
                float acc[32];
#pragma unroll
                for (int k = 0; k < 32; ++k) acc[k] = 0.f;
                float Rc[9], tc[3], Rpi[9], tp[3];
#pragma unroll
                for (int k = 0; k < 9; ++k) { Rc[k] = S.Rcurr[k]; Rpi[k] = S.Rprev_inv[k]; }
#pragma unroll
                for (int k = 0; k < 3; ++k) { tc[k] = S.tcurr[k]; tp[k] = S.tprev[k]; }
                if (L.icp.use_search) for (int i = gtid; i < N; i += gstride) icp_pixel<true>(L.icp, Rc, tc, Rpi, tp, i, acc);
                else if (tile_res) icp_tile_pass<kTrackThreads>(L.icp, tv, Rc, tc, Rpi, tp, acc);
                else icp_pass_nosearch<kTrackThreads>(L.icp, Rc, tc, Rpi, tp, begin, end, acc);
                TP_STAMP(2);
                block_partial32<kTrackThreads>(acc, s_w, part + (size_t)blockIdx.x * 64, tag);
                TP_STAMP(3);
            }
            if (p.rgb) {
                // ---- phase B: sigma of ALL residuals, then rgbStep (reduce.cu:718-811) from the slots ----
                all_reduce_int2(ip, itag, s_i);
                TP_STAMP(9);
                ++iphase;
                if (tid == 0) {      // RGBDOdometry.cpp:1017-1032
                    const int sigma = s_i[1], rgbSize = s_i[0];
                    float sigmaVal = sqrtf(((float)sigma / (float)rgbSize == 0.f) ? 1.f : (float)rgbSize);
                    const float rgbError = sqrtf((float)sigma) / (float)(rgbSize == 0 ? 1 : rgbSize);
                    const float lastErr = (j == 0) ? FLT_MAX : S.lastRGBError;
                    if (S.rgbOnly && rgbError > lastErr) {
                        S.done_level = l;
                        if (next_lower >= 0) {      // the next level warps with ITS camera matrix (RGBDOdometry.cpp:983-992)
                            double Rt[16];
#pragma unroll
                            for (int k = 0; k < 16; ++k) Rt[k] = S.resultRt[k];
                            update_krk(&S, Rt, next_lower);
                        }
                    } else {
                        S.lastRGBError = rgbError;
                        S.lastRGBCount = (float)rgbSize;
                        if (S.rgbOnly) sigmaVal = -1.f;
                        S.sigmaVal = sigmaVal;
                    }
                }
                __syncthreads();
                if (S.done_level == l) { ++phase; break; }      // identical decision in every CTA; this exchange's words are never read
                float acc[32];
#pragma unroll
                for (int k = 0; k < 32; ++k) acc[k] = 0.f;
                const float sigma = S.sigmaVal;
                TP_STAMP(10);
                rgb_step_pass<kTrackThreads>(L.step, sigma, s_slots, begin, end, ifx, ify, lcx, lcy, acc);
                TP_STAMP(11);
                block_partial32<kTrackThreads>(acc, s_w, part + (size_t)blockIdx.x * 64 + 32, tag);
            }
            TP_STAMP(4);
            all_reduce_partials<kTrackThreads>(part, tag, p.icp != 0, p.rgb != 0, s_d, s_tot);
            TP_STAMP(5);
            if (tid < 32) { if (p.icp) S.icp_sums[tid] = s_tot[tid]; if (p.rgb) S.rgb_sums[tid] = s_tot[32 + tid]; }
            __syncthreads();
            if (tid < 64) gn_update_warps(&S, next_level, p.dbg, &dbg_n);
            __syncthreads();
            TP_STAMP(6);
            ++phase;
        }
    }

    // ---- track_end_kernel + statistics ----
    if (blockIdx.x == 0 && tid == 0) {
        if (S.rgb) {
            const float dx = S.tcurr[0] - S.tprev[0], dy = S.tcurr[1] - S.tprev[1], dz = S.tcurr[2] - S.tprev[2];
            if ((double)sqrtf(dx * dx + dy * dy + dz * dz) > 0.3) {
                for (int k = 0; k < 9; ++k) S.Rcurr[k] = S.Rprev[k];
                for (int k = 0; k < 3; ++k) S.tcurr[k] = S.tprev[k];
            }
        }
        float cur[12], prev[12];
        for (int k = 0; k < 9; ++k) { cur[k] = S.Rcurr[k]; prev[k] = S.Rprev[k]; }
        for (int k = 0; k < 3; ++k) { cur[9 + k] = S.tcurr[k]; prev[9 + k] = S.tprev[k]; }
        for (int k = 0; k < 12; ++k) p.pose_out[k] = cur[k];
        if (p.last_pose_out) for (int k = 0; k < 12; ++k) p.last_pose_out[k] = prev[k];
        if (p.traj_out) for (int k = 0; k < 12; ++k) p.traj_out[k] = cur[k];
        if (p.inv_pose_out) pose_inverse_dev(cur, p.inv_pose_out);
        if (p.weighting_out) p.weighting_out[0] = velocity_weighting_dev(cur, prev, p.weight_multiplier);
        TrackState* g = p.st_global;
        g->lastICPError = S.lastICPError; g->lastICPCount = S.lastICPCount; g->lastRGBError = S.lastRGBError; g->lastRGBCount = S.lastRGBCount;
        g->lastSO3Error = S.lastSO3Error; g->lastSO3Count = S.lastSO3Count; g->icp_iterations_run = S.icp_iterations_run;
        for (int k = 0; k < 36; ++k) g->lastA[k] = S.lastA[k];
        for (int k = 0; k < 6; ++k) g->lastb[k] = S.lastb[k];
        for (int k = 0; k < 9; ++k) { g->Rcurr[k] = S.Rcurr[k]; g->Rprev[k] = S.Rprev[k]; g->Rprev_inv[k] = S.Rprev_inv[k]; }
        for (int k = 0; k < 3; ++k) { g->tcurr[k] = S.tcurr[k]; g->tprev[k] = S.tprev[k]; }
        g->icp = S.icp; g->rgb = S.rgb; g->rgbOnly = S.rgbOnly; g->so3 = S.so3; g->icpWeight = S.icpWeight; g->done_level = -1; g->sigmaVal = S.sigmaVal;
        for (int k = 0; k < 9; ++k) { g->krkinv[k] = S.krkinv[k]; g->so3_basis[k] = S.so3_basis[k]; g->so3_kinv[k] = S.so3_kinv[k]; g->so3_krlr[k] = S.so3_krlr[k]; }
        for (int k = 0; k < 3; ++k) g->kt[k] = S.kt[k];
    }
}

// ---- the parts of a tracking call that depend on the camera images alone, as kernels of their own: the frame pipeline runs them on
// its staging stream, one frame ahead of the tracker (hrbf_internal.h: CurrBank) ----

// Sobel of the next image + the pose-independent candidate mask of computeRgbResidual, all three levels (grid.y = level): the
// per-pixel code of the tracker's own prologue
struct SobelCandArgs { RgbResArgs r[3]; unsigned char* cand[3]; };
__global__ void __launch_bounds__(256) sobel_cand_kernel(const SobelCandArgs a)
{
    pdl_wait();
    const RgbResArgs& r = a.r[blockIdx.z];
    const int cols = r.cols, rows = r.rows;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= cols || y >= rows) return;
    const int k = y * cols + x;
    short* dIdx = const_cast<short*>(r.dIdx);
    short* dIdy = const_cast<short*>(r.dIdy);
    if (x >= 2 && x < cols - 1 && y >= 2 && y < rows - 1) {
        // interior: the 4 x 4 window of the zero test (rows y-2 .. y+1, columns x-2 .. x+1) contains the 3 x 3 Sobel window; the sums are
        // small integers, so the float sums of sobel_pixel are exact in any order and the kernel weights fold into two differences
        int p[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) p[u][v] = (int)__ldg(r.nextImage + (size_t)(y - 2 + u) * cols + (x - 2 + v));
        const int dx = (p[1][3] + 2 * p[2][3] + p[3][3]) - (p[1][1] + 2 * p[2][1] + p[3][1]);
        const int dy = (p[3][1] + 2 * p[3][2] + p[3][3]) - (p[1][1] + 2 * p[1][2] + p[1][3]);
        dIdx[k] = (short)dx; dIdy[k] = (short)dy;
        bool zero = false;
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) zero = zero || p[u][v] == 0;
        const short sx = (short)dx, sy = (short)dy;
        const bool cand = x < cols - 5 && !zero && (float)((sx * sx) + (sy * sy)) >= r.minScale && !isnan(__ldg(r.nextDepth + k));
        a.cand[blockIdx.z][k] = cand ? 1 : 0;
        return;
    }
    sobel_pixel(rows, cols, r.nextImage, dIdx, dIdy, x, y);
    a.cand[blockIdx.z][k] = rgb_static_candidate(r, k, dIdx[k], dIdy[k]) ? 1 : 0;      // plain loads of this thread's own stores
}

// The SO3 pre-alignment loop (RGBDOdometry.cpp:827-914) on the level-2 images of two consecutive camera frames: ONE CTA (4 800 pixels
// at 640x480; the loop is a chain of <= 10 dependent reductions, and it runs off the critical path), fp32 per-thread and per-warp
// sums, fp64 across warps, the same so3_pixel / so3_update as the tracker's in-kernel loop.
constexpr int kSo3Threads = 256;      // small enough to sit beside a 384-thread tracker (48 K + 16 K registers)
__global__ void __launch_bounds__(kSo3Threads, 4) so3_prealign_kernel(const unsigned char* __restrict__ lastImage, const unsigned char* __restrict__ nextImage,
                                                                   int rows, int cols, float fx, float fy, float cx, float cy, So3Pre* __restrict__ out)
{
    pdl_wait();
    __shared__ TrackState S;
    __shared__ float s_w[kSo3Threads / 32][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, N = rows * cols;
    if (tid == 0) {
        S.fx = fx; S.fy = fy; S.cx = cx; S.cy = cy;
        for (int k = 0; k < 9; ++k) { S.resultR[k] = S.lastResultR[k] = (k % 4 == 0) ? 1.0 : 0.0; S.R_lr[k] = (k % 4 == 0) ? 1.f : 0.f; }
        S.so3_lastError = FLT_MAX / 2; S.so3_lastCount = FLT_MAX / 2; S.so3_done = 0;
        S.lastSO3Error = 0; S.lastSO3Count = 0;
        const double I3[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
        update_so3_mats(&S, I3);
    }
    __syncthreads();
    for (int it = 0; it < 10; ++it) {
        if (S.so3_done) break;
        float acc[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) acc[k] = 0.f;
        for (int k = tid; k < N; k += kSo3Threads) so3_pixel(lastImage, nextImage, rows, cols, S.so3_basis, k, acc);
        s_w[warp][lane] = warp_reduce32_transpose(acc);
        __syncthreads();
        if (tid < 16) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < kSo3Threads / 32; ++w) t += (double)s_w[w][tid];
            S.so3_sums[tid] = t;
        }
        __syncthreads();
        if (tid == 0) so3_update(&S);
        __syncthreads();
    }
    if (tid < 9) out->resultR[tid] = S.resultR[tid];
    if (tid == 0) { out->lastSO3Error = S.lastSO3Error; out->lastSO3Count = S.lastSO3Count; }
}

}  // namespace hrbf
