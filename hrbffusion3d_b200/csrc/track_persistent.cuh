// track_persistent.cuh -- the whole coarse-to-fine tracking loop of RGBDOdometry::getIncrementalTransformation
// (Core/src/Utils/RGBDOdometry.cpp:796-1249) as ONE persistent cooperative kernel: one CTA per SM, grid-wide
// barriers between the reduction phases, the Gauss-Newton state in shared memory.
//
// Every CTA reduces the same per-CTA partial sums in the same fixed order and runs the same fp64 solve, so all
// CTAs hold bit-identical copies of the state and only ONE grid barrier is needed per reduction (none to
// broadcast the new pose).  Compared with one kernel per reduction (icp_reduce_kernel & co., kept for the
// step-function ABI) this removes ~70 dependent kernel boundaries per frame, the ticket atomics and the
// __threadfence of the last-block-done scheme.
#pragma once
#include "odometry_kernels.cuh"

namespace hrbf {

constexpr int kTrackThreads = 512;       // x 148 CTAs; 128 registers per thread
constexpr int kTrackWarps = kTrackThreads / 32;

struct TrackLevel {
    IcpArgs icp;
    RgbResArgs res;
    RgbStepArgs step;
    float* cloud;            // projectToPointCloud target (cudafuncs.cu:995-1013)
    int iters;
};
struct TrackParams {
    TrackLevel lvl[3];
    const unsigned char *so3_last, *so3_next;      // level-2 images of the SO3 pre-alignment
    int icp, rgb, rgbOnly, so3;
    float icpWeight;
    const float* prev_pose;                        // device R[9], t[3]
    float* pose_out;                               // device R[9], t[3]
    TrackState* st_global;                         // camera in, statistics out
    float* partials;                               // [2][gridDim.x][64]
    int* ipartials;                                // [2][gridDim.x][2]
    unsigned int* barrier;                         // zeroed before the launch
    long long* dbg;                                // optional: %globaltimer stamps of CTA 0 (profiling builds)
};

// Monotonic-counter grid barrier (all CTAs are co-resident: cooperative launch).  Release: the CTA's global
// writes are ordered before the arrive by __syncthreads + __threadfence; acquire: readers use ld.cg after it.
__device__ __forceinline__ void grid_sync(unsigned int* bar, unsigned int& target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(bar, 1u);
        unsigned int v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory"); } while (v < target);
    }
    __syncthreads();
}

// 32 per-thread sums -> this CTA's partial (written by warp 0 to dst[0..31])
__device__ __forceinline__ void block_partial32(float (&acc)[32], float (*s_w)[32], float* dst)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float r = warp_reduce32_transpose(acc);
    s_w[warp][lane] = r;
    __syncthreads();
    if (threadIdx.x < 32) {
        float p = 0.f;
#pragma unroll
        for (int w = 0; w < kTrackWarps; ++w) p += s_w[w][threadIdx.x];
        dst[threadIdx.x] = p;
    }
    __syncthreads();
}

// every CTA: grid totals of the 2 x 32 partial sums, fp64, fixed order -> out[64] (shared).
// Thread t owns value (t & 63) of CTA slice (t >> 6): its <= 8-deep batches of loads are all independent (one L2
// round trip per batch instead of one per CTA), the additions run in a fixed order -> identical in every CTA.
__device__ __forceinline__ void all_reduce_partials(const float* part /* [gridDim.x][64] */, double (*s_d)[64], double* out /* [64] */)
{
    constexpr int kSlices = kTrackThreads / 64, U = 8;
    static_assert(kTrackThreads % 64 == 0, "slice layout");
    const int v = threadIdx.x & 63, q = threadIdx.x >> 6;
    double a = 0.0;
    unsigned int b = q;
    for (; b + (U - 1) * kSlices < gridDim.x; b += U * kSlices) {
        float t[U];
#pragma unroll
        for (int u = 0; u < U; ++u) t[u] = __ldcg(part + (size_t)(b + u * kSlices) * 64 + v);
#pragma unroll
        for (int u = 0; u < U; ++u) a += (double)t[u];
    }
    {
        float t[U];
#pragma unroll
        for (int u = 0; u < U; ++u) t[u] = (b + u * kSlices < gridDim.x) ? __ldcg(part + (size_t)(b + u * kSlices) * 64 + v) : 0.f;
#pragma unroll
        for (int u = 0; u < U; ++u) a += (double)t[u];
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

// ---- the no-search ICP pixel split into load / gather / finish stages so that two pixels per thread are in flight
// (the pass is bound by two dependent L2 round trips per pixel, not by bandwidth or issue slots) ----
struct IcpCurr { float vx, vy, vz, nx, ny, nz, k1, k2; };
struct IcpModel { float vx, vy, vz, nx, ny, nz, k1, k2, w; int ok, ux, uy; float3 vg, ng; };

__device__ __forceinline__ IcpCurr icp_load_curr(const IcpArgs& a, int i)
{
    const int y = i / a.cols, x = i - y * a.cols, rows = a.rows;
    auto ld = [&](const float* p, int plane) { return __ldg(p + (size_t)(plane * rows + y) * a.cpitch + x); };
    IcpCurr c;
    c.vx = ld(a.vc, 0); c.vy = ld(a.vc, 1); c.vz = ld(a.vc, 2);
    c.nx = ld(a.nc, 0); c.ny = ld(a.nc, 1); c.nz = ld(a.nc, 2);
    c.k1 = ld(a.k1c, 3); c.k2 = ld(a.k2c, 3);
    return c;
}
__device__ __forceinline__ IcpModel icp_gather_model(const IcpArgs& a, const IcpCurr& c, const float* Rc, const float* tc, const float* Rpi, const float* tp)
{
    IcpModel m;
    const int rows = a.rows;
    m.vg = mul(Rc, make_float3(c.vx, c.vy, c.vz)) + make_float3(tc[0], tc[1], tc[2]);
    const float3 vcp = mul(Rpi, m.vg - make_float3(tp[0], tp[1], tp[2]));
    m.ux = __float2int_rn(vcp.x * a.fx / vcp.z + a.cx);
    m.uy = __float2int_rn(vcp.y * a.fy / vcp.z + a.cy);
    m.ok = !(m.ux < 0 || m.uy < 0 || m.ux >= a.cols || m.uy >= rows || vcp.z < 0) && !(isnan(c.vx) || isnan(c.nx) || isnan(c.k1) || isnan(c.k2));
    m.ng = mul(Rc, make_float3(c.nx, c.ny, c.nz));
    m.vx = m.vy = m.vz = m.nx = m.ny = m.nz = m.k1 = m.k2 = m.w = 0.f;
    if (m.ok) {
        auto ld = [&](const float* p, int plane) { return __ldg(p + (size_t)(plane * rows + m.uy) * a.gpitch + m.ux); };
        m.vx = ld(a.vg, 0); m.vy = ld(a.vg, 1); m.vz = ld(a.vg, 2);
        m.nx = ld(a.ng, 0); m.ny = ld(a.ng, 1); m.nz = ld(a.ng, 2);
        m.k1 = ld(a.k1g, 3); m.k2 = ld(a.k2g, 3);
        m.w = a.use_weight ? __ldg(a.w + (size_t)m.uy * a.wpitch + m.ux) : 1.f;
    }
    return m;
}
__device__ __forceinline__ void icp_finish(const IcpArgs& a, const IcpModel& m, const float* Rpi, const float* tp, int i, float (&acc)[32])
{
    float row[7] = { 0, 0, 0, 0, 0, 0, 0 };
    float weight = 1.f;
    bool found = false;
    if (m.ok) {
        const float3 vp = make_float3(m.vx, m.vy, m.vz), np = make_float3(m.nx, m.ny, m.nz);
        const float dist = norm(vp - m.vg), sine = norm(cross(m.ng, np));
        found = !(isnan(vp.x) || isnan(np.x) || isnan(m.k1) || isnan(m.k2)) && !(sine > a.angle_thres || dist > a.dist_thres);
        if (found) {
            const float3 tpv = make_float3(tp[0], tp[1], tp[2]);
            const float3 s_cp = mul(Rpi, m.vg - tpv), d_cp = mul(Rpi, vp - tpv), n_cp = mul(Rpi, np);
            if (a.use_weight) weight = isnan(m.w) ? 0.f : m.w;
            const float3 c = cross(s_cp, n_cp);
            row[0] = n_cp.x; row[1] = n_cp.y; row[2] = n_cp.z; row[3] = c.x; row[4] = c.y; row[5] = c.z;
            row[6] = dot(n_cp, s_cp - d_cp);
        }
    }
    if (a.corres) a.corres[i] = found ? make_int2(m.ux, m.uy) : make_int2(-1, -1);
    accumulate_row7(acc, row, weight, found);
}
// this CTA's contiguous pixel range [begin, end), two pixels in flight per thread
__device__ __forceinline__ void icp_pass_nosearch(const IcpArgs& a, const float* Rc, const float* tc, const float* Rpi, const float* tp,
                                                  int begin, int end, float (&acc)[32])
{
    for (int i = begin + (int)threadIdx.x; i < end; i += 2 * kTrackThreads) {
        const int j = i + kTrackThreads;
        const bool two = j < end;
        const IcpCurr c0 = icp_load_curr(a, i), c1 = icp_load_curr(a, two ? j : i);
        const IcpModel m0 = icp_gather_model(a, c0, Rc, tc, Rpi, tp), m1 = icp_gather_model(a, c1, Rc, tc, Rpi, tp);
        icp_finish(a, m0, Rpi, tp, i, acc);
        if (two) icp_finish(a, m1, Rpi, tp, j, acc);
    }
}

__device__ __forceinline__ long long gtimer() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define TP_STAMP(slot) do { if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) p.dbg[dbg_n < 500 ? dbg_n++ : 499] = ((long long)(slot) << 56) | (gtimer() & 0x00ffffffffffffffll); } while (0)

__global__ void __launch_bounds__(kTrackThreads, 1) track_persistent_kernel(const TrackParams p)
{
    int dbg_n = 0;
    __shared__ TrackState S;
    __shared__ float s_w[kTrackWarps][32];
    __shared__ double s_d[kTrackThreads / 64][64];
    __shared__ double s_tot[64];
    __shared__ int s_i[2];
    unsigned int bar_target = 0;
    unsigned int phase = 0;
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
        if (p.so3) update_so3_mats(&S, I3);
        else if (p.rgb) update_krk(&S, I4, first_level);
    }

    // ---- RGB branch prep: Sobel of the next image and the back-projected last depth, all levels ----
    if (p.rgb) {
        for (int l = 0; l < 3; ++l) {
            const RgbResArgs& r = p.lvl[l].res;
            const int N = r.rows * r.cols;
            const int div = 1 << l;
            const float ifx = 1.0f / (p.st_global->fx / div), ify = 1.0f / (p.st_global->fy / div), cx = p.st_global->cx / div, cy = p.st_global->cy / div;
            for (int k = gtid; k < N; k += gstride) {
                const int y = k / r.cols, x = k - y * r.cols;
                sobel_pixel(r.rows, r.cols, r.nextImage, const_cast<short*>(r.dIdx), const_cast<short*>(r.dIdy), x, y);
                project_pixel(r.rows, r.cols, r.lastDepth, p.lvl[l].cloud, ifx, ify, cx, cy, x, y);
            }
        }
        grid_sync(p.barrier, bar_target);       // rgb_step reads the cloud at OTHER pixels
    } else {
        __syncthreads();
    }

    // ---- SO3 pre-alignment (RGBDOdometry.cpp:827-914), level 2 ----
    if (p.so3) {
        const int rows = p.lvl[2].res.rows, cols = p.lvl[2].res.cols, N = rows * cols;
        for (int it = 0; it < 10; ++it) {
            if (S.so3_done) break;
            float acc[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) acc[k] = 0.f;
            // so3_pixel reads s_m = basis[9], kinv[9], krlr[9]: contiguous in TrackState
            for (int k = gtid; k < N; k += gstride) so3_pixel(p.so3_last, p.so3_next, rows, cols, S.so3_basis, k, acc);
            float* part = p.partials + (size_t)(phase & 1) * gridDim.x * 64;
            block_partial32(acc, s_w, part + (size_t)blockIdx.x * 64);
            grid_sync(p.barrier, bar_target);
            all_reduce_partials(part, s_d, s_tot);
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
        for (int j = 0; j < L.iters; ++j) {
            if (S.done_level == l) break;
            const int next_level = (j + 1 < L.iters) ? l : next_lower;
            if (p.rgb) {
                // computeRgbResidual: correspondences + {count, sum diff^2}
                int cnt = 0, sig = 0;
                for (int k = gtid; k < N; k += gstride) rgb_residual_pixel(L.res, S.krkinv, k, cnt, sig);      // krkinv[9], kt[3] contiguous
                cnt = __reduce_add_sync(0xffffffffu, cnt);
                sig = __reduce_add_sync(0xffffffffu, sig);
                if (tid < 2) s_i[tid] = 0;
                __syncthreads();
                if ((tid & 31) == 0 && (cnt | sig)) { atomicAdd(&s_i[0], cnt); atomicAdd(&s_i[1], sig); }
                __syncthreads();
                int* ip = p.ipartials + (size_t)(phase & 1) * gridDim.x * 2;
                if (tid < 2) ip[blockIdx.x * 2 + tid] = s_i[tid];
                grid_sync(p.barrier, bar_target);
                if (tid < 2) s_i[tid] = 0;
                __syncthreads();
                int c2 = 0, g2 = 0;
                for (unsigned int b = tid; b < gridDim.x; b += blockDim.x) { c2 += __ldcg(ip + b * 2); g2 += __ldcg(ip + b * 2 + 1); }
                c2 = __reduce_add_sync(0xffffffffu, c2);
                g2 = __reduce_add_sync(0xffffffffu, g2);
                if ((tid & 31) == 0 && (c2 | g2)) { atomicAdd(&s_i[0], c2); atomicAdd(&s_i[1], g2); }
                __syncthreads();
                if (tid == 0) {      // RGBDOdometry.cpp:1017-1032
                    const int sigma = s_i[1], rgbSize = s_i[0];
                    float sigmaVal = sqrtf(((float)sigma / (float)rgbSize == 0.f) ? 1.f : (float)rgbSize);
                    const float rgbError = sqrtf((float)sigma) / (float)(rgbSize == 0 ? 1 : rgbSize);
                    const float lastErr = (j == 0) ? FLT_MAX : S.lastRGBError;
                    if (S.rgbOnly && rgbError > lastErr) S.done_level = l;
                    else {
                        S.lastRGBError = rgbError;
                        S.lastRGBCount = (float)rgbSize;
                        if (S.rgbOnly) sigmaVal = -1.f;
                        S.sigmaVal = sigmaVal;
                    }
                }
                __syncthreads();
                ++phase;
                if (S.done_level == l) break;
            }
            float* part = p.partials + (size_t)(phase & 1) * gridDim.x * 64;
            TP_STAMP(1);
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
                else {
                    const int begin = (int)(((long long)N * blockIdx.x) / gridDim.x), end = (int)(((long long)N * (blockIdx.x + 1)) / gridDim.x);
                    icp_pass_nosearch(L.icp, Rc, tc, Rpi, tp, begin, end, acc);
                }
                TP_STAMP(2);
                block_partial32(acc, s_w, part + (size_t)blockIdx.x * 64);
                TP_STAMP(3);
            }
            if (p.rgb) {
                float acc[32];
#pragma unroll
                for (int k = 0; k < 32; ++k) acc[k] = 0.f;
                const float sigma = S.sigmaVal;
                for (int i = gtid; i < N; i += gstride) rgb_step_pixel(L.step, sigma, i, acc);
                block_partial32(acc, s_w, part + (size_t)blockIdx.x * 64 + 32);
            }
            grid_sync(p.barrier, bar_target);
            TP_STAMP(4);
            all_reduce_partials(part, s_d, s_tot);
            TP_STAMP(5);
            if (tid < 32) { if (p.icp) S.icp_sums[tid] = s_tot[tid]; if (p.rgb) S.rgb_sums[tid] = s_tot[32 + tid]; }
            __syncthreads();
            if (tid == 0) gn_update(&S, l, next_level);
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
        for (int k = 0; k < 9; ++k) p.pose_out[k] = S.Rcurr[k];
        for (int k = 0; k < 3; ++k) p.pose_out[9 + k] = S.tcurr[k];
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

}  // namespace hrbf
