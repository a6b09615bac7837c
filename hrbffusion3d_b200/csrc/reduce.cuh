// reduce.cuh -- single-pass warp-shuffle tree reductions and the device-side
// Gauss-Newton state shared by the ICP / RGB / SO3 kernels.
//
// Replaces the reference's two-kernel scheme (per-block partials + reduceSum<<<1,1024>>>,
// Core/src/Cuda/reduce.cu:91-251) and its host-side solve (Utils/RGBDOdometry.cpp:1162-1204).
#pragma once
#include "common.cuh"

namespace hrbf {

// Device-resident tracking state (one per hrbf_odometry / per reduce workspace).
struct TrackState {
    // previous pose and current estimate (RGBDOdometry.cpp:810-814, 920-922)
    float Rprev[9], tprev[3], Rprev_inv[9];
    float Rcurr[9], tcurr[3];
    double resultRt[16];                 // RGBDOdometry.cpp:924
    // per-iteration photometric warp (RGBDOdometry.cpp:983-992)
    float krkinv[9], kt[3];
    // SO3 pre-alignment (RGBDOdometry.cpp:825-914)
    double resultR[9], lastResultR[9];
    float R_lr[9];
    float so3_basis[9], so3_kinv[9], so3_krlr[9];
    float so3_lastError, so3_lastCount;
    int so3_done;
    // camera (level 0) and options
    float fx, fy, cx, cy;
    float icpWeight;
    int icp, rgb, rgbOnly, so3;
    int done_level;                      // level at which rgbOnly broke out early (RGBDOdometry.cpp:1020-1023), -1 = none
    // per-iteration reduction results
    int rgb_count, rgb_sigma;            // int2 of computeRgbResidual
    float sigmaVal;
    double icp_sums[32], rgb_sums[32], so3_sums[16];
    // statistics mirrored to the host (RGBDOdometry.h:124-134)
    float lastICPError, lastICPCount, lastRGBError, lastRGBCount, lastSO3Error, lastSO3Count;
    double lastA[36], lastb[6];
    double gn_t[3], gn_R[9];             // scratch of the warp-parallel update (translation and rotation increment)
    int icp_iterations_run;
    unsigned int ticket;                 // last-block-done counter
    int pad_;
};

struct ReduceWork {
    TrackState st;
    float partials[kMaxReduceBlocks][32];
};

// Transposing butterfly: 32 per-lane values -> lane l ends with the warp total of value l.
// 31 shuffles instead of 32*5.
__device__ __forceinline__ float warp_reduce32_transpose(float (&v)[32])
{
    const unsigned lane = threadIdx.x & 31u;
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool upper = (lane & s) != 0;
#pragma unroll
        for (int j = 0; j < s; ++j) {
            const float send = upper ? v[j] : v[j + s];
            const float keep = upper ? v[j + s] : v[j];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return v[0];
}

// Block (kReduceThreads) -> grid reduction of 32 floats.  Returns true in EVERY thread of the
// block that finished last; `total` (shared, double[32]) then holds the grid totals, summed in a
// fixed order (independent of block completion order -> deterministic).
__device__ __forceinline__ bool grid_reduce32(float (&v)[32], float (*partials)[32], unsigned int* ticket,
                                              double* total /* shared double[32] */)
{
    __shared__ float s_w[kReduceThreads / 32][32];
    __shared__ double s_d[kReduceThreads / 32][32];
    __shared__ unsigned int s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    const float r = warp_reduce32_transpose(v);
    s_w[warp][lane] = r;
    __syncthreads();
    if (threadIdx.x < 32) {
        float p = 0.f;
#pragma unroll
        for (int w = 0; w < kReduceThreads / 32; ++w) p += s_w[w][threadIdx.x];
        partials[blockIdx.x][threadIdx.x] = p;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(ticket, 1u);
        s_last = (t == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
    // final: 8 slices x 32 values, fp64, fixed order
    {
        // each warp sums every 8th partial; ALL of a warp's loads are issued together (one L2 round trip for up to
        // 8 x 40 = 320 blocks), the additions stay in a fixed order
        double acc = 0.0;
        constexpr int W = kReduceThreads / 32, U = 40;
        for (unsigned int b0 = warp; b0 < gridDim.x; b0 += U * W) {
            float t[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { const unsigned int b = b0 + u * W; t[u] = b < gridDim.x ? __ldcg(&partials[b][lane]) : 0.f; }
#pragma unroll
            for (int u = 0; u < U; ++u) acc += (double)t[u];
        }
        s_d[warp][lane] = acc;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        double acc = 0.0;
#pragma unroll
        for (int w = 0; w < kReduceThreads / 32; ++w) acc += s_d[w][threadIdx.x];
        total[threadIdx.x] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) *ticket = 0u;   // re-arm for the next launch
    return true;
}

}  // namespace hrbf
