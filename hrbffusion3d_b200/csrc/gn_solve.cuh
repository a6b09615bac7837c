// gn_solve.cuh -- device-side Gauss-Newton bookkeeping: fp64 6x6 / 3x3 LDLT, Rodrigues, SE3
// composition.  Runs in ONE thread of the last block of a reduction kernel, so the whole
// coarse-to-fine loop never leaves the GPU.  Everything on the fast path is fully unrolled with
// compile-time indices so the small matrices live in registers (no local-memory traffic).
//
// Restates the host code of Utils/RGBDOdometry.cpp:825-914 (SO3), :983-992 (K R K^-1, K t),
// :1162-1204 (normal equations, solve, pose update) and Utils/OdometryProvider.h:35-93.
#pragma once
#include "reduce.cuh"
#include <float.h>

namespace hrbf {

// (a0 b0 + a1 b1) + a2 b2 with every product and sum rounded separately: the float pose compositions of the host code
// (RGBDOdometry.cpp:902, 1196-1204) are plain C++ / Eigen fixed-size products, no FMA
__device__ __forceinline__ float dot3_rn(float a0, float b0, float a1, float b1, float a2, float b2)
{
    return __fadd_rn(__fadd_rn(__fmul_rn(a0, b0), __fmul_rn(a1, b1)), __fmul_rn(a2, b2));
}

// Slow path: LDL^T with symmetric diagonal pivoting; a vanishing pivot contributes 0 (what
// Eigen's ldlt().solve() does for a singular system, e.g. no correspondences at all).
template <int N>
__device__ __noinline__ void ldlt_solve_pivoted(const double* Ain, const double* bin, double* x)
{
    double A[N * N], y[N];
    int perm[N];
    for (int i = 0; i < N * N; ++i) A[i] = Ain[i];
    for (int i = 0; i < N; ++i) perm[i] = i;
    for (int k = 0; k < N; ++k) {
        int p = k;
        double big = fabs(A[k * N + k]);
        for (int i = k + 1; i < N; ++i)
            if (fabs(A[i * N + i]) > big) { big = fabs(A[i * N + i]); p = i; }
        if (p != k) {
            for (int j = 0; j < N; ++j) { double t = A[k * N + j]; A[k * N + j] = A[p * N + j]; A[p * N + j] = t; }
            for (int j = 0; j < N; ++j) { double t = A[j * N + k]; A[j * N + k] = A[j * N + p]; A[j * N + p] = t; }
            int t = perm[k]; perm[k] = perm[p]; perm[p] = t;
        }
        const double d = A[k * N + k];
        if (fabs(d) <= DBL_MIN) continue;
        for (int i = k + 1; i < N; ++i) A[i * N + k] /= d;
        for (int i = k + 1; i < N; ++i)
            for (int j = k + 1; j <= i; ++j) {
                A[i * N + j] -= A[i * N + k] * d * A[j * N + k];
                A[j * N + i] = A[i * N + j];
            }
    }
    for (int i = 0; i < N; ++i) y[i] = bin[perm[i]];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < i; ++j) y[i] -= A[i * N + j] * y[j];
    for (int i = 0; i < N; ++i) { const double d = A[i * N + i]; y[i] = (fabs(d) > DBL_MIN) ? y[i] / d : 0.0; }
    for (int i = N - 1; i >= 0; --i)
        for (int j = i + 1; j < N; ++j) y[i] -= A[j * N + i] * y[j];
    for (int i = 0; i < N; ++i) x[perm[i]] = y[i];
}

// 1/d for a pivot of the normal matrix (|d| within float range): float seed + two Newton steps in fp64 (~1e-15 relative),
// a fraction of the instruction count of the IEEE fp64 division on the single thread that runs the solve
__device__ __forceinline__ double rcp_pivot(double d)
{
    double r = (double)(1.0f / (float)d);
    r = r * (2.0 - d * r);
    r = r * (2.0 - d * r);
    return r;
}

// Fast path: unpivoted LDL^T in registers (the normal matrix is SPD whenever tracking has
// support).  Returns false if a pivot is not safely positive -> caller takes the pivoted path.
template <int N>
__device__ __forceinline__ bool ldlt_solve_spd(const double (&A)[N * N], const double (&b)[N], double (&x)[N])
{
    double L[N][N], D[N], Dinv[N], y[N];
    double amax = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) amax = fmax(amax, fabs(A[i * N + i]));
    const double tiny = amax * 1e-13;
    bool ok = amax > 1e-30 && amax < 1e30;               // rcp_pivot seeds in float
#pragma unroll
    for (int j = 0; j < N; ++j) {
        double d = A[j * N + j];
#pragma unroll
        for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k] * D[k];
        D[j] = d;
        ok = ok && (d > tiny);
        const double inv = rcp_pivot(d);
        Dinv[j] = inv;
#pragma unroll
        for (int i = j + 1; i < N; ++i) {
            double s = A[i * N + j];
#pragma unroll
            for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k] * D[k];
            L[i][j] = s * inv;
        }
    }
    if (!ok) return false;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double s = b[i];
#pragma unroll
        for (int k = 0; k < i; ++k) s -= L[i][k] * y[k];
        y[i] = s;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) y[i] *= Dinv[i];
#pragma unroll
    for (int i = N - 1; i >= 0; --i) {
        double s = y[i];
#pragma unroll
        for (int k = i + 1; k < N; ++k) s -= L[k][i] * x[k];
        x[i] = s;
    }
    return true;
}

template <int N>
__device__ __forceinline__ void ldlt_solve(const double (&A)[N * N], const double (&b)[N], double (&x)[N])
{
    if (!ldlt_solve_spd<N>(A, b, x)) {
        // copies keep the caller's arrays in registers (their address never escapes)
        double Ac[N * N], bc[N], xc[N];
#pragma unroll
        for (int i = 0; i < N * N; ++i) Ac[i] = A[i];
#pragma unroll
        for (int i = 0; i < N; ++i) bc[i] = b[i];
        ldlt_solve_pivoted<N>(Ac, bc, xc);
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = xc[i];
    }
}

// OdometryProvider.h:35-69
__device__ __forceinline__ void rodrigues(const double (&w)[3], double (&R)[9])
{
    double rx = w[0], ry = w[1], rz = w[2];
    const double t2 = rx * rx + ry * ry + rz * rz;
    if (t2 < 0.0625) {
        // Gauss-Newton increments are tiny: R = cos(t) I + A [w]x + B w w^T with A = sin(t)/t and B = (1 - cos t)/t^2 as
        // series in t^2 (truncation < 1e-19 for t < 0.25): no square root, division or library sincos on the critical thread
        const double A = 1.0 + t2 * (-1.0 / 6 + t2 * (1.0 / 120 + t2 * (-1.0 / 5040 + t2 * (1.0 / 362880 + t2 * (-1.0 / 39916800 + t2 * (1.0 / 6227020800.0))))));
        const double B = 0.5 + t2 * (-1.0 / 24 + t2 * (1.0 / 720 + t2 * (-1.0 / 40320 + t2 * (1.0 / 3628800 + t2 * (-1.0 / 479001600 + t2 * (1.0 / 87178291200.0))))));
        const double c = 1.0 - B * t2;
        R[0] = c + B * rx * rx; R[1] = B * rx * ry - A * rz; R[2] = B * rx * rz + A * ry;
        R[3] = B * rx * ry + A * rz; R[4] = c + B * ry * ry; R[5] = B * ry * rz - A * rx;
        R[6] = B * rx * rz - A * ry; R[7] = B * ry * rz + A * rx; R[8] = c + B * rz * rz;
        return;
    }
    const double theta = sqrt(t2);
    R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
    if (theta >= DBL_EPSILON) {
        double s, c;
        sincos(theta, &s, &c);
        const double c1 = 1. - c, it = 1. / theta;
        rx *= it; ry *= it; rz *= it;
        R[0] = c + c1 * rx * rx; R[1] = c1 * rx * ry - s * rz; R[2] = c1 * rx * rz + s * ry;
        R[3] = c1 * rx * ry + s * rz; R[4] = c + c1 * ry * ry; R[5] = c1 * ry * rz - s * rx;
        R[6] = c1 * rx * rz - s * ry; R[7] = c1 * ry * rz + s * rx; R[8] = c + c1 * rz * rz;
    }
}

__device__ __forceinline__ void inv3(const double (&m)[9], double (&o)[9])
{
    const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
    const double det = m[0] * c00 + m[1] * c01 + m[2] * c02, id = 1.0 / det;
    o[0] = c00 * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
    o[3] = c01 * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
    o[6] = c02 * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}
__device__ __forceinline__ void inv3f(const float* m, float* o)
{
    // Rprev.inverse() on the host (RGBDOdometry.cpp:920): cofactors / determinant in float, every operation rounded separately
    auto d2 = [](float a, float b, float c, float d) { return __fsub_rn(__fmul_rn(a, b), __fmul_rn(c, d)); };
    const float c00 = d2(m[4], m[8], m[5], m[7]), c01 = d2(m[5], m[6], m[3], m[8]), c02 = d2(m[3], m[7], m[4], m[6]);
    const float det = dot3_rn(m[0], c00, m[1], c01, m[2], c02), id = __fdiv_rn(1.0f, det);
    o[0] = __fmul_rn(c00, id); o[1] = __fmul_rn(d2(m[2], m[7], m[1], m[8]), id); o[2] = __fmul_rn(d2(m[1], m[5], m[2], m[4]), id);
    o[3] = __fmul_rn(c01, id); o[4] = __fmul_rn(d2(m[0], m[8], m[2], m[6]), id); o[5] = __fmul_rn(d2(m[2], m[3], m[0], m[5]), id);
    o[6] = __fmul_rn(c02, id); o[7] = __fmul_rn(d2(m[1], m[6], m[0], m[7]), id); o[8] = __fmul_rn(d2(m[0], m[4], m[1], m[3]), id);
}
__device__ __forceinline__ void mul3(const double (&a)[9], const double (&b)[9], double (&o)[9])
{
    double r[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) r[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
#pragma unroll
    for (int k = 0; k < 9; ++k) o[k] = r[k];
}

// K of pyramid level `level` (Cuda/types.cuh:93-97: float division by 2^level)
__device__ __forceinline__ void level_K(const TrackState* st, int level, double (&K)[9])
{
    const int div = 1 << level;
    const float fx = st->fx / div, fy = st->fy / div, cx = st->cx / div, cy = st->cy / div;
    K[0] = fx; K[1] = 0; K[2] = cx; K[3] = 0; K[4] = fy; K[5] = cy; K[6] = 0; K[7] = 0; K[8] = 1;
}

// RGBDOdometry.cpp:983-992 : Rt = resultRt^-1, K R K^-1 and K t for the photometric warp
__device__ __forceinline__ void update_krk(TrackState* st, const double (&Rt)[16], int level)
{
    double K[9], Kinv[9], Rm[9], Rinv[9], tm[3], KR[9], KRK[9];
    level_K(st, level, K);
    inv3(K, Kinv);
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) Rm[a * 3 + b] = Rt[a * 4 + b];
    inv3(Rm, Rinv);
#pragma unroll
    for (int a = 0; a < 3; ++a) tm[a] = -(Rinv[a * 3] * Rt[3] + Rinv[a * 3 + 1] * Rt[7] + Rinv[a * 3 + 2] * Rt[11]);
    mul3(K, Rinv, KR);
    mul3(KR, Kinv, KRK);
#pragma unroll
    for (int k = 0; k < 9; ++k) st->krkinv[k] = (float)KRK[k];
#pragma unroll
    for (int a = 0; a < 3; ++a) st->kt[a] = (float)(K[a * 3] * tm[0] + K[a * 3 + 1] * tm[1] + K[a * 3 + 2] * tm[2]);
}

// RGBDOdometry.cpp:851-862 : homography K R K^-1, K^-1, K R for the SO3 step (level 2)
__device__ __forceinline__ void update_so3_mats(TrackState* st, const double (&resultR)[9])
{
    double K[9], Kinv[9], KR[9], H[9];
    level_K(st, 2, K);
    inv3(K, Kinv);
    mul3(K, resultR, KR);
    mul3(KR, Kinv, H);
#pragma unroll
    for (int k = 0; k < 9; ++k) { st->so3_basis[k] = (float)H[k]; st->so3_kinv[k] = (float)Kinv[k]; st->so3_krlr[k] = (float)KR[k]; }
}

// 27 upper-triangular sums -> symmetric A (6x6) and b, rounded through float like the
// reference's download into Eigen float matrices (reduce.cu:677-689, RGBDOdometry.cpp:1163-1166)
__device__ __forceinline__ void unpack_se3(const double* s, double (&A)[36], double (&b)[6])
{
    int shift = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = i; j < 7; ++j) {
            const double value = (double)(float)s[shift++];
            if (j == 6) b[i] = value; else { A[j * 6 + i] = value; A[i * 6 + j] = value; }
        }
}

// One Gauss-Newton update (RGBDOdometry.cpp:1135-1204) from the reduced sums held in the state.
// next_level: pyramid level of the NEXT iteration (-1: none).
__device__ __forceinline__ void gn_update(TrackState* st, int cur_level, int next_level)
{
    (void)cur_level;
    double lastA[36], lastb[6], result[6];
    const bool icp = st->icp != 0, rgb = st->rgb != 0;
    if (icp) {
        const float r0 = (float)st->icp_sums[27], r1 = (float)st->icp_sums[28];
        st->lastICPError = sqrtf(r0) / r1;
        st->lastICPCount = r1;
        st->icp_iterations_run++;
    }
    if (icp && rgb) {
        double Ai[36], bi[6];
        unpack_se3(st->rgb_sums, lastA, lastb);
        unpack_se3(st->icp_sums, Ai, bi);
        const double w = st->icpWeight;
#pragma unroll
        for (int k = 0; k < 36; ++k) lastA[k] = lastA[k] + w * w * Ai[k];
#pragma unroll
        for (int k = 0; k < 6; ++k) lastb[k] = lastb[k] + w * bi[k];
    } else if (icp) {
        unpack_se3(st->icp_sums, lastA, lastb);
    } else {
        unpack_se3(st->rgb_sums, lastA, lastb);
    }
    ldlt_solve<6>(lastA, lastb, result);
#pragma unroll
    for (int k = 0; k < 36; ++k) st->lastA[k] = lastA[k];
#pragma unroll
    for (int k = 0; k < 6; ++k) st->lastb[k] = lastb[k];

    // OdometryProvider.h:71-93 : resultRt = [exp(w) | t] * resultRt
    double Rupd[9], old[16], nrt[16];
    const double wv[3] = { result[3], result[4], result[5] };
    rodrigues(wv, Rupd);
#pragma unroll
    for (int k = 0; k < 16; ++k) old[k] = st->resultRt[k];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
            nrt[a * 4 + b] = Rupd[a * 3 + 0] * old[0 * 4 + b] + Rupd[a * 3 + 1] * old[1 * 4 + b] + Rupd[a * 3 + 2] * old[2 * 4 + b] + result[a] * old[3 * 4 + b];
#pragma unroll
    for (int b = 0; b < 4; ++b) nrt[12 + b] = old[12 + b];
#pragma unroll
    for (int k = 0; k < 16; ++k) st->resultRt[k] = nrt[k];

    // RGBDOdometry.cpp:1196-1204 : currentT = [Rprev|tprev] * rgbOdom^-1 in float
    float Rf[9], tf[3], ti[3], Rp[9], tp[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int b = 0; b < 3; ++b) { Rf[a * 3 + b] = (float)nrt[a * 4 + b]; Rp[a * 3 + b] = st->Rprev[a * 3 + b]; }
        tf[a] = (float)nrt[a * 4 + 3];
        tp[a] = st->tprev[a];
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) ti[a] = -dot3_rn(Rf[0 * 3 + a], tf[0], Rf[1 * 3 + a], tf[1], Rf[2 * 3 + a], tf[2]);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int b = 0; b < 3; ++b) st->Rcurr[a * 3 + b] = dot3_rn(Rp[a * 3], Rf[b * 3], Rp[a * 3 + 1], Rf[b * 3 + 1], Rp[a * 3 + 2], Rf[b * 3 + 2]);
        st->tcurr[a] = __fadd_rn(dot3_rn(Rp[a * 3], ti[0], Rp[a * 3 + 1], ti[1], Rp[a * 3 + 2], ti[2]), tp[a]);
    }
    if (next_level >= 0 && rgb) update_krk(st, nrt, next_level);
}

// gn_update for a state in SHARED memory, run by warps 0 and 1 of the CTA (call with all of their 64 threads):
// the normal equations, the pose composition and the new camera pose are spread over lanes; the fp64 solve itself stays on
// one lane; the photometric warp matrices (update_krk) run on warp 1 while warp 0 composes the pose.
#ifndef GN_STAMP
#define GN_STAMP(k)
#endif
__device__ __forceinline__ void gn_update_warps(TrackState* st, int next_level, long long* dbg = nullptr, int* dbg_n_ptr = nullptr)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool icp = st->icp != 0, rgb = st->rgb != 0;
    if (warp == 0) {
        for (int e = lane; e < 42; e += 32) {
            int i, j;
            if (e < 36) { i = e / 6; j = e - 6 * i; if (i > j) { const int t = i; i = j; j = t; } } else { i = e - 36; j = 6; }
            const int idx = i * 7 - (i * (i - 1)) / 2 + (j - i);
            const double vr = (double)(float)st->rgb_sums[idx], vi = (double)(float)st->icp_sums[idx];
            const double w = st->icpWeight;
            double val;
            if (icp && rgb) val = (e < 36) ? vr + w * w * vi : vr + w * vi;
            else val = icp ? vi : vr;
            if (e < 36) st->lastA[e] = val; else st->lastb[e - 36] = val;
        }
        if (lane == 0 && icp) {
            const float r0 = (float)st->icp_sums[27], r1 = (float)st->icp_sums[28];
            st->lastICPError = sqrtf(r0) / r1;
            st->lastICPCount = r1;
            st->icp_iterations_run++;
        }
        __syncwarp();
        GN_STAMP(12);
        if (lane == 0) {
            double A[36], b[6], x[6], Rupd[9];
#pragma unroll
            for (int k = 0; k < 36; ++k) A[k] = st->lastA[k];
#pragma unroll
            for (int k = 0; k < 6; ++k) b[k] = st->lastb[k];
            ldlt_solve<6>(A, b, x);
            GN_STAMP(13);
            const double wv[3] = { x[3], x[4], x[5] };
            rodrigues(wv, Rupd);
            GN_STAMP(14);
#pragma unroll
            for (int k = 0; k < 3; ++k) st->gn_t[k] = x[k];
#pragma unroll
            for (int k = 0; k < 9; ++k) st->gn_R[k] = Rupd[k];
        }
        __syncwarp();
        // resultRt = [exp(w) | t] * resultRt (OdometryProvider.h:71-93): one entry of the top 3 rows per lane
        double nv = 0.0;
        if (lane < 12) {
            const int a = lane >> 2, b = lane & 3;
            nv = st->gn_R[a * 3 + 0] * st->resultRt[0 * 4 + b] + st->gn_R[a * 3 + 1] * st->resultRt[1 * 4 + b] + st->gn_R[a * 3 + 2] * st->resultRt[2 * 4 + b] +
                 st->gn_t[a] * st->resultRt[3 * 4 + b];
        }
        __syncwarp();
        if (lane < 12) st->resultRt[lane] = nv;
        GN_STAMP(15);
    }
    asm volatile("bar.sync 1, 64;" ::: "memory");
    if (warp == 0) {
        // currentT = [Rprev|tprev] * rgbOdom^-1 in float (RGBDOdometry.cpp:1196-1204): one entry per lane
        if (lane < 12) {
            float Rf[9], tf[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
#pragma unroll
                for (int b = 0; b < 3; ++b) Rf[a * 3 + b] = (float)st->resultRt[a * 4 + b];
                tf[a] = (float)st->resultRt[a * 4 + 3];
            }
            if (lane < 9) {
                const int a = lane / 3, b = lane - 3 * a;
                st->Rcurr[lane] = dot3_rn(st->Rprev[a * 3], Rf[b * 3], st->Rprev[a * 3 + 1], Rf[b * 3 + 1], st->Rprev[a * 3 + 2], Rf[b * 3 + 2]);
            } else {
                const int a = lane - 9;
                float ti[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) ti[k] = -dot3_rn(Rf[0 * 3 + k], tf[0], Rf[1 * 3 + k], tf[1], Rf[2 * 3 + k], tf[2]);
                st->tcurr[a] = __fadd_rn(dot3_rn(st->Rprev[a * 3], ti[0], st->Rprev[a * 3 + 1], ti[1], st->Rprev[a * 3 + 2], ti[2]), st->tprev[a]);
            }
        }
    } else if (lane == 0 && next_level >= 0 && rgb) {
        double nrt[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) nrt[k] = st->resultRt[k];
        update_krk(st, nrt, next_level);
    }
}

// SO3 control flow of one iteration (RGBDOdometry.cpp:879-912) from st->so3_sums
__device__ __forceinline__ void so3_update(TrackState* st)
{
    const double* s = st->so3_sums;
    double Ad[9], bd[3], xd[3];
    {
        int shift = 0;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = i; j < 4; ++j) {
                const double value = (double)(float)s[shift++];
                if (j == 3) bd[i] = value; else { Ad[j * 3 + i] = value; Ad[i * 3 + j] = value; }
            }
    }
    const float r0 = (float)s[9], r1 = (float)s[10];
    const float err = sqrtf(r0) / r1;
    st->lastSO3Error = err;
    st->lastSO3Count = r1;
    const float lastError = st->so3_lastError, lastCount = st->so3_lastCount;
    if (err < lastError && lastCount == r1) { st->so3_done = 1; return; }
    if ((double)err > (double)lastError + 0.001) {
        st->lastSO3Error = lastError;
        st->lastSO3Count = lastCount;
#pragma unroll
        for (int k = 0; k < 9; ++k) st->resultR[k] = st->lastResultR[k];
        st->so3_done = 1;
        return;
    }
    st->so3_lastError = err;
    st->so3_lastCount = r1;
#pragma unroll
    for (int k = 0; k < 9; ++k) st->lastResultR[k] = st->resultR[k];
    ldlt_solve<3>(Ad, bd, xd);
    const double xw[3] = { (double)(float)xd[0], (double)(float)xd[1], (double)(float)xd[2] };   // Eigen solves this one in float
    double upd[9], nrd[9];
    rodrigues(xw, upd);
    float Rl[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) Rl[k] = st->R_lr[k];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            // R_lr = rotUpdate.cast<float>() * R_lr (RGBDOdometry.cpp:902): products and sums rounded one by one, as the host code does
            const float v = __fadd_rn(__fadd_rn(__fmul_rn((float)upd[i * 3], Rl[j]), __fmul_rn((float)upd[i * 3 + 1], Rl[3 + j])), __fmul_rn((float)upd[i * 3 + 2], Rl[6 + j]));
            st->R_lr[i * 3 + j] = v;
            nrd[i * 3 + j] = (double)v;
            st->resultR[i * 3 + j] = (double)v;
        }
    update_so3_mats(st, nrd);
}

}  // namespace hrbf
