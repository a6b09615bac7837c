// gn_solve.cuh -- device-side Gauss-Newton bookkeeping: fp64 6x6 / 3x3 LDLT, Rodrigues, SE3
// composition.  Runs in ONE thread of the last block of a reduction kernel, so the whole
// coarse-to-fine loop never leaves the GPU.
//
// Restates the host code of Utils/RGBDOdometry.cpp:825-914 (SO3), :983-992 (K R K^-1, K t),
// :1162-1204 (normal equations, solve, pose update) and Utils/OdometryProvider.h:35-93.
#pragma once
#include "reduce.cuh"
#include <float.h>

namespace hrbf {

// A.ldlt().solve(b) stand-in: LDL^T with symmetric diagonal pivoting, fp64; a vanishing pivot
// contributes 0 (Eigen's solve() behaviour).  N = 6 (SE3) or 3 (SO3).
template <int N>
__device__ inline void ldlt_solve(const double* Ain, const double* bin, double* x)
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

// OdometryProvider.h:35-69
__device__ inline void rodrigues(const double* w, double* R)
{
    double rx = w[0], ry = w[1], rz = w[2];
    const double theta = sqrt(rx * rx + ry * ry + rz * rz);
    for (int k = 0; k < 9; ++k) R[k] = (k % 4 == 0) ? 1.0 : 0.0;
    if (theta >= DBL_EPSILON) {
        const double c = cos(theta), s = sin(theta), c1 = 1. - c, it = 1. / theta;
        rx *= it; ry *= it; rz *= it;
        const double rrt[9] = { rx * rx, rx * ry, rx * rz, rx * ry, ry * ry, ry * rz, rx * rz, ry * rz, rz * rz };
        const double rx_[9] = { 0, -rz, ry, rz, 0, -rx, -ry, rx, 0 };
        for (int k = 0; k < 9; ++k) R[k] = c * ((k % 4 == 0) ? 1.0 : 0.0) + c1 * rrt[k] + s * rx_[k];
    }
}

__device__ inline void inv3(const double* m, double* o)
{
    const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
    const double det = m[0] * c00 + m[1] * c01 + m[2] * c02, id = 1.0 / det;
    o[0] = c00 * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
    o[3] = c01 * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
    o[6] = c02 * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}
__device__ inline void inv3f(const float* m, float* o)
{
    const float c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
    const float det = m[0] * c00 + m[1] * c01 + m[2] * c02, id = 1.0f / det;
    o[0] = c00 * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
    o[3] = c01 * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
    o[6] = c02 * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}
__device__ inline void mul3(const double* a, const double* b, double* o)
{
    double r[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
    for (int k = 0; k < 9; ++k) o[k] = r[k];
}

// K of pyramid level `level` (Cuda/types.cuh:93-97: float division by 2^level)
__device__ inline void level_K(const TrackState* st, int level, double* K)
{
    const int div = 1 << level;
    const float fx = st->fx / div, fy = st->fy / div, cx = st->cx / div, cy = st->cy / div;
    K[0] = fx; K[1] = 0; K[2] = cx; K[3] = 0; K[4] = fy; K[5] = cy; K[6] = 0; K[7] = 0; K[8] = 1;
}

// RGBDOdometry.cpp:983-992 : Rt = resultRt^-1, K R K^-1 and K t for the photometric warp
__device__ inline void update_krk(TrackState* st, int level)
{
    double K[9], Kinv[9], Rm[9], Rinv[9], tm[3], KR[9], KRK[9];
    level_K(st, level, K);
    inv3(K, Kinv);
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) Rm[a * 3 + b] = st->resultRt[a * 4 + b];
    inv3(Rm, Rinv);
    for (int a = 0; a < 3; ++a)
        tm[a] = -(Rinv[a * 3] * st->resultRt[3] + Rinv[a * 3 + 1] * st->resultRt[7] + Rinv[a * 3 + 2] * st->resultRt[11]);
    mul3(K, Rinv, KR);
    mul3(KR, Kinv, KRK);
    for (int k = 0; k < 9; ++k) st->krkinv[k] = (float)KRK[k];
    for (int a = 0; a < 3; ++a) st->kt[a] = (float)(K[a * 3] * tm[0] + K[a * 3 + 1] * tm[1] + K[a * 3 + 2] * tm[2]);
}

// RGBDOdometry.cpp:851-862 : homography K R K^-1, K^-1, K R for the SO3 step (level 2)
__device__ inline void update_so3_mats(TrackState* st)
{
    double K[9], Kinv[9], KR[9], H[9];
    level_K(st, 2, K);
    inv3(K, Kinv);
    mul3(K, st->resultR, KR);
    mul3(KR, Kinv, H);
    for (int k = 0; k < 9; ++k) { st->so3_basis[k] = (float)H[k]; st->so3_kinv[k] = (float)Kinv[k]; st->so3_krlr[k] = (float)KR[k]; }
}

// unpack 27 upper-triangular sums into symmetric A (6x6) and b (reduce.cu:677-689), as floats
__device__ inline void unpack_se3(const double* s, float* A, float* b)
{
    int shift = 0;
    for (int i = 0; i < 6; ++i)
        for (int j = i; j < 7; ++j) {
            const float value = (float)s[shift++];
            if (j == 6) b[i] = value; else A[j * 6 + i] = A[i * 6 + j] = value;
        }
}

// One Gauss-Newton update (RGBDOdometry.cpp:1135-1204) from the reduced sums held in the state.
// next_level: pyramid level of the NEXT iteration (-1: none); cur_level: level just processed.
__device__ inline void gn_update(TrackState* st, int cur_level, int next_level)
{
    float A_icp[36], b_icp[6], A_rgb[36], b_rgb[6];
    double lastA[36], lastb[6], result[6];
    if (st->icp) {
        unpack_se3(st->icp_sums, A_icp, b_icp);
        const float r0 = (float)st->icp_sums[27], r1 = (float)st->icp_sums[28];
        st->lastICPError = sqrtf(r0) / r1;
        st->lastICPCount = r1;
        st->icp_iterations_run++;
    }
    if (st->rgb) unpack_se3(st->rgb_sums, A_rgb, b_rgb);
    if (st->icp && st->rgb) {
        const double w = st->icpWeight;
        for (int k = 0; k < 36; ++k) lastA[k] = (double)A_rgb[k] + w * w * (double)A_icp[k];
        for (int k = 0; k < 6; ++k) lastb[k] = (double)b_rgb[k] + w * (double)b_icp[k];
    } else if (st->icp) {
        for (int k = 0; k < 36; ++k) lastA[k] = A_icp[k];
        for (int k = 0; k < 6; ++k) lastb[k] = b_icp[k];
    } else {
        for (int k = 0; k < 36; ++k) lastA[k] = A_rgb[k];
        for (int k = 0; k < 6; ++k) lastb[k] = b_rgb[k];
    }
    ldlt_solve<6>(lastA, lastb, result);
    for (int k = 0; k < 36; ++k) st->lastA[k] = lastA[k];
    for (int k = 0; k < 6; ++k) st->lastb[k] = lastb[k];

    // OdometryProvider.h:71-93 : resultRt = [exp(w) | t] * resultRt
    double Rupd[9], Rt[16], nrt[16];
    rodrigues(result + 3, Rupd);
    for (int k = 0; k < 16; ++k) Rt[k] = 0.0;
    for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 3; ++b) Rt[a * 4 + b] = Rupd[a * 3 + b];
        Rt[a * 4 + 3] = result[a];
    }
    Rt[15] = 1.0;
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += Rt[a * 4 + k] * st->resultRt[k * 4 + b];
            nrt[a * 4 + b] = s;
        }
    for (int k = 0; k < 16; ++k) st->resultRt[k] = nrt[k];

    // RGBDOdometry.cpp:1196-1204 : currentT = [Rprev|tprev] * rgbOdom^-1 in float
    float Rf[9], tf[3], ti[3];
    for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 3; ++b) Rf[a * 3 + b] = (float)nrt[a * 4 + b];
        tf[a] = (float)nrt[a * 4 + 3];
    }
    for (int a = 0; a < 3; ++a) ti[a] = -(Rf[0 * 3 + a] * tf[0] + Rf[1 * 3 + a] * tf[1] + Rf[2 * 3 + a] * tf[2]);
    for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 3; ++b)
            st->Rcurr[a * 3 + b] = st->Rprev[a * 3] * Rf[b * 3] + st->Rprev[a * 3 + 1] * Rf[b * 3 + 1] + st->Rprev[a * 3 + 2] * Rf[b * 3 + 2];
        st->tcurr[a] = st->Rprev[a * 3] * ti[0] + st->Rprev[a * 3 + 1] * ti[1] + st->Rprev[a * 3 + 2] * ti[2] + st->tprev[a];
    }
    (void)cur_level;
    if (next_level >= 0 && st->rgb) update_krk(st, next_level);
    st->rgb_count = 0;
    st->rgb_sigma = 0;
}

// SO3 control flow of one iteration (RGBDOdometry.cpp:879-912) from st->so3_sums
__device__ inline void so3_update(TrackState* st)
{
    const double* s = st->so3_sums;
    float jtj[9], jtr[3];
    int shift = 0;
    for (int i = 0; i < 3; ++i)
        for (int j = i; j < 4; ++j) {
            const float value = (float)s[shift++];
            if (j == 3) jtr[i] = value; else jtj[j * 3 + i] = jtj[i * 3 + j] = value;
        }
    const float r0 = (float)s[9], r1 = (float)s[10];
    st->lastSO3Error = sqrtf(r0) / r1;
    st->lastSO3Count = r1;
    if (st->lastSO3Error < st->so3_lastError && st->so3_lastCount == st->lastSO3Count) { st->so3_done = 1; return; }
    if ((double)st->lastSO3Error > (double)st->so3_lastError + 0.001) {
        st->lastSO3Error = st->so3_lastError;
        st->lastSO3Count = st->so3_lastCount;
        for (int k = 0; k < 9; ++k) st->resultR[k] = st->lastResultR[k];
        st->so3_done = 1;
        return;
    }
    st->so3_lastError = st->lastSO3Error;
    st->so3_lastCount = st->lastSO3Count;
    for (int k = 0; k < 9; ++k) st->lastResultR[k] = st->resultR[k];
    double Ad[9], bd[3], xd[3], upd[9];
    for (int k = 0; k < 9; ++k) Ad[k] = jtj[k];
    for (int k = 0; k < 3; ++k) bd[k] = jtr[k];
    ldlt_solve<3>(Ad, bd, xd);
    for (int k = 0; k < 3; ++k) xd[k] = (double)(float)xd[k];   // Eigen solves this one in float
    rodrigues(xd, upd);
    float nr[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            nr[i * 3 + j] = (float)upd[i * 3] * st->R_lr[j] + (float)upd[i * 3 + 1] * st->R_lr[3 + j] + (float)upd[i * 3 + 2] * st->R_lr[6 + j];
    for (int k = 0; k < 9; ++k) { st->R_lr[k] = nr[k]; st->resultR[k] = nr[k]; }
    update_so3_mats(st);
}

}  // namespace hrbf
