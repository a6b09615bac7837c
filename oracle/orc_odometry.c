/*
 * oracle/orc_odometry.c -- CPU ORACLE (test infrastructure, never shipped) for
 * SURVEY.md section 8 rows 1-5 (+ the RGB / SO3 branch, 8f.1).
 *
 * Restates, in plain C, the arithmetic of
 *   Core/src/Cuda/reduce.cu            (icpStep / rgbStep / computeRgbResidual / so3Step)
 *   Core/src/Cuda/cudafuncs.cu         (pyramid / map preparation kernels)
 *   Core/src/Utils/RGBDOdometry.cpp    (host Gauss-Newton loop)
 *   Core/src/Utils/OdometryProvider.h  (rodrigues, computeUpdateSE3)
 * Per-pixel arithmetic is fp32 like the reference kernels; the cross-pixel sums
 * (whose order in the reference depends on the launch shape) are accumulated in
 * fp64 so the oracle is the most accurate statement of the sum.
 */
#include "orc.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <limits.h>

static inline float orc_qnan(void) { union { uint32_t u; float f; } v; v.u = 0x7fffffffu; return v.f; }

/* __float2int_rn: round-to-nearest-even, NaN -> 0, saturating (CUDA semantics) */
static inline int f2i_rn(float x)
{
    if (isnan(x)) return 0;
    if (x >= 2147483648.0f) return INT_MAX;
    if (x <= -2147483648.0f) return INT_MIN;
    return (int)rintf(x);
}

static inline void m33_mul_v(const float M[9], const float v[3], float o[3])
{
    float a = M[0] * v[0] + M[1] * v[1] + M[2] * v[2];
    float b = M[3] * v[0] + M[4] * v[1] + M[5] * v[2];
    float c = M[6] * v[0] + M[7] * v[1] + M[8] * v[2];
    o[0] = a; o[1] = b; o[2] = c;
}
static inline void cross3(const float a[3], const float b[3], float o[3])
{
    float x = a[1] * b[2] - a[2] * b[1];
    float y = a[2] * b[0] - a[0] * b[2];
    float z = a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z;
}
static inline float dot3(const float a[3], const float b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline float norm3(const float a[3]) { return sqrtf(dot3(a, a)); }

/* =========================================================== row 5 ===== */

/* cudafuncs.cu:344-383 : AoS RGBA32F -> SoA, invalid (z==0 || n.w<=0) -> NaN in all 4 planes */
void orc_copyMaps(int rows, int cols, const float* v_aos, const float* n_aos, float* vmap, float* nmap)
{
    const float qn = orc_qnan();
    const size_t P = (size_t)rows * cols;
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            size_t i = (size_t)y * cols + x;
            const float* v = v_aos + 4 * i;
            const float* n = n_aos + 4 * i;
            int ok = !(v[2] == 0.0f) && n[3] > 0.0f;
            for (int c = 0; c < 4; ++c) {
                vmap[c * P + i] = ok ? v[c] : qn;
                nmap[c * P + i] = ok ? n[c] : qn;
            }
        }
}

/* cudafuncs.cu:405-431 : valid iff |k| < thr and k is not NaN */
void orc_copyCurvatureMap(int rows, int cols, const float* c_aos, float* cmap, float thr)
{
    const float qn = orc_qnan();
    const size_t P = (size_t)rows * cols;
    for (size_t i = 0; i < P; ++i) {
        const float* c = c_aos + 4 * i;
        int ok = c[3] < thr && c[3] > -thr && !isnan(c[3]);
        for (int k = 0; k < 4; ++k) cmap[k * P + i] = ok ? c[k] : qn;
    }
}

/* cudafuncs.cu:452-470 : valid iff w > 0 */
void orc_copyicpWeightMap(int rows, int cols, const float* w_src, float* w_dst)
{
    const float qn = orc_qnan();
    const size_t P = (size_t)rows * cols;
    for (size_t i = 0; i < P; ++i) w_dst[i] = (w_src[i] > 0.0f) ? w_src[i] : qn;
}

/* cudafuncs.cu:526-587 : 2x2 mean; any NaN in plane x -> only plane x of dst set to NaN */
void orc_resizeMap(int drows, int dcols, const float* src, float* dst, int normalize)
{
    const float qn = orc_qnan();
    const int srows = drows * 2, scols = dcols * 2;
    const size_t SP = (size_t)srows * scols, DP = (size_t)drows * dcols;
    for (int y = 0; y < drows; ++y)
        for (int x = 0; x < dcols; ++x) {
            size_t s00 = (size_t)(2 * y) * scols + 2 * x, s01 = s00 + 1, s10 = s00 + scols, s11 = s10 + 1;
            size_t d = (size_t)y * dcols + x;
            float x00 = src[s00], x01 = src[s01], x10 = src[s10], x11 = src[s11];
            if (isnan(x00) || isnan(x01) || isnan(x10) || isnan(x11)) { dst[d] = qn; continue; }
            float n[3], w;
            n[0] = (x00 + x01 + x10 + x11) / 4;
            n[1] = (src[SP + s00] + src[SP + s01] + src[SP + s10] + src[SP + s11]) / 4;
            n[2] = (src[2 * SP + s00] + src[2 * SP + s01] + src[2 * SP + s10] + src[2 * SP + s11]) / 4;
            w = (src[3 * SP + s00] + src[3 * SP + s01] + src[3 * SP + s10] + src[3 * SP + s11]) / 4;
            if (normalize) { /* operators.cuh:82-86 normalized() = a * rsqrt(dot) */
                float rn = 1.0f / sqrtf(dot3(n, n));
                n[0] *= rn; n[1] *= rn; n[2] *= rn;
            }
            dst[d] = n[0]; dst[DP + d] = n[1]; dst[2 * DP + d] = n[2]; dst[3 * DP + d] = w;
        }
}

/* cudafuncs.cu:618-674 : validity decided on plane w only; NaN -> planes x and w */
void orc_resizeCMap(int drows, int dcols, const float* src, float* dst)
{
    const float qn = orc_qnan();
    const int srows = drows * 2, scols = dcols * 2;
    const size_t SP = (size_t)srows * scols, DP = (size_t)drows * dcols;
    for (int y = 0; y < drows; ++y)
        for (int x = 0; x < dcols; ++x) {
            size_t s00 = (size_t)(2 * y) * scols + 2 * x, s01 = s00 + 1, s10 = s00 + scols, s11 = s10 + 1;
            size_t d = (size_t)y * dcols + x;
            float w00 = src[3 * SP + s00], w01 = src[3 * SP + s01], w10 = src[3 * SP + s10], w11 = src[3 * SP + s11];
            if (isnan(w00) || isnan(w01) || isnan(w10) || isnan(w11)) { dst[d] = qn; dst[3 * DP + d] = qn; continue; }
            dst[d] = (src[s00] + src[s01] + src[s10] + src[s11]) / 4;
            dst[DP + d] = (src[SP + s00] + src[SP + s01] + src[SP + s10] + src[SP + s11]) / 4;
            dst[2 * DP + d] = (src[2 * SP + s00] + src[2 * SP + s01] + src[2 * SP + s10] + src[2 * SP + s11]) / 4;
            dst[3 * DP + d] = (w00 + w01 + w10 + w11) / 4;
        }
}

/* cudafuncs.cu:694-726 */
void orc_resizeicpWeightMap(int drows, int dcols, const float* src, float* dst)
{
    const float qn = orc_qnan();
    const int scols = dcols * 2;
    for (int y = 0; y < drows; ++y)
        for (int x = 0; x < dcols; ++x) {
            size_t s00 = (size_t)(2 * y) * scols + 2 * x;
            float a = src[s00], b = src[s00 + 1], c = src[s00 + scols], d = src[s00 + scols + 1];
            dst[(size_t)y * dcols + x] = (isnan(a) || isnan(b) || isnan(c) || isnan(d)) ? qn : (a + b + c + d) / 4;
        }
}

/* cudafuncs.cu:213-257 : rigid move; NaN pixel -> only plane x written (NaN) */
void orc_tranformMaps(int rows, int cols, const float* vsrc, const float* nsrc,
                      const float R[9], const float t[3], float* vdst, float* ndst)
{
    const float qn = orc_qnan();
    const size_t P = (size_t)rows * cols;
    for (size_t i = 0; i < P; ++i) {
        float vx = vsrc[i];
        if (!isnan(vx)) {
            float v[3] = { vx, vsrc[P + i], vsrc[2 * P + i] }, o[3];
            float w = vsrc[3 * P + i];
            m33_mul_v(R, v, o);
            vdst[i] = o[0] + t[0]; vdst[P + i] = o[1] + t[1]; vdst[2 * P + i] = o[2] + t[2]; vdst[3 * P + i] = w;
        } else vdst[i] = qn;
        float nx = nsrc[i];
        if (!isnan(nx)) {
            float n[3] = { nx, nsrc[P + i], nsrc[2 * P + i] }, o[3];
            float w = nsrc[3 * P + i];
            m33_mul_v(R, n, o);
            ndst[i] = o[0]; ndst[P + i] = o[1]; ndst[2 * P + i] = o[2]; ndst[3 * P + i] = w;
        } else ndst[i] = qn;
    }
}

/* cudafuncs.cu:279-322 : curvature directions rotate, values pass through */
void orc_transformCurvMaps(int rows, int cols, const float* k1src, const float* k2src,
                           const float R[9], const float t[3], float* k1dst, float* k2dst)
{
    (void)t;
    const float qn = orc_qnan();
    const size_t P = (size_t)rows * cols;
    const float* S[2] = { k1src, k2src };
    float* D[2] = { k1dst, k2dst };
    for (int m = 0; m < 2; ++m)
        for (size_t i = 0; i < P; ++i) {
            float x = S[m][i];
            if (!isnan(x)) {
                float v[3] = { x, S[m][P + i], S[m][2 * P + i] }, o[3];
                float k = S[m][3 * P + i];
                m33_mul_v(R, v, o);
                D[m][i] = o[0]; D[m][P + i] = o[1]; D[m][2 * P + i] = o[2]; D[m][3 * P + i] = k;
            } else D[m][i] = qn;
        }
}

/* cudafuncs.cu:57-94 : 5x5 depth pyramid with 3-sigma gate (sigma_color = 30, :103) */
void orc_pyrDownDepth(int srows, int scols, const float* src, float* dst)
{
    const int drows = srows / 2, dcols = scols / 2, D = 5;
    const float sigma_color = 30.0f;
    const float weights[3] = { 0.375f, 0.25f, 0.0625f };
    for (int y = 0; y < drows; ++y)
        for (int x = 0; x < dcols; ++x) {
            float center = src[(size_t)(2 * y) * scols + 2 * x];
            int x_mi = (0 > 2 * x - D / 2 ? 0 : 2 * x - D / 2) - 2 * x;
            int y_mi = (0 > 2 * y - D / 2 ? 0 : 2 * y - D / 2) - 2 * y;
            int x_ma = (scols < 2 * x - D / 2 + D ? scols : 2 * x - D / 2 + D) - 2 * x;
            int y_ma = (srows < 2 * y - D / 2 + D ? srows : 2 * y - D / 2 + D) - 2 * y;
            float sum = 0, wall = 0;
            for (int yi = y_mi; yi < y_ma; ++yi)
                for (int xi = x_mi; xi < x_ma; ++xi) {
                    float val = src[(size_t)(2 * y + yi) * scols + 2 * x + xi];
                    if (fabsf(val - center) < 3 * sigma_color) {
                        float w = weights[abs(xi)] * weights[abs(yi)];
                        sum += val * w; wall += w;
                    }
                }
            dst[(size_t)y * dcols + x] = sum / wall;
        }
}

/* cudafuncs.cu:109-136 */
void orc_createVMap(orc_cam k, int rows, int cols, const float* depth, float* vmap, float cutoff, float factor)
{
    const float qn = orc_qnan();
    const size_t P = (size_t)rows * cols;
    const float fx_inv = 1.f / k.fx, fy_inv = 1.f / k.fy;
    for (int v = 0; v < rows; ++v)
        for (int u = 0; u < cols; ++u) {
            size_t i = (size_t)v * cols + u;
            float z = depth[i] * factor;
            if (z != 0 && z < cutoff) {
                vmap[i] = z * (u - k.cx) * fx_inv;
                vmap[P + i] = z * (v - k.cy) * fy_inv;
                vmap[2 * P + i] = z;
                vmap[3 * P + i] = 1.0f;
            } else vmap[i] = qn;
        }
}

/* cudafuncs.cu:154-195 : forward-difference normals */
void orc_createNMap(int rows, int cols, const float* vmap, float* nmap)
{
    const float qn = orc_qnan();
    const size_t P = (size_t)rows * cols;
    for (int v = 0; v < rows; ++v)
        for (int u = 0; u < cols; ++u) {
            size_t i = (size_t)v * cols + u;
            if (u == cols - 1 || v == rows - 1) { nmap[i] = qn; continue; }
            float a = vmap[i], b = vmap[i + 1], c = vmap[i + cols];
            if (!isnan(a) && !isnan(b) && !isnan(c)) {
                float v00[3] = { a, vmap[P + i], vmap[2 * P + i] };
                float v01[3] = { b, vmap[P + i + 1], vmap[2 * P + i + 1] };
                float v10[3] = { c, vmap[P + i + cols], vmap[2 * P + i + cols] };
                float d1[3] = { v01[0] - v00[0], v01[1] - v00[1], v01[2] - v00[2] };
                float d2[3] = { v10[0] - v00[0], v10[1] - v00[1], v10[2] - v00[2] };
                float r[3]; cross3(d1, d2, r);
                float rn = 1.0f / sqrtf(dot3(r, r));
                nmap[i] = r[0] * rn; nmap[P + i] = r[1] * rn; nmap[2 * P + i] = r[2] * rn; nmap[3 * P + i] = 1.0f;
            } else nmap[i] = qn;
        }
}

/* cudafuncs.cu:874-885 */
void orc_verticesToDepth(int rows, int cols, const float* v_aos, float* depth, float cutoff)
{
    const float qn = orc_qnan();
    for (size_t i = 0; i < (size_t)rows * cols; ++i) {
        float z = v_aos[4 * i + 2];
        depth[i] = (z > cutoff || z <= 0) ? qn : z;
    }
}

static const float kGauss25[25] = { 1, 4, 6, 4, 1, 4, 16, 24, 16, 4, 6, 24, 36, 24, 6, 4, 16, 24, 16, 4, 1, 4, 6, 4, 1 };

/* cudafuncs.cu:493-524 (window clipped with `cols-1` upper bound and a kernel
 * index measured from the clipped end -- reproduced as is) */
void orc_pyrDownGaussF(int srows, int scols, const float* src, float* dst)
{
    const int drows = srows / 2, dcols = scols / 2, D = 5;
    for (int y = 0; y < drows; ++y)
        for (int x = 0; x < dcols; ++x) {
            int tx = 2 * x - D / 2 + D; if (tx > scols - 1) tx = scols - 1;
            int ty = 2 * y - D / 2 + D; if (ty > srows - 1) ty = srows - 1;
            int cy = 2 * y - D / 2; if (cy < 0) cy = 0;
            float sum = 0; int count = 0;
            for (; cy < ty; ++cy) {
                int cx = 2 * x - D / 2; if (cx < 0) cx = 0;
                for (; cx < tx; ++cx) {
                    float s = src[(size_t)cy * scols + cx];
                    if (!isnan(s)) {
                        float g = kGauss25[(ty - cy - 1) * 5 + (tx - cx - 1)];
                        sum = fmaf(s, g, sum);   /* nvcc contracts the reference's `sum += src * gauss` into one FMA (cudafuncs.cu:516) */
                        count = (int)((float)count + g);
                    }
                }
            }
            dst[(size_t)y * dcols + x] = (float)(sum / (float)count);
        }
}

/* cudafuncs.cu:818-848 : as above on u8 with "> 0" validity; float -> u8 truncation
 * (count == 0 gives NaN in the reference; CUDA's float->u8 of NaN is 0) */
void orc_pyrDownUcharGauss(int srows, int scols, const unsigned char* src, unsigned char* dst)
{
    const int drows = srows / 2, dcols = scols / 2, D = 5;
    for (int y = 0; y < drows; ++y)
        for (int x = 0; x < dcols; ++x) {
            int tx = 2 * x - D / 2 + D; if (tx > scols - 1) tx = scols - 1;
            int ty = 2 * y - D / 2 + D; if (ty > srows - 1) ty = srows - 1;
            int cy = 2 * y - D / 2; if (cy < 0) cy = 0;
            float sum = 0; int count = 0;
            for (; cy < ty; ++cy) {
                int cx = 2 * x - D / 2; if (cx < 0) cx = 0;
                for (; cx < tx; ++cx) {
                    unsigned char s = src[(size_t)cy * scols + cx];
                    if (s > 0) {
                        float g = kGauss25[(ty - cy - 1) * 5 + (tx - cx - 1)];
                        sum += (float)s * g;
                        count = (int)((float)count + g);
                    }
                }
            }
            float r = sum / (float)count;
            dst[(size_t)y * dcols + x] = isnan(r) ? 0 : (unsigned char)(int)r;
        }
}

/* cudafuncs.cu:898-911 : texel (x=R,y=G,z=B) weighted 0.114/0.299/0.587, truncated */
void orc_rgbaToIntensity(int rows, int cols, const unsigned char* rgba, unsigned char* dst)
{
    for (size_t i = 0; i < (size_t)rows * cols; ++i) {
        const unsigned char* s = rgba + 4 * i;
        int value = (int)((float)s[0] * 0.114f + (float)s[1] * 0.299f + (float)s[2] * 0.587f);
        dst[i] = (unsigned char)value;
    }
}

/* cudafuncs.cu:930-954 + :970-976 : 3x3 Sobel; the kernel index counts down from
 * 8 over the IN-BOUNDS taps only (border misalignment reproduced) */
void orc_sobel(int rows, int cols, const unsigned char* src, short* dx, short* dy)
{
    static const float gx[9] = { 1, 0, -1, 2, 0, -2, 1, 0, -1 };
    static const float gy[9] = { 1, 2, 1, 0, 0, 0, -1, -2, -1 };
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            float dxv = 0, dyv = 0; int k = 8;
            int j0 = y - 1 < 0 ? 0 : y - 1, j1 = y + 1 > rows - 1 ? rows - 1 : y + 1;
            int i0 = x - 1 < 0 ? 0 : x - 1, i1 = x + 1 > cols - 1 ? cols - 1 : x + 1;
            for (int j = j0; j <= j1; ++j)
                for (int i = i0; i <= i1; ++i) {
                    float s = (float)src[(size_t)j * cols + i];
                    dxv += s * gx[k]; dyv += s * gy[k]; --k;
                }
            dx[(size_t)y * cols + x] = (short)dxv;
            dy[(size_t)y * cols + x] = (short)dyv;
        }
}

/* cudafuncs.cu:995-1013 */
void orc_projectToPointCloud(int rows, int cols, const float* depth, float* cloud3, orc_cam k)
{
    const float invFx = 1.0f / k.fx, invFy = 1.0f / k.fy;
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            size_t i = (size_t)y * cols + x;
            float z = depth[i];
            cloud3[3 * i + 0] = (float)((x - k.cx) * z * invFx);
            cloud3[3 * i + 1] = (float)((y - k.cy) * z * invFy);
            cloud3[3 * i + 2] = z;
        }
}

/* ========================================================= rows 1-3 ===== */

static void unpack_se3(const double s[29], float A[36], float b[6], float residual[2])
{
    /* reduce.cu:677-692 : upper-triangular 6x7 -> symmetric A, b */
    int shift = 0;
    for (int i = 0; i < 6; ++i)
        for (int j = i; j < 7; ++j) {
            float value = (float)s[shift++];
            if (j == 6) b[i] = value; else A[j * 6 + i] = A[i * 6 + j] = value;
        }
    if (residual) { residual[0] = (float)s[27]; residual[1] = (float)s[28]; }
}

static void accum_row7(double acc[29], const float row[7], float weight, int found)
{
    /* reduce.cu:511-545 : weight * row[i] * row[j], i <= j, then residual, inliers */
    int k = 0;
    for (int i = 0; i < 7; ++i)
        for (int j = i; j < 7; ++j) {
            float p = weight * row[i] * row[j];
            acc[k++] += (double)p;
        }
    acc[28] += found ? 1.0 : 0.0;
}

/* reduce.cu:317-435 (search) + :437-548 (getProducts) + :550-572 (sum) */
void orc_icpStep(int rows, int cols,
                 const float Rcurr[9], const float tcurr[3],
                 const float* vmap_curr, const float* nmap_curr,
                 const float* ck1_curr, const float* ck2_curr,
                 const float Rprev_inv[9], const float tprev[3], orc_cam intr,
                 const float* vmap_g_prev, const float* nmap_g_prev,
                 const float* ck1_g_prev, const float* ck2_g_prev,
                 const float* icpw_g_prev, const orc_icp_opts* o,
                 float A[36], float b[6], float residual[2], double* sums29, int* corres)
{
    const size_t P = (size_t)rows * cols;
    double total[29];
    memset(total, 0, sizeof total);
    const int R = o->use_search ? o->radius : 0;
    const int D = R * 2 + 1;

#pragma omp parallel
    {
        double acc[29];
        memset(acc, 0, sizeof acc);
#pragma omp for schedule(static)
        for (int y = 0; y < rows; ++y) {
            for (int x = 0; x < cols; ++x) {
                size_t i = (size_t)y * cols + x;
                float row[7] = { 0, 0, 0, 0, 0, 0, 0 };
                float weight = 1.0f;
                int found = 0, cx_best = -1, cy_best = -1;
                float n_g[3] = { 0, 0, 0 }, d_g[3] = { 0, 0, 0 }, s_g[3] = { 0, 0, 0 };

                do {
                    float vcurr[3] = { vmap_curr[i], vmap_curr[P + i], vmap_curr[2 * P + i] };
                    float vg[3], tmp[3], vcp[3];
                    m33_mul_v(Rcurr, vcurr, vg);
                    vg[0] += tcurr[0]; vg[1] += tcurr[1]; vg[2] += tcurr[2];
                    tmp[0] = vg[0] - tprev[0]; tmp[1] = vg[1] - tprev[1]; tmp[2] = vg[2] - tprev[2];
                    m33_mul_v(Rprev_inv, tmp, vcp);
                    int ux = f2i_rn(vcp[0] * intr.fx / vcp[2] + intr.cx);
                    int uy = f2i_rn(vcp[1] * intr.fy / vcp[2] + intr.cy);
                    if (ux < 0 || uy < 0 || ux >= cols || uy >= rows || vcp[2] < 0) break;

                    float ncurr[3] = { nmap_curr[i], nmap_curr[P + i], nmap_curr[2 * P + i] }, ng[3];
                    m33_mul_v(Rcurr, ncurr, ng);
                    float k1c = ck1_curr[3 * P + i], k2c = ck2_curr[3 * P + i];
                    if (isnan(vcurr[0]) || isnan(ncurr[0]) || isnan(k1c) || isnan(k2c)) break;

                    /* window pass 1: collect candidates, track the largest distance (D_p_R) */
                    float DpR = -1e8f;
                    int cnt = 0;
                    float best_p = 1e8f;
                    /* candidates are re-evaluated in pass 2 (the reference stores <= 25 in local arrays) */
                    for (int pass = 0; pass < 2; ++pass) {
                        for (int cy = uy - D / 2; cy < uy + D / 2 + 1; ++cy)
                            for (int cx = ux - D / 2; cx < ux + D / 2 + 1; ++cx) {
                                if (cx < 0 || cy < 0 || cx >= cols || cy >= rows) continue;
                                size_t j = (size_t)cy * cols + cx;
                                float vp[3] = { vmap_g_prev[j], vmap_g_prev[P + j], vmap_g_prev[2 * P + j] };
                                float np[3] = { nmap_g_prev[j], nmap_g_prev[P + j], nmap_g_prev[2 * P + j] };
                                float k1 = ck1_g_prev[3 * P + j], k2 = ck2_g_prev[3 * P + j];
                                float df[3] = { vp[0] - vg[0], vp[1] - vg[1], vp[2] - vg[2] };
                                float dist = norm3(df);
                                float cr[3]; cross3(ng, np, cr);
                                float sine = norm3(cr);
                                if (isnan(vp[0]) || isnan(np[0]) || isnan(k1) || isnan(k2)) continue;
                                if (sine > o->angle_thres || dist > o->dist_thres) continue;
                                if (pass == 0) { if (dist > DpR) DpR = dist; cnt++; continue; }
                                /* pass 2 : reduce.cu:409-430 */
                                float a1 = fabsf(k1), a2 = fabsf(k2);
                                float ckmax = a1 > a2 ? a1 : a2;
                                float D_p = dist / DpR;
                                float D_n = 1 - dot3(np, ng);
                                float D_c = 1 - expf(-fabsf(k1 - k1c) / ckmax) * expf(-fabsf(k2 - k2c) / ckmax);
                                float p = o->use_search ? 0.333f * D_p + 0.333f * D_n + 0.333f * D_c : 1.0f;
                                if (p < best_p) {
                                    cx_best = cx; cy_best = cy;
                                    memcpy(d_g, vp, sizeof d_g); memcpy(n_g, np, sizeof n_g);
                                    best_p = p;
                                }
                                found = 1;
                            }
                        if (pass == 0 && cnt == 0) break;
                    }
                    /* When every candidate's score is NaN (ckmax == 0 or D_p_R == 0) nothing is selected, yet the
                     * reference still returns if_found = true with UNINITIALISED vprev_g / nprev_g and
                     * corres = (-1,-1) (reduce.cu:404-434): undefined behaviour.  Observed with the reference's own
                     * kernel on a B200 (GPUTest pair, zero curvature): NaN sums in one run, zero sums in the next.
                     * The oracle defines the case as "no correspondence". */
                    if (cx_best < 0) found = 0;
                    if (found) memcpy(s_g, vg, sizeof s_g);
                } while (0);

                if (corres) { corres[2 * i] = cx_best; corres[2 * i + 1] = cy_best; }

                if (found) {
                    float t1[3], s_cp[3], d_cp[3], n_cp[3];
                    t1[0] = s_g[0] - tprev[0]; t1[1] = s_g[1] - tprev[1]; t1[2] = s_g[2] - tprev[2];
                    m33_mul_v(Rprev_inv, t1, s_cp);
                    t1[0] = d_g[0] - tprev[0]; t1[1] = d_g[1] - tprev[1]; t1[2] = d_g[2] - tprev[2];
                    m33_mul_v(Rprev_inv, t1, d_cp);
                    m33_mul_v(Rprev_inv, n_g, n_cp);
                    if (o->use_weight) {
                        float w = icpw_g_prev[(size_t)cy_best * cols + cx_best];
                        weight = isnan(w) ? 0.0f : w;
                    }
                    row[0] = n_cp[0]; row[1] = n_cp[1]; row[2] = n_cp[2];
                    cross3(s_cp, n_cp, row + 3);
                    float df[3] = { s_cp[0] - d_cp[0], s_cp[1] - d_cp[1], s_cp[2] - d_cp[2] };
                    row[6] = dot3(n_cp, df);
                }
                accum_row7(acc, row, weight, found);
            }
        }
#pragma omp critical
        for (int k = 0; k < 29; ++k) total[k] += acc[k];
    }
    if (sums29) memcpy(sums29, total, sizeof total);
    unpack_se3(total, A, b, residual);
}

/* reduce.cu:986-1060 (per pixel) + :1062-1080, 1152-1153 (int2 sum) */
void orc_computeRgbResidual(int rows, int cols, float minScale,
                            const short* dIdx, const short* dIdy,
                            const float* lastDepth, const float* nextDepth,
                            const unsigned char* lastImage, const unsigned char* nextImage,
                            orc_dataterm* corresImg, float maxDepthDelta,
                            const float kt[3], const float K[9],
                            int* sigmaSum, int* count)
{
    long long cnt = 0, sig = 0;
    for (int i = 0; i < rows; ++i)
        for (int j0 = 0; j0 < cols; ++j0) {
            size_t k = (size_t)i * cols + j0;
            orc_dataterm c; memset(&c, 0, sizeof c);
            if (j0 < cols - 5 && i < rows - 1) {
                int valid = 1;
                int u1 = i + 2 < rows ? i + 2 : rows, v1 = j0 + 2 < cols ? j0 + 2 : cols;
                for (int u = (i - 2 > 0 ? i - 2 : 0); u < u1; ++u)
                    for (int v = (j0 - 2 > 0 ? j0 - 2 : 0); v < v1; ++v)
                        valid = valid && (nextImage[(size_t)u * cols + v] > 0);
                if (valid) {
                    short valx = dIdx[k], valy = dIdy[k];
                    float mTwo = (float)((valx * valx) + (valy * valy));
                    if (mTwo >= minScale) {
                        int y = i, x = j0;
                        float d1 = nextDepth[k];
                        if (!isnan(d1)) {
                            /* the warp d1 * (K.x * x + K.y * y + K.z) + kt as nvcc contracts it in the reference's build (read off the
                             * SASS of residualKernel in oracle/_ref/libref_reduce.so): FMUL y*K.y, FFMA x*K.x + ., FADD + K.z, FFMA d1 * . + kt.
                             * The division stays IEEE (the reference build's MUFU.RCP cannot be restated) */
#define ORC_WARP_ROW(r, t) fmaf(d1, fmaf((float)x, K[3 * (r)], (float)y * K[3 * (r) + 1]) + K[3 * (r) + 2], (t))
                            float td1 = ORC_WARP_ROW(2, kt[2]);
                            int u0 = f2i_rn(ORC_WARP_ROW(0, kt[0]) / td1);
                            int v0 = f2i_rn(ORC_WARP_ROW(1, kt[1]) / td1);
#undef ORC_WARP_ROW
                            if (u0 >= 0 && v0 >= 0 && u0 < cols && v0 < rows) {
                                float d0 = lastDepth[(size_t)v0 * cols + u0];
                                if (d0 > 0 && fabsf(td1 - d0) <= maxDepthDelta && lastImage[(size_t)v0 * cols + u0] != 0) {
                                    c.zx = (short)u0; c.zy = (short)v0; c.ox = (short)x; c.oy = (short)y;
                                    c.diff = (float)nextImage[k] - (float)lastImage[(size_t)v0 * cols + u0];
                                    c.valid = 1;
                                    cnt += 1;
                                    sig += (int)(c.diff * c.diff);
                                }
                            }
                        }
                    }
                }
            }
            corresImg[k] = c;
        }
    *count = (int)cnt; *sigmaSum = (int)sig;
}

/* reduce.cu:718-811 + host unpack :881-895 */
void orc_rgbStep(int rows, int cols, const orc_dataterm* corresImg, float sigma,
                 const float* cloud3, float fx, float fy,
                 const short* dIdx, const short* dIdy, int use_grad_weight, float sobelScale,
                 float A[36], float b[6], double* sums29)
{
    double acc[29]; memset(acc, 0, sizeof acc);
    for (size_t i = 0; i < (size_t)rows * cols; ++i) {
        const orc_dataterm* c = &corresImg[i];
        float row[7] = { 0, 0, 0, 0, 0, 0, 0 };
        float rgb_weight = 1.0f;
        if (c->valid) {
            float w = sigma + fabsf(c->diff);
            w = w > 1.19209290E-07F ? 1.0f / w : 1.0f;
            if (sigma == -1) w = 1;
            row[6] = -w * c->diff;
            const float* cp = cloud3 + 3 * ((size_t)c->zy * cols + c->zx);
            float invz = (float)(1.0 / cp[2]);
            float gxv = w * sobelScale * dIdx[(size_t)c->oy * cols + c->ox];
            float gyv = w * sobelScale * dIdy[(size_t)c->oy * cols + c->ox];
            float v0 = gxv * fx * invz, v1 = gyv * fy * invz;
            float v2 = -(v0 * cp[0] + v1 * cp[1]) * invz;
            row[0] = v0; row[1] = v1; row[2] = v2;
            row[3] = -cp[2] * v1 + cp[1] * v2;
            row[4] = cp[2] * v0 - cp[0] * v2;
            row[5] = -cp[1] * v0 + cp[0] * v1;
            if (use_grad_weight) {
                float gm = sqrtf(gxv * gxv + gyv * gyv);
                rgb_weight = (float)exp(-0.5 * (10 / gm) * (10 / gm));
            }
        }
        accum_row7(acc, row, rgb_weight, c->valid);
    }
    if (sums29) memcpy(sums29, acc, sizeof acc);
    unpack_se3(acc, A, b, NULL);
}

static inline void grad_u8(const unsigned char* img, int cols, int x, int y, float* gx, float* gy)
{
    /* reduce.cu:1172-1188 */
    float actu = (float)img[(size_t)y * cols + x];
    float back = (float)img[(size_t)y * cols + x - 1], fore = (float)img[(size_t)y * cols + x + 1];
    *gx = ((back + actu) / 2.0f) - ((fore + actu) / 2.0f);
    back = (float)img[(size_t)(y - 1) * cols + x]; fore = (float)img[(size_t)(y + 1) * cols + x];
    *gy = ((back + actu) / 2.0f) - ((fore + actu) / 2.0f);
}

/* reduce.cu:1190-1273 + unpack :1344-1358 */
void orc_so3Step(int rows, int cols, const unsigned char* lastImage, const unsigned char* nextImage,
                 const float B[9], const float kinv[9], const float krlr[9],
                 float A[9], float bv[3], float residual[2], double* sums11)
{
    double acc[11]; memset(acc, 0, sizeof acc);
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            float up[3] = { (float)x, (float)y, 1.0f }, wp[3];
            m33_mul_v(B, up, wp);
            int wx = f2i_rn(wp[0] / wp[2]), wy = f2i_rn(wp[1] / wp[2]);
            int found = wx >= 1 && wx < cols - 1 && wy >= 1 && wy < rows - 1 &&
                        x >= 1 && x < cols - 1 && y >= 1 && y < rows - 1;
            float row[4] = { 0, 0, 0, 0 };
            if (found) {
                float gnx, gny, glx, gly;
                grad_u8(nextImage, cols, wx, wy, &gnx, &gny);
                grad_u8(lastImage, cols, x, y, &glx, &gly);
                float gx = (gnx + glx) / 2.0f, gy = (gny + gly) / 2.0f;
                float pt[3]; m33_mul_v(kinv, up, pt);
                float z2 = pt[2] * pt[2];
                float a = krlr[0], b = krlr[1], c = krlr[2], d = krlr[3], e = krlr[4], f = krlr[5],
                      g = krlr[6], h = krlr[7], ii = krlr[8];
                float lp[3] = { ((pt[2] * (d * gy + a * gx)) - (gy * g * y) - (gx * g * x)) / z2,
                                ((pt[2] * (e * gy + b * gx)) - (gy * h * y) - (gx * h * x)) / z2,
                                ((pt[2] * (f * gy + c * gx)) - (gy * ii * y) - (gx * ii * x)) / z2 };
                float jr[3]; cross3(lp, pt, jr);
                row[0] = jr[0]; row[1] = jr[1]; row[2] = jr[2];
                row[3] = -((float)nextImage[(size_t)wy * cols + wx] - (float)lastImage[(size_t)y * cols + x]);
            }
            int k = 0;
            for (int i = 0; i < 4; ++i)
                for (int j = i; j < 4; ++j) acc[k++] += (double)(row[i] * row[j]);
            acc[10] += found ? 1.0 : 0.0;
        }
    if (sums11) memcpy(sums11, acc, sizeof acc);
    int shift = 0;
    for (int i = 0; i < 3; ++i)
        for (int j = i; j < 4; ++j) {
            float value = (float)acc[shift++];
            if (j == 3) bv[i] = value; else A[j * 3 + i] = A[i * 3 + j] = value;
        }
    residual[0] = (float)acc[9]; residual[1] = (float)acc[10];
}

/* ============================================================ solves ===== */

/* OdometryProvider.h:35-69 */
void orc_rodrigues(const double w[3], double R[9])
{
    double rx = w[0], ry = w[1], rz = w[2];
    double theta = sqrt(rx * rx + ry * ry + rz * rz);
    for (int k = 0; k < 9; ++k) R[k] = (k % 4 == 0) ? 1.0 : 0.0;
    if (theta >= DBL_EPSILON) {
        double c = cos(theta), s = sin(theta), c1 = 1. - c, it = 1. / theta;
        rx *= it; ry *= it; rz *= it;
        double rrt[9] = { rx * rx, rx * ry, rx * rz, rx * ry, ry * ry, ry * rz, rx * rz, ry * rz, rz * rz };
        double rx_[9] = { 0, -rz, ry, rz, 0, -rx, -ry, rx, 0 };
        for (int k = 0; k < 9; ++k) R[k] = c * ((k % 4 == 0) ? 1.0 : 0.0) + c1 * rrt[k] + s * rx_[k];
    }
}

/* Stand-in for Eigen's A.ldlt().solve(b) (RGBDOdometry.cpp:1173-1185; Eigen is an
 * un-vendored, un-pinned dependency): LDL^T with symmetric diagonal pivoting, and a
 * (near-)zero pivot contributes 0 to the solution as Eigen's solve does. */
static void ldlt_solve_n(int n, const double* Ain, const double* bin, double* x)
{
    double A[36], b[6], y[6];
    int perm[6];
    for (int i = 0; i < n * n; ++i) A[i] = Ain[i];
    for (int i = 0; i < n; ++i) { b[i] = bin[i]; perm[i] = i; }
    for (int k = 0; k < n; ++k) {
        int p = k; double big = fabs(A[k * n + k]);
        for (int i = k + 1; i < n; ++i) if (fabs(A[i * n + i]) > big) { big = fabs(A[i * n + i]); p = i; }
        if (p != k) {
            for (int j = 0; j < n; ++j) { double t = A[k * n + j]; A[k * n + j] = A[p * n + j]; A[p * n + j] = t; }
            for (int j = 0; j < n; ++j) { double t = A[j * n + k]; A[j * n + k] = A[j * n + p]; A[j * n + p] = t; }
            int t = perm[k]; perm[k] = perm[p]; perm[p] = t;
        }
        double d = A[k * n + k];
        if (fabs(d) <= DBL_MIN) continue;
        for (int i = k + 1; i < n; ++i) A[i * n + k] /= d;
        for (int i = k + 1; i < n; ++i)
            for (int j = k + 1; j <= i; ++j) {
                A[i * n + j] -= A[i * n + k] * d * A[j * n + k];
                A[j * n + i] = A[i * n + j];
            }
    }
    for (int i = 0; i < n; ++i) y[i] = b[perm[i]];
    for (int i = 0; i < n; ++i) for (int j = 0; j < i; ++j) y[i] -= A[i * n + j] * y[j];
    for (int i = 0; i < n; ++i) { double d = A[i * n + i]; y[i] = (fabs(d) > DBL_MIN) ? y[i] / d : 0.0; }
    for (int i = n - 1; i >= 0; --i) for (int j = i + 1; j < n; ++j) y[i] -= A[j * n + i] * y[j];
    for (int i = 0; i < n; ++i) x[perm[i]] = y[i];
}

void orc_ldlt_solve6(const double A[36], const double b[6], double x[6]) { ldlt_solve_n(6, A, b, x); }

/* OdometryProvider.h:71-93 computeUpdateSE3: resultRt (row-major 4x4, in/out) = [rodrigues(omega) | t; 0 1] * resultRt with
 * xi = (t, omega); rgbOdom (iso16, row-major) = the float cast of its rotation and translation */
void orc_computeUpdateSE3(double resultRt[16], const double result[6], float iso16[16])
{
    double Rupd[9], Rt[16] = { 0 }, nrt[16];
    orc_rodrigues(result + 3, Rupd);
    for (int a = 0; a < 3; ++a) { for (int b2 = 0; b2 < 3; ++b2) Rt[a * 4 + b2] = Rupd[a * 3 + b2]; Rt[a * 4 + 3] = result[a]; }
    Rt[15] = 1;
    for (int a = 0; a < 4; ++a) for (int b2 = 0; b2 < 4; ++b2) {
        double s = 0; for (int k = 0; k < 4; ++k) s += Rt[a * 4 + k] * resultRt[k * 4 + b2];
        nrt[a * 4 + b2] = s;
    }
    memcpy(resultRt, nrt, sizeof nrt);
    for (int k = 0; k < 16; ++k) iso16[k] = (k % 5 == 0) ? 1.0f : 0.0f;
    for (int a = 0; a < 3; ++a) { for (int b2 = 0; b2 < 3; ++b2) iso16[a * 4 + b2] = (float)resultRt[a * 4 + b2]; iso16[a * 4 + 3] = (float)resultRt[a * 4 + 3]; }
}

/* RGBDOdometry.cpp:900 : the 3x3 SO3 system is solved in float by Eigen; the
 * oracle solves the float-valued system in double and rounds (differences are
 * below float epsilon on a well-conditioned 3x3). */
void orc_ldlt_solve3f(const float A[9], const float b[3], float x[3])
{
    double Ad[9], bd[3], xd[3];
    for (int i = 0; i < 9; ++i) Ad[i] = A[i];
    for (int i = 0; i < 3; ++i) bd[i] = b[i];
    ldlt_solve_n(3, Ad, bd, xd);
    for (int i = 0; i < 3; ++i) x[i] = (float)xd[i];
}

static void inv3d(const double m[9], double o[9])
{
    double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
    double det = m[0] * c00 + m[1] * c01 + m[2] * c02, id = 1.0 / det;
    o[0] = c00 * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
    o[3] = c01 * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
    o[6] = c02 * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}
static void inv3f(const float m[9], float o[9])
{
    float c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
    float det = m[0] * c00 + m[1] * c01 + m[2] * c02, id = 1.0f / det;
    o[0] = c00 * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
    o[3] = c01 * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
    o[6] = c02 * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}
static void mul3d(const double a[9], const double b[9], double o[9])
{
    double r[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
        r[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
    memcpy(o, r, sizeof r);
}

/* ============================================================ row 4 ===== */

struct orc_odom {
    int width, height;
    orc_cam intr;
    float distThres, angleThres;
    float sobelScale, maxDepthDeltaRGB, maxDepthRGB;
    float minGrad[ORC_NUM_PYRS];
    float* maps[9][ORC_NUM_PYRS];          /* see orc_odom_map() */
    float* vmaps_tmp;                      /* AoS copy of the last vertex texture (RGBDOdometry.cpp:189,217) */
    float* depth_tmp[ORC_NUM_PYRS];
    float* lastDepth[ORC_NUM_PYRS]; float* nextDepth[ORC_NUM_PYRS];
    unsigned char* lastImage[ORC_NUM_PYRS]; unsigned char* nextImage[ORC_NUM_PYRS]; unsigned char* lastNextImage[ORC_NUM_PYRS];
    short* dIdx[ORC_NUM_PYRS]; short* dIdy[ORC_NUM_PYRS];
    float* cloud[ORC_NUM_PYRS];
    orc_dataterm* corresImg[ORC_NUM_PYRS];
};

static orc_cam cam_level(orc_cam k, int level)
{   /* types.cuh:93-97 */
    int div = 1 << level;
    orc_cam r = { k.fx / div, k.fy / div, k.cx / div, k.cy / div };
    return r;
}

orc_odom* orc_odom_create(int width, int height, float cx, float cy, float fx, float fy, float distThresh, float angleThresh)
{
    orc_odom* o = (orc_odom*)calloc(1, sizeof *o);
    o->width = width; o->height = height;
    o->intr.cx = cx; o->intr.cy = cy; o->intr.fx = fx; o->intr.fy = fy;
    o->distThres = distThresh; o->angleThres = angleThresh;
    o->sobelScale = (float)(1.0 / pow(2.0, 3)); o->maxDepthDeltaRGB = 0.07f; o->maxDepthRGB = 6.0f;
    o->minGrad[0] = 5; o->minGrad[1] = 3; o->minGrad[2] = 1;
    for (int l = 0; l < ORC_NUM_PYRS; ++l) {
        size_t P = (size_t)(height >> l) * (width >> l);
        for (int m = 0; m < 8; ++m) o->maps[m][l] = (float*)calloc(4 * P, sizeof(float));
        o->maps[8][l] = (float*)calloc(P, sizeof(float));
        o->depth_tmp[l] = (float*)calloc(P, sizeof(float));
        o->lastDepth[l] = (float*)calloc(P, sizeof(float)); o->nextDepth[l] = (float*)calloc(P, sizeof(float));
        o->lastImage[l] = (unsigned char*)calloc(P, 1); o->nextImage[l] = (unsigned char*)calloc(P, 1);
        o->lastNextImage[l] = (unsigned char*)calloc(P, 1);
        o->dIdx[l] = (short*)calloc(P, sizeof(short)); o->dIdy[l] = (short*)calloc(P, sizeof(short));
        o->cloud[l] = (float*)calloc(3 * P, sizeof(float));
        o->corresImg[l] = (orc_dataterm*)calloc(P, sizeof(orc_dataterm));
    }
    o->vmaps_tmp = (float*)calloc((size_t)height * width * 4, sizeof(float));
    return o;
}

void orc_odom_destroy(orc_odom* o)
{
    if (!o) return;
    for (int l = 0; l < ORC_NUM_PYRS; ++l) {
        for (int m = 0; m < 9; ++m) free(o->maps[m][l]);
        free(o->depth_tmp[l]); free(o->lastDepth[l]); free(o->nextDepth[l]);
        free(o->lastImage[l]); free(o->nextImage[l]); free(o->lastNextImage[l]);
        free(o->dIdx[l]); free(o->dIdy[l]); free(o->cloud[l]); free(o->corresImg[l]);
    }
    free(o->vmaps_tmp); free(o);
}

const float* orc_odom_map(const orc_odom* o, int which, int level) { return o->maps[which][level]; }
const unsigned char* orc_odom_image(const orc_odom* o, int which, int level)
{ return which == 0 ? o->lastImage[level] : which == 1 ? o->nextImage[level] : o->lastNextImage[level]; }
const float* orc_odom_depth(const orc_odom* o, int which, int level) { return which == 0 ? o->lastDepth[level] : o->nextDepth[level]; }

enum { M_VG = 0, M_NG, M_K1G, M_K2G, M_VC, M_NC, M_K1C, M_K2C, M_W };

/* RGBDOdometry.cpp:161-181 (the GPUTest path: depth pyramid -> vmap/nmap per level) */
void orc_odom_initICP_depth(orc_odom* o, const float* depth_raw, float depthCutoff, float depthFactor)
{
    memcpy(o->depth_tmp[0], depth_raw, sizeof(float) * (size_t)o->width * o->height);
    for (int i = 1; i < ORC_NUM_PYRS; ++i)
        orc_pyrDownDepth(o->height >> (i - 1), o->width >> (i - 1), o->depth_tmp[i - 1], o->depth_tmp[i]);
    for (int i = 0; i < ORC_NUM_PYRS; ++i) {
        orc_createVMap(cam_level(o->intr, i), o->height >> i, o->width >> i, o->depth_tmp[i], o->maps[M_VC][i], depthCutoff, depthFactor);
        orc_createNMap(o->height >> i, o->width >> i, o->maps[M_VC][i], o->maps[M_NC][i]);
    }
}

/* RGBDOdometry.cpp:183-206 */
void orc_odom_initICP(orc_odom* o, const float* vert_aos, const float* norm_aos, float depthCutoff)
{
    (void)depthCutoff;
    memcpy(o->vmaps_tmp, vert_aos, sizeof(float) * 4 * (size_t)o->width * o->height);
    orc_copyMaps(o->height, o->width, vert_aos, norm_aos, o->maps[M_VC][0], o->maps[M_NC][0]);
    for (int i = 1; i < ORC_NUM_PYRS; ++i) {
        orc_resizeMap(o->height >> i, o->width >> i, o->maps[M_VC][i - 1], o->maps[M_VC][i], 0);
        orc_resizeMap(o->height >> i, o->width >> i, o->maps[M_NC][i - 1], o->maps[M_NC][i], 1);
    }
}

static void pose_split(const float pose[16], float R[9], float t[3])
{
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) R[i * 3 + j] = pose[i * 4 + j]; t[i] = pose[i * 4 + 3]; }
}

/* RGBDOdometry.cpp:208-247 */
void orc_odom_initICPModel(orc_odom* o, const float* vert_aos, const float* norm_aos, float depthCutoff, const float pose[16])
{
    (void)depthCutoff;
    float R[9], t[3]; pose_split(pose, R, t);
    memcpy(o->vmaps_tmp, vert_aos, sizeof(float) * 4 * (size_t)o->width * o->height);
    orc_copyMaps(o->height, o->width, vert_aos, norm_aos, o->maps[M_VG][0], o->maps[M_NG][0]);
    for (int i = 1; i < ORC_NUM_PYRS; ++i) {
        orc_resizeMap(o->height >> i, o->width >> i, o->maps[M_VG][i - 1], o->maps[M_VG][i], 0);
        orc_resizeMap(o->height >> i, o->width >> i, o->maps[M_NG][i - 1], o->maps[M_NG][i], 1);
    }
    for (int i = 0; i < ORC_NUM_PYRS; ++i)
        orc_tranformMaps(o->height >> i, o->width >> i, o->maps[M_VG][i], o->maps[M_NG][i], R, t, o->maps[M_VG][i], o->maps[M_NG][i]);
}

/* RGBDOdometry.cpp:660-687 */
static void populateRGBDData(orc_odom* o, const unsigned char* rgba, float** depths, unsigned char** images)
{
    orc_verticesToDepth(o->height, o->width, o->vmaps_tmp, depths[0], o->maxDepthRGB);
    for (int i = 0; i + 1 < ORC_NUM_PYRS; ++i) orc_pyrDownGaussF(o->height >> i, o->width >> i, depths[i], depths[i + 1]);
    orc_rgbaToIntensity(o->height, o->width, rgba, images[0]);
    for (int i = 0; i + 1 < ORC_NUM_PYRS; ++i) orc_pyrDownUcharGauss(o->height >> i, o->width >> i, images[i], images[i + 1]);
}
void orc_odom_initRGB(orc_odom* o, const unsigned char* rgba) { populateRGBDData(o, rgba, o->nextDepth, o->nextImage); }
void orc_odom_initRGBModel(orc_odom* o, const unsigned char* rgba) { populateRGBDData(o, rgba, o->lastDepth, o->lastImage); }
/* RGBDOdometry.cpp:777-794 */
void orc_odom_initFirstRGB(orc_odom* o, const unsigned char* rgba)
{
    orc_rgbaToIntensity(o->height, o->width, rgba, o->lastNextImage[0]);
    for (int i = 0; i + 1 < ORC_NUM_PYRS; ++i) orc_pyrDownUcharGauss(o->height >> i, o->width >> i, o->lastNextImage[i], o->lastNextImage[i + 1]);
}

/* RGBDOdometry.cpp:701-723 */
void orc_odom_initCurvature(orc_odom* o, const float* k1_aos, const float* k2_aos, float thr)
{
    orc_copyCurvatureMap(o->height, o->width, k1_aos, o->maps[M_K1C][0], thr);
    orc_copyCurvatureMap(o->height, o->width, k2_aos, o->maps[M_K2C][0], thr);
    for (int i = 1; i < ORC_NUM_PYRS; ++i) {
        orc_resizeCMap(o->height >> i, o->width >> i, o->maps[M_K1C][i - 1], o->maps[M_K1C][i]);
        orc_resizeCMap(o->height >> i, o->width >> i, o->maps[M_K2C][i - 1], o->maps[M_K2C][i]);
    }
}
/* RGBDOdometry.cpp:725-759 */
void orc_odom_initCurvatureModel(orc_odom* o, const float* k1_aos, const float* k2_aos, const float pose[16], float thr)
{
    float R[9], t[3]; pose_split(pose, R, t);
    orc_copyCurvatureMap(o->height, o->width, k1_aos, o->maps[M_K1G][0], thr);
    orc_copyCurvatureMap(o->height, o->width, k2_aos, o->maps[M_K2G][0], thr);
    for (int i = 1; i < ORC_NUM_PYRS; ++i) {
        orc_resizeCMap(o->height >> i, o->width >> i, o->maps[M_K1G][i - 1], o->maps[M_K1G][i]);
        orc_resizeCMap(o->height >> i, o->width >> i, o->maps[M_K2G][i - 1], o->maps[M_K2G][i]);
    }
    for (int i = 0; i < ORC_NUM_PYRS; ++i)
        orc_transformCurvMaps(o->height >> i, o->width >> i, o->maps[M_K1G][i], o->maps[M_K2G][i], R, t, o->maps[M_K1G][i], o->maps[M_K2G][i]);
}
/* RGBDOdometry.cpp:761-775 */
void orc_odom_initICPweight(orc_odom* o, const float* w)
{
    orc_copyicpWeightMap(o->height, o->width, w, o->maps[M_W][0]);
    for (int i = 1; i < ORC_NUM_PYRS; ++i)
        orc_resizeicpWeightMap(o->height >> i, o->width >> i, o->maps[M_W][i - 1], o->maps[M_W][i]);
}

void orc_odom_fillNeutralCurvature(orc_odom* o)
{
    for (int l = 0; l < ORC_NUM_PYRS; ++l) {
        size_t P = (size_t)(o->height >> l) * (o->width >> l);
        for (int m = M_K1G; m <= M_K2G; ++m) memset(o->maps[m][l], 0, 4 * P * sizeof(float));
        for (int m = M_K1C; m <= M_K2C; ++m) memset(o->maps[m][l], 0, 4 * P * sizeof(float));
        for (size_t i = 0; i < P; ++i) o->maps[M_W][l][i] = 1.0f;
    }
}

/* RGBDOdometry.cpp:796-1249 */
void orc_odom_getIncrementalTransformation(orc_odom* o, float trans[3], float rot[9],
                                           const orc_track_opts* opt, orc_track_stats* st)
{
    const int icp = !opt->rgbOnly && opt->icpWeight > 0;
    const int rgb = opt->rgbOnly || opt->icpWeight < 100;
    float Rprev[9], tprev[3], Rcurr[9], tcurr[3];
    memcpy(Rprev, rot, sizeof Rprev); memcpy(tprev, trans, sizeof tprev);
    memcpy(Rcurr, Rprev, sizeof Rcurr); memcpy(tcurr, tprev, sizeof tcurr);
    orc_track_stats local; if (!st) st = &local;
    memset(st, 0, sizeof *st);

    if (rgb)
        for (int i = 0; i < ORC_NUM_PYRS; ++i)
            orc_sobel(o->height >> i, o->width >> i, o->nextImage[i], o->dIdx[i], o->dIdy[i]);

    double resultR[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };

    if (opt->so3) {   /* :827-914 */
        const int lvl = 2;
        orc_cam kl = cam_level(o->intr, lvl);
        float R_lr[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
        double K[9] = { kl.fx, 0, kl.cx, 0, kl.fy, kl.cy, 0, 0, 1 }, Kinv[9];
        inv3d(K, Kinv);
        float lastError = FLT_MAX / 2, lastCount = FLT_MAX / 2;
        double lastResultR[9]; memcpy(lastResultR, resultR, sizeof resultR);
        for (int it = 0; it < 10; ++it) {
            double H[9], KR[9];
            mul3d(K, resultR, KR); mul3d(KR, Kinv, H);
            float Bf[9], kinvf[9], krlrf[9];
            for (int k = 0; k < 9; ++k) { Bf[k] = (float)H[k]; kinvf[k] = (float)Kinv[k]; krlrf[k] = (float)KR[k]; }
            float jtj[9], jtr[3], residual[2];
            orc_so3Step(o->height >> lvl, o->width >> lvl, o->lastNextImage[lvl], o->nextImage[lvl], Bf, kinvf, krlrf, jtj, jtr, residual, NULL);
            st->lastSO3Error = sqrtf(residual[0]) / residual[1];
            st->lastSO3Count = residual[1];
            if (st->lastSO3Error < lastError && lastCount == st->lastSO3Count) break;
            else if (st->lastSO3Error > lastError + 0.001) {
                st->lastSO3Error = lastError; st->lastSO3Count = lastCount;
                memcpy(resultR, lastResultR, sizeof resultR);
                break;
            }
            lastError = st->lastSO3Error; lastCount = st->lastSO3Count;
            memcpy(lastResultR, resultR, sizeof resultR);
            float delta[3]; orc_ldlt_solve3f(jtj, jtr, delta);
            double dd[3] = { delta[0], delta[1], delta[2] }, upd[9];
            orc_rodrigues(dd, upd);
            float updf[9], nr[9];
            for (int k = 0; k < 9; ++k) updf[k] = (float)upd[k];
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
                nr[i * 3 + j] = updf[i * 3] * R_lr[j] + updf[i * 3 + 1] * R_lr[3 + j] + updf[i * 3 + 2] * R_lr[6 + j];
            memcpy(R_lr, nr, sizeof nr);
            for (int k = 0; k < 9; ++k) resultR[k] = R_lr[k];
        }
    }

    int iterations[ORC_NUM_PYRS];
    iterations[0] = opt->fastOdom ? 3 : 10;
    iterations[1] = opt->pyramid ? 5 : 0;
    iterations[2] = opt->pyramid ? 4 : 0;

    float Rprev_inv[9]; inv3f(Rprev, Rprev_inv);

    double resultRt[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
    if (opt->so3) for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) resultRt[i * 4 + j] = resultR[i * 3 + j];

    orc_icp_opts io = { opt->use_search, opt->search_radius, opt->if_curvature_info, o->distThres, o->angleThres };

    for (int i = ORC_NUM_PYRS - 1; i >= 0; --i) {
        const int rows = o->height >> i, cols = o->width >> i;
        orc_cam kl = cam_level(o->intr, i);
        if (rgb) orc_projectToPointCloud(rows, cols, o->lastDepth[i], o->cloud[i], kl);
        double K[9] = { kl.fx, 0, kl.cx, 0, kl.fy, kl.cy, 0, 0, 1 }, Kinv[9];
        inv3d(K, Kinv);
        st->lastRGBError = FLT_MAX;

        for (int j = 0; j < iterations[i]; ++j) {
            /* Rt = resultRt^-1 (rigid inverse in fp64), KRK^-1, K t  (:983-992) */
            double Rm[9], tm[3];
            for (int a = 0; a < 3; ++a) for (int b2 = 0; b2 < 3; ++b2) Rm[a * 3 + b2] = resultRt[a * 4 + b2];
            double Rinv[9]; inv3d(Rm, Rinv);
            for (int a = 0; a < 3; ++a)
                tm[a] = -(Rinv[a * 3] * resultRt[3] + Rinv[a * 3 + 1] * resultRt[7] + Rinv[a * 3 + 2] * resultRt[11]);
            double KR[9], KRK[9]; mul3d(K, Rinv, KR); mul3d(KR, Kinv, KRK);
            float krkInv[9]; for (int k = 0; k < 9; ++k) krkInv[k] = (float)KRK[k];
            float kt[3];
            for (int a = 0; a < 3; ++a) kt[a] = (float)(K[a * 3] * tm[0] + K[a * 3 + 1] * tm[1] + K[a * 3 + 2] * tm[2]);

            int sigma = 0, rgbSize = 0;
            if (rgb)
                orc_computeRgbResidual(rows, cols, (float)(pow(o->minGrad[i], 2.0) / pow(o->sobelScale, 2.0)),
                                       o->dIdx[i], o->dIdy[i], o->lastDepth[i], o->nextDepth[i],
                                       o->lastImage[i], o->nextImage[i], o->corresImg[i], o->maxDepthDeltaRGB,
                                       kt, krkInv, &sigma, &rgbSize);
            /* :1017 operator-precedence quirk: sqrt(((float)sigma / rgbSize == 0) ? 1 : rgbSize) */
            float sigmaVal = sqrtf(((float)sigma / rgbSize == 0) ? 1 : rgbSize);
            float rgbError = sqrtf((float)sigma) / (rgbSize == 0 ? 1 : rgbSize);
            if (opt->rgbOnly && rgbError > st->lastRGBError) break;
            st->lastRGBError = rgbError; st->lastRGBCount = (float)rgbSize;
            if (opt->rgbOnly) sigmaVal = -1;

            float A_icp[36] = { 0 }, b_icp[6] = { 0 }, A_rgbd[36] = { 0 }, b_rgbd[6] = { 0 }, residual[2] = { 0, 0 };
            if (icp) {
                orc_icpStep(rows, cols, Rcurr, tcurr, o->maps[M_VC][i], o->maps[M_NC][i], o->maps[M_K1C][i], o->maps[M_K2C][i],
                            Rprev_inv, tprev, kl, o->maps[M_VG][i], o->maps[M_NG][i], o->maps[M_K1G][i], o->maps[M_K2G][i],
                            o->maps[M_W][i], &io, A_icp, b_icp, residual, NULL, NULL);
                st->icp_iterations_run++;
            }
            st->lastICPError = sqrtf(residual[0]) / residual[1];
            st->lastICPCount = residual[1];
            if (rgb)
                orc_rgbStep(rows, cols, o->corresImg[i], sigmaVal, o->cloud[i], kl.fx, kl.fy, o->dIdx[i], o->dIdy[i],
                            opt->rgb_grad_weight, o->sobelScale, A_rgbd, b_rgbd, NULL);

            double lastA[36], lastb[6], result[6];
            if (icp && rgb) {
                double w = opt->icpWeight;
                for (int k = 0; k < 36; ++k) lastA[k] = (double)A_rgbd[k] + w * w * (double)A_icp[k];
                for (int k = 0; k < 6; ++k) lastb[k] = (double)b_rgbd[k] + w * (double)b_icp[k];
            } else if (icp) {
                for (int k = 0; k < 36; ++k) lastA[k] = A_icp[k];
                for (int k = 0; k < 6; ++k) lastb[k] = b_icp[k];
            } else {
                for (int k = 0; k < 36; ++k) lastA[k] = A_rgbd[k];
                for (int k = 0; k < 6; ++k) lastb[k] = b_rgbd[k];
            }
            orc_ldlt_solve6(lastA, lastb, result);
            memcpy(st->lastA, lastA, sizeof lastA); memcpy(st->lastb, lastb, sizeof lastb);

            /* OdometryProvider.h:71-93 : resultRt = exp(xi) * resultRt, rgbOdom = float(resultRt) */
            float iso[16];
            orc_computeUpdateSE3(resultRt, result, iso);

            /* :1196-1204 : currentT = [Rprev|tprev] * rgbOdom^-1, all in float; an
             * Isometry3f inverse is the transpose of the float-cast rotation */
            float Rf[9], tf[3];
            for (int a = 0; a < 3; ++a) { for (int b2 = 0; b2 < 3; ++b2) Rf[a * 3 + b2] = iso[a * 4 + b2]; tf[a] = iso[a * 4 + 3]; }
            float ti[3];
            for (int a = 0; a < 3; ++a) ti[a] = -(Rf[0 * 3 + a] * tf[0] + Rf[1 * 3 + a] * tf[1] + Rf[2 * 3 + a] * tf[2]);
            for (int a = 0; a < 3; ++a) {
                for (int b2 = 0; b2 < 3; ++b2)
                    Rcurr[a * 3 + b2] = Rprev[a * 3] * Rf[b2 * 3] + Rprev[a * 3 + 1] * Rf[b2 * 3 + 1] + Rprev[a * 3 + 2] * Rf[b2 * 3 + 2];
                tcurr[a] = Rprev[a * 3] * ti[0] + Rprev[a * 3 + 1] * ti[1] + Rprev[a * 3 + 2] * ti[2] + tprev[a];
            }
        }
    }

    if (rgb) {   /* :1232-1236 */
        float d[3] = { tcurr[0] - tprev[0], tcurr[1] - tprev[1], tcurr[2] - tprev[2] };
        if (norm3(d) > 0.3) { memcpy(Rcurr, Rprev, sizeof Rcurr); memcpy(tcurr, tprev, sizeof tcurr); }
    }
    if (opt->so3)   /* :1239-1245 */
        for (int i = 0; i < ORC_NUM_PYRS; ++i) { unsigned char* t = o->lastNextImage[i]; o->lastNextImage[i] = o->nextImage[i]; o->nextImage[i] = t; }

    memcpy(trans, tcurr, sizeof tcurr); memcpy(rot, Rcurr, sizeof Rcurr);
}
