/*
 * oracle/orc_prep.c -- CPU ORACLE (test infrastructure, never shipped) for SURVEY.md section 8 row 10:
 * the per-frame preprocessing passes that produce the ICP "curr" maps, plus FillIn / Resize.
 *
 * Restates the GLSL passes (driver Core/src/HRBFFusion.cpp:1016-1021, 1262-1346; ComputePack targets :896-933)
 *   Shaders/depth_bilateral.frag, depth_metric_raw.frag, depth_metric_filtered.frag,
 *   depth_vertex_normal_radius.frag:23-68, geometry.glsl:40-46,90-244, surfels.glsl:19-46,
 *   depth_curvature_gradient.frag:28-142, hrbfbase.glsl:37-124,147-195, depth_update_normalrad.frag,
 *   depth_confidence_evaluation.frag, fill_vertex.frag, fill_normal.frag, fill_curvature.frag, fill_rgb.frag
 *   (Shaders/FillIn.cpp), resize.frag (Shaders/Resize.cpp) + HRBFFusion::denseEnough (HRBFFusion.cpp:974-987)
 * GL behaviour is DEFINED with integer semantics (SURVEY 8a hazards): a fragment is pixel (px,py) with
 * texcoord ((px+.5)/cols, (py+.5)/rows); GL_NEAREST fetches texel floor(u*cols) clamped to the edge
 * (Pangolin GlTexture: GL_CLAMP_TO_EDGE -- third-party, un-vendored, unpinned); float loop counters over
 * texture space become integer texel ranges [max(0,p-w), min(size-1,p+w)], x outer / y inner.  When a 7x7
 * window is clamped at the left/top border the shader's loop starts at texture coordinate 0.0 exactly, so the
 * pixel coordinate it derives (i*cols) is the INTEGER texel index there instead of index+0.5; that quirk is kept.
 * Parity unpinned: the reference holds no fixture for these passes.
 */
#include "orc.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define P_(p) ((size_t)(p)->cols * (p)->rows)

/* GLSL exp() is only specified to a few ULP and differs between drivers; the bilateral weights feed the PCA
 * normal estimation, which amplifies 1-ulp depth differences to ~1e-3 in the normal (E[x^2]-E[x]^2 cancellation,
 * geometry.glsl:163-187).  The oracle therefore DEFINES exp for this pass by a fixed sequence of IEEE fp32
 * operations (explicit fused multiply-adds in the Horner scheme and the two accumulations -- what a GPU compiler makes of the
 * shader's `a * b + c` -- and nothing else contracted): 2^(x*log2 e) with round-to-nearest range reduction and a degree-6 polynomial, relative
 * error < 4e-6 (range reduction at large |x|, where the weight is negligible) -- any implementation that repeats the sequence reproduces the filtered depth bit for bit. */
float orc_exp_bilateral(float x)
{
    if (!(x > -87.0f)) return 0.0f;
    const float t = x * 1.44269504088896341f;
    const float n = rintf(t);
    const float f = t - n;                      /* [-0.5, 0.5] */
    float p = 1.54035304e-4f;                   /* 2^f, minimax-ish Taylor coefficients ln2^k / k! */
    p = fmaf(p, f, 1.33335581e-3f);
    p = fmaf(p, f, 9.61812911e-3f);
    p = fmaf(p, f, 5.55041087e-2f);
    p = fmaf(p, f, 2.40226507e-1f);
    p = fmaf(p, f, 6.93147181e-1f);
    p = fmaf(p, f, 1.0f);
    return ldexpf(p, (int)n);
}

/* depth_bilateral.frag */
void orc_filterDepth(const orc_prep_params* p, const unsigned short* raw, float* filtered)
{
    const int W = p->cols, H = p->rows;
    const float adj = 1.0f / (p->depthFactor * 1000.0f);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const float value = (float)raw[(size_t)y * W + x] / adj;
            float out = 0.0f;
            if (!(value > p->maxD * 1000.0f || value < 300.0f)) {
                if (p->bilateral) {
                    const float ss = 0.024691358f, sc = 0.000555556f;
                    const int R = 6, D = R * 2 + 1;
                    const int tx = (x - D / 2 + D) < W ? (x - D / 2 + D) : W, ty = (y - D / 2 + D) < H ? (y - D / 2 + D) : H;
                    float sum1 = 0, sum2 = 0;
                    for (int cy = (y - D / 2 > 0 ? y - D / 2 : 0); cy < ty; ++cy)
                        for (int cx = (x - D / 2 > 0 ? x - D / 2 : 0); cx < tx; ++cx) {
                            const float tmp = (float)raw[(size_t)cy * W + cx] / adj;
                            const float space2 = ((float)x - (float)cx) * ((float)x - (float)cx) + ((float)y - (float)cy) * ((float)y - (float)cy);
                            const float color2 = (value - tmp) * (value - tmp);
                            const float weight = orc_exp_bilateral(-fmaf(color2, sc, space2 * ss));
                            sum1 = fmaf(tmp, weight, sum1);
                            sum2 += weight;
                        }
                    out = (sum1 / sum2) * adj;
                } else {
                    out = (float)raw[(size_t)y * W + x];      /* depth_guass.frag path is not restated: bilateral is the default */
                }
            }
            filtered[(size_t)y * W + x] = out;
        }
}

/* depth_metric_raw.frag, depth_metric_filtered.frag */
void orc_metriciseDepth(const orc_prep_params* p, const unsigned short* raw, const float* filtered, float* metric, float* metric_filtered)
{
    const size_t P = P_(p);
    const unsigned hi = (unsigned)(p->maxD / p->depthFactor), lo = (unsigned)(0.3f / p->depthFactor);
    const float hif = p->maxD / p->depthFactor, lof = 0.3f / p->depthFactor;
    for (size_t i = 0; i < P; ++i) {
        const unsigned v = raw[i];
        metric[i] = (v > hi || v < lo) ? 0.0f : (float)v * p->depthFactor;
        const float f = filtered[i];
        metric_filtered[i] = (f > hif || f < lof) ? 0.0f : f * p->depthFactor;
    }
}

/* surfels.glsl:19-34 ; cam.z = 1/fx, cam.w = 1/fy */
static float get_radius(float icx, float icy, float depth, float norm_z)
{
    const float meanFocal = ((1.0f / fabsf(icx)) + (1.0f / fabsf(icy))) / 2.0f;
    const float radius = (depth / meanFocal) * 1.41421356237f;
    float radius_n = radius / fabsf(norm_z);
    const float two = 2.0f * radius;
    return two < radius_n ? two : radius_n;           /* min(2r, r_n): NaN r_n -> GLSL min(x,y) = y<x?y:x -> x */
}
float orc_getRadius(float icx, float icy, float depth, float norm_z) { return get_radius(icx, icy, depth, norm_z); }

/* surfels.glsl:37-46 */
static float confidence_fn(float cx, float cy, float x, float y, float max_dist, float weighting)
{
    const float dx = x - cx, dy = y - cy;
    const float radialDist = sqrtf(dx * dx + dy * dy) / max_dist;
    return expf((-(radialDist * radialDist) / 0.72f)) * weighting;
}
float orc_confidence(float cx, float cy, float x, float y, float max_dist, float w) { return confidence_fn(cx, cy, x, y, max_dist, w); }

/* geometry.glsl:76-84 */
static void roots2(float b, float c, float r[3])
{
    float d = b * b - 4.0f * c;
    if (d < 0.0f) d = 0.0f;
    const float sd = sqrtf(d);
    r[0] = 0.0f; r[1] = 0.5f * (b + sd); r[2] = 0.5f * (b - sd);
}
/* geometry.glsl:86-160 ; m is symmetric, m[c][r] = mat[col][row] */
static void compute_roots(float m[3][3], float r[3])
{
    const float c0 = m[0][0] * m[1][1] * m[2][2] + 2.0f * m[1][0] * m[2][0] * m[2][1] - m[0][0] * m[2][1] * m[2][1]
                   - m[1][1] * m[2][0] * m[2][0] - m[2][2] * m[1][0] * m[1][0];
    const float c1 = m[0][0] * m[1][1] - m[1][0] * m[1][0] + m[0][0] * m[2][2] - m[2][0] * m[2][0] + m[1][1] * m[2][2] - m[2][1] * m[2][1];
    const float c2 = m[0][0] + m[1][1] + m[2][2];
    if (fabsf(c0) < 0.000001f) { roots2(c2, c1, r); return; }
    const float s_inv3 = 1.0f / 3.0f, s_sqrt3 = sqrtf(3.0f);
    const float c2_over_3 = c2 * s_inv3;
    float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
    if (a_over_3 > 0.0f) a_over_3 = 0.0f;
    const float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
    float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
    if (q > 0.0f) q = 0.0f;
    const float rho = sqrtf(-a_over_3);
    const float theta = atan2f(sqrtf(-q), half_b) * s_inv3;
    const float cos_theta = cosf(theta), sin_theta = sinf(theta);
    r[0] = c2_over_3 + 2.0f * rho * cos_theta;
    r[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
    r[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
    float t;
    if (r[0] >= r[1]) { t = r[0]; r[0] = r[1]; r[1] = t; }
    if (r[1] >= r[2]) {
        t = r[1]; r[1] = r[2]; r[2] = t;
        if (r[0] >= r[1]) { t = r[0]; r[0] = r[1]; r[1] = t; }
    }
    if (r[0] <= 0) roots2(c2, c1, r);
}

/* ---- shader-literal windows -------------------------------------------------------------------------------------------
 * The window loops of geometry.glsl:198-212 and depth_curvature_gradient.frag:62-75 run on FLOAT counters in texture space:
 *     for (float i = tx_min; i <= tx_max; i += indexXStep)      tx_min/max = clamp(texcoord.x -/+ indexXStep * winMultiply)
 * In fp32 the accumulated i overshoots tx_max by an ulp for many columns, and the last column of the window is then never
 * visited (e.g. 247 of the 634 interior columns at 640 px see 6 samples, not 7); the sample coordinate i * cols handed to
 * getVertex is px + 0.5 only up to that round-off.  The oracle (and the CUDA kernels, which follow it) run these LITERAL float
 * loops: that is what tests/test_oracle_vs_reference_glsl.py compares bit for bit with the reference's own shader text compiled
 * for the CPU.  orc_set_float_loops(0) switches to the INTENDED window (integer offsets -win..win, coordinates exactly qx + 0.5;
 * round 1's restatement) -- kept only so that the same test can measure what that idealisation changed (3.7e-5 of pose per frame). */
static int g_float_loops = 1;
void orc_set_float_loops(int on) { g_float_loops = on; }
int orc_get_float_loops(void) { return g_float_loops; }
/* literal mode, which texcoord a pass sees for pixel p: a full-screen fragment pass gets (p + 0.5) / n; the vertex pass of
 * GlobalModel::fuse reads it from the uv VBO, built as float(p) / n + 1.0 / (2 * n) with the sum in double (GlobalModel.cpp:87-96) */
static _Thread_local int t_uv_vbo_coords = 0;
void orc_set_uv_vbo_coords(int on) { t_uv_vbo_coords = on; }
static int texel_of(float u, int n)      /* GL_NEAREST with 8 fractional bits of fixed point, see oracle/glsl_cpu.h texel() */
{
    const float fixed = floorf(u * (float)n * 256.0f + 0.5f);
    int i = (int)floorf(fixed / 256.0f);
    return i < 0 ? 0 : (i >= n ? n - 1 : i);
}
int orc_texel_of(float u, int n) { return texel_of(u, n); }
/* one axis of the window of pixel p: texel index and the float coordinate i * n of every visited sample; returns their number */
static int float_window(int p, int n, float win, int* texels, float* coords)
{
    const float step = 1.0f / (float)n;
    const float tc = t_uv_vbo_coords ? (float)((double)((float)p / (float)n) + 1.0 / (double)(2 * (float)n)) : ((float)p + 0.5f) / (float)n;
    float lo = tc - step * win, hi = tc + step * win;
    if (lo < 0.0f) lo = 0.0f;
    if (hi > 1.0f) hi = 1.0f;
    int k = 0;
    for (float i = lo; i <= hi && k < 16; i += step) { texels[k] = texel_of(i, n); coords[k] = i * (float)n; ++k; }
    return k;
}

/* test hook: the literal window of pixel p along an axis of n texels (uv != 0: uv-VBO texcoords) */
int orc_float_window(int p, int n, float win, int uv, int* texels, float* coords)
{
    const int saved = t_uv_vbo_coords;
    t_uv_vbo_coords = uv;
    const int k = float_window(p, n, win, texels, coords);
    t_uv_vbo_coords = saved;
    return k;
}

/* geometry.glsl:190-244 getNormalPCA(vPosition (z only), texCoord of pixel (px,py), win = 3, depth map) */
void orc_getNormalPCA(const orc_prep_params* p, const float* depth, int px, int py, float vz, float n[3])
{
    const int W = p->cols, H = p->rows, win = 3;
    const float icx = (float)(1.0 / (double)p->fx), icy = (float)(1.0 / (double)p->fy);
    const int x0 = px - win < 0 ? 0 : px - win, x1 = px + win > W - 1 ? W - 1 : px + win;
    const int y0 = py - win < 0 ? 0 : py - win, y1 = py + win > H - 1 ? H - 1 : py + win;
    const int xclamped = px - win < 0 || ((float)px + 0.5f) / (float)W - (1.0f / (float)W) * 3.0f < 0.0f;
    const int yclamped = py - win < 0 || ((float)py + 0.5f) / (float)H - (1.0f / (float)H) * 3.0f < 0.0f;
    float accu[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
    int N = 0;
    n[0] = n[1] = n[2] = 0.0f;
    int txs[16], tys[16], nx = 0, ny = 0;
    float cxs[16], cys[16];
    if (g_float_loops) {
        nx = float_window(px, W, 3.0f, txs, cxs); ny = float_window(py, H, 3.0f, tys, cys);
    } else {
        for (int qx = x0; qx <= x1; ++qx, ++nx) { txs[nx] = qx; cxs[nx] = xclamped ? (float)qx : (float)qx + 0.5f; }
        for (int qy = y0; qy <= y1; ++qy, ++ny) { tys[ny] = qy; cys[ny] = yclamped ? (float)qy : (float)qy + 0.5f; }
    }
    for (int ix = 0; ix < nx; ++ix)
        for (int iy = 0; iy < ny; ++iy) {
            const float z = depth[(size_t)tys[iy] * W + txs[ix]];
            const float fx_ = cxs[ix], fy_ = cys[iy];
            const float X = (fx_ - p->cx) * z * icx, Y = (fy_ - p->cy) * z * icy;
            if (z > 0.3f && fabsf(z - vz) < 0.05f) {
                accu[0] += X * X; accu[1] += X * Y; accu[2] += X * z; accu[3] += Y * Y; accu[4] += Y * z; accu[5] += z * z;
                accu[6] += X; accu[7] += Y; accu[8] += z;
                ++N;
            }
        }
    if (N < 8) return;
    for (int k = 0; k < 9; ++k) accu[k] /= (float)N;
    float cov[3][3];   /* cov[col][row] */
    cov[0][0] = accu[0] - accu[6] * accu[6];
    cov[1][0] = accu[1] - accu[6] * accu[7];
    cov[2][0] = accu[2] - accu[6] * accu[8];
    cov[1][1] = accu[3] - accu[7] * accu[7];
    cov[2][1] = accu[4] - accu[7] * accu[8];
    cov[2][2] = accu[5] - accu[8] * accu[8];
    cov[0][1] = cov[1][0]; cov[0][2] = cov[2][0]; cov[1][2] = cov[2][1];
    float scale = cov[0][0];
    if (cov[1][0] > scale) scale = cov[1][0];
    { float b = cov[2][0] > cov[1][1] ? cov[2][0] : cov[1][1]; if (b > scale) scale = b; }
    { float b = cov[2][1] > cov[2][2] ? cov[2][1] : cov[2][2]; if (b > scale) scale = b; }
    float sm[3][3];
    for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) sm[c][r] = cov[c][r] / scale;
    float ev[3];
    compute_roots(cov, ev);
    const float eigenvalue = ev[0] * scale;
    sm[0][0] -= eigenvalue; sm[1][1] -= eigenvalue; sm[2][2] -= eigenvalue;
    const float r0[3] = { sm[0][0], sm[1][0], sm[2][0] }, r1[3] = { sm[0][1], sm[1][1], sm[2][1] }, r2[3] = { sm[0][2], sm[1][2], sm[2][2] };
    float v1[3] = { r0[1] * r1[2] - r0[2] * r1[1], r0[2] * r1[0] - r0[0] * r1[2], r0[0] * r1[1] - r0[1] * r1[0] };
    float v2[3] = { r0[1] * r2[2] - r0[2] * r2[1], r0[2] * r2[0] - r0[0] * r2[2], r0[0] * r2[1] - r0[1] * r2[0] };
    float v3[3] = { r1[1] * r2[2] - r1[2] * r2[1], r1[2] * r2[0] - r1[0] * r2[2], r1[0] * r2[1] - r1[1] * r2[0] };
    const float l1 = sqrtf(v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2]), l2 = sqrtf(v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2]),
                l3 = sqrtf(v3[0] * v3[0] + v3[1] * v3[1] + v3[2] * v3[2]);
    float* nn = (l1 >= l2 && l1 >= l3) ? v1 : (l2 >= l1 && l2 >= l3) ? v2 : v3;
    float s = 1.0f;
    if (nn[2] < 0) s = -1.0f;
    const float a = s * nn[0], b = s * nn[1], c = s * nn[2];
    const float len = sqrtf(a * a + b * b + c * c);
    n[0] = a / len; n[1] = b / len; n[2] = c / len;
}

/* depth_vertex_normal_radius.frag:23-68 (PCA path; the central-difference path needs preprocessingNormalEstimationPCA = 0) */
void orc_computeVertexNormalRadius(const orc_prep_params* p, const float* metric, const float* metric_filtered,
                                   float* vertex_raw, float* vertex_filtered, float* normal, float* radius)
{
    const int W = p->cols, H = p->rows;
    const float icx = (float)(1.0 / (double)p->fx), icy = (float)(1.0 / (double)p->fy);
    const float max_dist = sqrtf(((float)H * 0.5f) * ((float)H * 0.5f) + ((float)W * 0.5f) * ((float)W * 0.5f));
#pragma omp parallel for schedule(dynamic, 8)
    for (int py = 0; py < H; ++py)
        for (int px = 0; px < W; ++px) {
            const size_t o = (size_t)py * W + px;
            const float x = (float)px + 0.5f, y = (float)py + 0.5f;
            const float z = metric[o], zf = metric_filtered[o];
            /* getVertex(..., int(x), int(y), ...) : integer pixel coordinates */
            float v[3] = { ((float)px - p->cx) * z * icx, ((float)py - p->cy) * z * icy, z };
            float vf[3] = { ((float)px - p->cx) * zf * icx, ((float)py - p->cy) * zf * icy, zf };
            float n[3] = { 0, 0, 0 };
            if (p->pca) orc_getNormalPCA(p, metric_filtered, px, py, zf, n);
            float rad = p->radiusMultiplier * get_radius(icx, icy, vf[2], n[2]);
            const float nl = sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
            if (nl < 0.3f || v[2] < 0.3f || vf[2] < 0.3f) {
                v[0] = v[1] = v[2] = vf[0] = vf[1] = vf[2] = n[0] = n[1] = n[2] = 0.0f;
                rad = 0.0f;
            }
            float* o4 = vertex_raw + 4 * o; o4[0] = v[0]; o4[1] = v[1]; o4[2] = v[2]; o4[3] = confidence_fn(p->cx, p->cy, x, y, max_dist, 1.0f);
            o4 = vertex_filtered + 4 * o; o4[0] = vf[0]; o4[1] = vf[1]; o4[2] = vf[2]; o4[3] = 1.0f;
            o4 = normal + 4 * o; o4[0] = n[0]; o4[1] = n[1]; o4[2] = n[2]; o4[3] = rad;
            if (radius) radius[o] = rad;
        }
}

/* hrbfbase.glsl:37-69 + 147-166 over a neighbour list (vc: xyz, nr: normal xyz + support radius) */
static void hrbf_gradient49(const float p[3], const float (*vc)[4], const float (*nr)[4], int n, float g[3])
{
    g[0] = g[1] = g[2] = 0;
    for (int i = 0; i < n; ++i) {
        const float sx = 10.0f * nr[i][0], sy = 10.0f * nr[i][1], sz = 10.0f * nr[i][2];
        const float vx = p[0] - vc[i][0], vy = p[1] - vc[i][1], vz = p[2] - vc[i][2];
        const float d2 = vx * vx + vy * vy + vz * vz;
        const float T2 = nr[i][3] * nr[i][3];
        float h[9];
        if (d2 > T2) { for (int k = 0; k < 9; ++k) h[k] = 0; }
        else if (d2 == 0.0f) { for (int k = 0; k < 9; ++k) h[k] = 0; h[0] = h[4] = h[8] = -20.0f / T2; }
        else {
            const float r = sqrtf(d2 / T2), s = 1.0f - r, s2 = s * s;
            const float t1 = 20.0f * s2 / (T2 * T2 * r), t2 = -r * s * T2;
            h[0] = t1 * (3.0f * (vx * vx) + t2); h[1] = t1 * 3.0f * vx * vy; h[2] = t1 * 3.0f * vx * vz;
            h[3] = h[1]; h[4] = t1 * (3.0f * (vy * vy) + t2); h[5] = t1 * 3.0f * vy * vz;
            h[6] = h[2]; h[7] = h[5]; h[8] = t1 * (3.0f * (vz * vz) + t2);
        }
        g[0] -= sx * h[0] + sy * h[1] + sz * h[2];
        g[1] -= sx * h[3] + sy * h[4] + sz * h[5];
        g[2] -= sx * h[6] + sy * h[7] + sz * h[8];
    }
}

/* hrbfbase.glsl:72-124 (getWeightT) + 168-195 (hrbfHessianMatrix; note g[3], g[6], g[7] are copies) */
static void hrbf_hessian49(const float p[3], const float (*vc)[4], const float (*nr)[4], int n, float g[9])
{
    for (int k = 0; k < 9; ++k) g[k] = 0.0f;
    for (int i = 0; i < n; ++i) {
        const float sx = 10.0f * nr[i][0], sy = 10.0f * nr[i][1], sz = 10.0f * nr[i][2];
        const float vx = p[0] - vc[i][0], vy = p[1] - vc[i][1], vz = p[2] - vc[i][2];
        const float d2 = vx * vx + vy * vy + vz * vz;
        const float T2 = nr[i][3] * nr[i][3];
        float t[27];
        if (d2 > T2 || d2 == 0.0f) { for (int k = 0; k < 27; ++k) t[k] = 0.0f; }
        else {
            const float r = sqrtf(d2 / T2);
            const float s = 1.0f - r;
            const float s2 = r - 2 + 1 / r;
            const float s3 = 60 / (T2 * T2);
            const float s4 = 1 / (r * r);
            const float prx = vx / (T2 * r), pry = vy / (T2 * r), prz = vz / (T2 * r);
            t[0] = s3 * (T2 * s * s * prx + 2 * vx * s2 + vx * vx * (prx - s4 * prx));
            t[1] = s3 * vy * ((prx - s4 * prx) * vx + s2);
            t[2] = s3 * vz * ((prx - s4 * prx) * vx + s2);
            t[3] = s3 * (T2 * s * s * pry + vx * vx * (pry - s4 * pry));
            t[4] = s3 * vx * ((pry - s4 * pry) * vy + s2);
            t[5] = s3 * vx * vz * (pry - s4 * pry);
            t[6] = s3 * (T2 * s * s * prz + vx * vx * (prz - s4 * prz));
            t[7] = s3 * vx * vy * (prz - s4 * prz);
            t[8] = s3 * vx * ((prz - s4 * prz) * vz + s2);
            t[9] = t[1];
            t[10] = s3 * (T2 * s * s * prx + vy * vy * (prx - s4 * prx));
            t[11] = s3 * vy * vz * (prx - s4 * prx);
            t[12] = t[4];
            t[13] = s3 * (T2 * s * s * pry + 2 * vy * s2 + vy * vy * (pry - s4 * pry));
            t[14] = s3 * vz * ((pry - s4 * pry) * vy + s2);
            t[15] = t[7];
            t[16] = s3 * (T2 * s * s * prz + vy * vy * (prz - s4 * prz));
            t[17] = s3 * vy * ((prz - s4 * prz) * vz + s2);
            t[18] = t[2]; t[19] = t[11];
            t[20] = s3 * (T2 * s * s * prx + vz * vz * (prx - s4 * prx));
            t[21] = t[5]; t[22] = t[14];
            t[23] = s3 * (T2 * s * s * pry + vz * vz * (pry - s4 * pry));
            t[24] = t[8]; t[25] = t[17];
            t[26] = s3 * (T2 * s * s * prz + 2 * vz * s2 + vz * vz * (prz - s4 * prz));
        }
        g[0] -= sx * t[0] + sy * t[1] + sz * t[2];
        g[1] -= sx * t[3] + sy * t[4] + sz * t[5];
        g[2] -= sx * t[6] + sy * t[7] + sz * t[8];
        g[3] = g[1];
        g[4] -= sx * t[12] + sy * t[13] + sz * t[14];
        g[5] -= sx * t[15] + sy * t[16] + sz * t[17];
        g[6] = g[2];
        g[7] = g[5];
        g[8] -= sx * t[24] + sy * t[25] + sz * t[26];
    }
}

/* depth_curvature_gradient.frag:28-142 */
void orc_computeCurvatureGradient(const orc_prep_params* p, const float* vertex_filtered, const float* normal,
                                  float* curv1, float* curv2, float* gradient_mag, float* normal_opt)
{
    const int W = p->cols, H = p->rows, win = (int)p->curvWindow;
    const float icx = (float)(1.0 / (double)p->fx), icy = (float)(1.0 / (double)p->fy);
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = 0; py < H; ++py)
        for (int px = 0; px < W; ++px) {
            const size_t o = (size_t)py * W + px;
            const float* vf = vertex_filtered + 4 * o;
            const float* vn = normal + 4 * o;
            float kmax[4] = { 0, 0, 0, 1000.0f }, kmin[4] = { 0, 0, 0, 1000.0f }, gm = 0.0f, nopt[4] = { 0, 0, 0, 0 };
            if (vf[2] > 0.3f && sqrtf(vn[0] * vn[0] + vn[1] * vn[1] + vn[2] * vn[2]) > 0.5f) {
                float k1 = 1000.0f, k2 = 1000.0f, pmax[3] = { 0, 0, 0 }, pmin[3] = { 0, 0, 0 };
                float vc[100][4], nr[100][4];
                int N = 0;
                const int x0 = px - win < 0 ? 0 : px - win, x1 = px + win > W - 1 ? W - 1 : px + win;
                const int y0 = py - win < 0 ? 0 : py - win, y1 = py + win > H - 1 ? H - 1 : py + win;
                int txs[16], tys[16], nx = 0, ny = 0;
                if (g_float_loops) {
                    float unused[16];
                    nx = float_window(px, W, p->curvWindow, txs, unused); ny = float_window(py, H, p->curvWindow, tys, unused);
                } else {
                    for (int qx = x0; qx <= x1 && nx < 16; ++qx) txs[nx++] = qx;
                    for (int qy = y0; qy <= y1 && ny < 16; ++qy) tys[ny++] = qy;
                }
                for (int ix = 0; ix < nx; ++ix)
                    for (int iy = 0; iy < ny; ++iy) {
                        const int qx = txs[ix], qy = tys[iy];
                        const float* v = vertex_filtered + 4 * ((size_t)qy * W + qx);
                        const float* n = normal + 4 * ((size_t)qy * W + qx);
                        if (fabsf(v[2] - vf[2]) < 0.10f && v[2] > 0.3f && sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]) > 0.8f && N < 100) {
                            vc[N][0] = v[0]; vc[N][1] = v[1]; vc[N][2] = v[2]; vc[N][3] = 1.0f;
                            memcpy(nr[N], n, 16);
                            ++N;
                        }
                    }
                if (N > 15) {
                    float g[3], hs[9];
                    hrbf_gradient49(vf, vc, nr, N, g);
                    gm = fabsf(g[0] * vn[0] + g[1] * vn[1] + g[2] * vn[2]);
                    const float gl = sqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
                    nopt[0] = g[0] / gl; nopt[1] = g[1] / gl; nopt[2] = g[2] / gl; nopt[3] = vn[3];
                    (void)icx; (void)icy;
                    hrbf_hessian49(vf, vc, nr, N, hs);
                    const float gx = g[0], gy = g[1], gz = g[2];
                    const float h_x = -gx / gz, h_y = -gy / gz;
                    const float gz3 = gz * gz * gz;
                    const float h_xx = (2 * gx * gz * hs[2] - gx * gx * hs[8] - gz * gz * hs[0]) / gz3;
                    const float h_xy = (gx * gz * hs[5] + gy * gz * hs[2] - gx * gy * hs[8] - gz * gz * hs[1]) / gz3;
                    const float h_yy = (2 * gy * gz * hs[5] - gy * gy * hs[8] - gz * gz * hs[4]) / gz3;
                    const float E = 1 + h_x * h_x, F = h_x * h_y, G = 1 + h_y * h_y;
                    const float len = sqrtf(h_x * h_x + h_y * h_y + 1);
                    const float L = h_xx / len, M = h_xy / len, Nn = h_yy / len;
                    const float cg = (L * Nn - M * M) / (E * G - F * F);
                    const float cm = (E * Nn + G * L - 2 * F * M) / (2 * (E * G - F * F));
                    if (!isnan(cg) && !isnan(cm)) {
                        float delta = cm * cm - cg;
                        if (delta < 0.0f) delta = 0.0f;
                        k1 = cm + sqrtf(delta); k2 = cm - sqrtf(delta);
                        const float lmax = -(M - k1 * F) / (Nn - k1 * G), lmin = -(M - k2 * F) / (Nn - k2 * G);
                        /* r_u + lambda r_v = (1, lambda, h_x + lambda h_y), normalised */
                        float a[3] = { 1.0f, lmax, h_x + lmax * h_y }, b[3] = { 1.0f, lmin, h_x + lmin * h_y };
                        /* literal mode: the shader evaluates r_u + lambda * r_v component-wise, so an infinite lambda (degenerate
                         * direction, y and z are NaN anyway) also turns x = 1 + lambda * 0 into NaN; the default keeps x = 1 -> 0 */
                        if (g_float_loops) { a[0] = 1.0f + lmax * 0.0f; a[1] = 0.0f + lmax * 1.0f; b[0] = 1.0f + lmin * 0.0f; b[1] = 0.0f + lmin * 1.0f; }
                        const float la = sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]), lb = sqrtf(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
                        for (int k = 0; k < 3; ++k) { pmax[k] = a[k] / la; pmin[k] = b[k] / lb; }
                    }
                }
                kmax[0] = pmax[0]; kmax[1] = pmax[1]; kmax[2] = pmax[2]; kmax[3] = k1;
                kmin[0] = pmin[0]; kmin[1] = pmin[1]; kmin[2] = pmin[2]; kmin[3] = k2;
            }
            memcpy(curv1 + 4 * o, kmax, 16); memcpy(curv2 + 4 * o, kmin, 16);
            gradient_mag[o] = gm;
            memcpy(normal_opt + 4 * o, nopt, 16);
        }
}

/* depth_confidence_evaluation.frag */
void orc_vertexConfidence(const orc_prep_params* p, const float* gradient_mag, float weighting, int useConfEval, float epsilon, float* confidence)
{
    const int W = p->cols, H = p->rows;
    const float max_dist = sqrtf(((float)H * 0.5f) * ((float)H * 0.5f) + ((float)W * 0.5f) * ((float)W * 0.5f));
    for (int py = 0; py < H; ++py)
        for (int px = 0; px < W; ++px) {
            const size_t o = (size_t)py * W + px;
            float c = confidence_fn(p->cx, p->cy, (float)px + 0.5f, (float)py + 0.5f, max_dist, weighting);
            if (useConfEval > 0) c = c * expf(-epsilon / sqrtf(gradient_mag[o]));
            confidence[o] = c;
        }
}

/* fill_vertex.frag, fill_normal.frag, fill_curvature.frag, fill_rgb.frag (FillIn.cpp clears every target to 0) */
void orc_fillIn(const orc_prep_params* p, int passthrough, float lambda, float curvThr,
                const float* eVertex, const float* eIcpW, const float* eNormal, const float* eK1, const float* eK2, const unsigned char* eImage,
                const float* vertexFiltered, const float* normal, const float* k1, const float* k2, const float* confidence, const unsigned char* rgb,
                float* oVertex, float* oIcpW, float* oNormal, float* oK1, float* oK2, unsigned char* oImage)
{
    const size_t P = P_(p);
    for (size_t o = 0; o < P; ++o) {
        /* vertex + icp weight */
        const float* s = eVertex + 4 * o;
        float v[4] = { 0, 0, 0, 0 }, w = 0.0f;
        if (s[2] == 0 || passthrough == 1) {
            const float* fv = vertexFiltered + 4 * o;
            const float r1 = k1[4 * o + 3], r2 = k2[4 * o + 3];
            if (r1 > -curvThr && r1 < curvThr && r2 > -curvThr && r2 < curvThr) {
                const float vConf = confidence[o];
                const float a1 = fabsf(r1), a2 = fabsf(r2), cmax = a1 > a2 ? a1 : a2;
                w = (1.0f / (fv[2] * fv[2])) * (vConf / 256.0f + expf(-0.5f * (lambda * lambda) / (cmax * cmax)));
                v[0] = fv[0]; v[1] = fv[1]; v[2] = fv[2]; v[3] = vConf;
            }
        } else { memcpy(v, s, 16); w = eIcpW[o]; }
        memcpy(oVertex + 4 * o, v, 16); oIcpW[o] = w;
        /* normal */
        const float* en = eNormal + 4 * o;
        if (sqrtf(en[0] * en[0] + en[1] * en[1] + en[2] * en[2]) < 0.8f || passthrough == 1) memcpy(oNormal + 4 * o, normal + 4 * o, 16);
        else memcpy(oNormal + 4 * o, en, 16);
        /* curvature */
        if (eK1[4 * o + 3] > 300 || eK2[4 * o + 3] > 300 || passthrough == 1) { memcpy(oK1 + 4 * o, k1 + 4 * o, 16); memcpy(oK2 + 4 * o, k2 + 4 * o, 16); }
        else { memcpy(oK1 + 4 * o, eK1 + 4 * o, 16); memcpy(oK2 + 4 * o, eK2 + 4 * o, 16); }
        /* colour: existing RGBA8 (x+y+z == 0 -> raw RGB with alpha 255) */
        const unsigned char* ei = eImage + 4 * o;
        if ((ei[0] == 0 && ei[1] == 0 && ei[2] == 0) || passthrough == 1) { oImage[4 * o] = rgb[3 * o]; oImage[4 * o + 1] = rgb[3 * o + 1]; oImage[4 * o + 2] = rgb[3 * o + 2]; oImage[4 * o + 3] = 255; }
        else memcpy(oImage + 4 * o, ei, 4);
    }
}

/* Shaders/Resize.cpp (1/20 nearest at texel centres: texel (20i+10, 20j+10)) + HRBFFusion.cpp:974-987 */
int orc_denseEnough(int rows, int cols, const float* vertex, float thresh)
{
    const int f = 20, w = cols / f, h = rows / f;
    int sum = 0;
    for (int j = 0; j < h; ++j)
        for (int i = 0; i < w; ++i) sum += vertex[4 * ((size_t)(f * j + f / 2) * cols + (f * i + f / 2)) + 2] > 0;
    const float per = (float)sum / (float)(h * w);
    return per > thresh;
}
