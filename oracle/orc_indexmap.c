/*
 * oracle/orc_indexmap.c -- CPU ORACLE (test infrastructure, never shipped) for SURVEY.md
 * section 8 rows 6-7: the surfel index-map splat and the per-pixel HRBF ray-cast prediction.
 *
 * Restates the GLSL passes
 *   Core/src/Shaders/index_map.vert:34-66, index_map.frag:35-43   (driver IndexMap.cpp:193-267)
 *   Core/src/Shaders/predict_hrbf.frag:40-311, hrbfbase.glsl:7-166, utils.glsl:11-15,
 *   color.glsl:19-34                                              (driver IndexMap.cpp:413-518)
 * in plain C.  GL fixed-function behaviour is DEFINED here with integer semantics (SURVEY 8a
 * "parity hazards"): a GL point covers the pixel floor(window x,y); depth test keeps the nearest
 * z, ties go to the lowest surfel id (GL_LESS + in-order rasterisation); GL_NEAREST fetches texel
 * (px+dx, py+dy); float loop counters with epsilon become integer ring offsets.
 * Parity unpinned: the reference holds no fixture for these passes.
 */
#include "orc.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

static void rigid_inverse(const float pose[16], float Ri[9], float ti[3])
{   /* pose.inverse() of a rigid transform: R^T, -R^T t (float) */
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Ri[i * 3 + j] = pose[j * 4 + i];
    for (int i = 0; i < 3; ++i) ti[i] = -(Ri[i * 3] * pose[3] + Ri[i * 3 + 1] * pose[7] + Ri[i * 3 + 2] * pose[11]);
}

/* IndexMap.cpp:193-267, index_map.vert:34-66 */
void orc_predictIndices(const float pose[16], const float* surfels, int count,
                        const orc_splat_params* p, const float* active_kf, int kf_dim,
                        uint32_t* index, float* vertConf, float* colorTime, float* normRad,
                        float* curvMax, float* curvMin)
{
    const int W = p->cols, H = p->rows;
    const size_t P = (size_t)W * H;
    float Ri[9], ti[3];
    rigid_inverse(pose, Ri, ti);
    float* zbuf = (float*)malloc(P * sizeof(float));
    for (size_t i = 0; i < P; ++i) zbuf[i] = INFINITY;
    memset(index, 0, P * sizeof(uint32_t));
    memset(vertConf, 0, 4 * P * sizeof(float)); memset(colorTime, 0, 4 * P * sizeof(float));
    memset(normRad, 0, 4 * P * sizeof(float)); memset(curvMax, 0, 4 * P * sizeof(float)); memset(curvMin, 0, 4 * P * sizeof(float));

    for (int id = 0; id < count; ++id) {
        const float* s = surfels + (size_t)id * 20;
        float X = Ri[0] * s[0] + Ri[1] * s[1] + Ri[2] * s[2] + ti[0];
        float Y = Ri[3] * s[0] + Ri[4] * s[1] + Ri[5] * s[2] + ti[1];
        float Z = Ri[6] * s[0] + Ri[7] * s[1] + Ri[8] * s[2] + ti[2];
        /* uint(vColorTime.y) indexes the active-keyframe mask (index_map.vert:41-43) */
        float sub = s[5];
        int kf = (sub >= 0.0f && sub < (float)kf_dim) ? (int)sub : -1;
        float active = (kf >= 0) ? active_kf[kf] : 0.0f;
        if (Z > p->maxDepth || Z < 0 || active == 0.0f) continue;     /* :45-51 */
        if (!(Z < p->maxDepth)) continue;                             /* depth == 1.0 fails GL_LESS against the clear value */
        float xw = (p->fx * X) / Z + p->cx, yw = (p->fy * Y) / Z + p->cy;   /* :54-55, window coords */
        if (!(xw >= 0.0f && xw < (float)W && yw >= 0.0f && yw < (float)H)) continue;   /* clipped */
        int px = (int)floorf(xw), py = (int)floorf(yw);
        size_t k = (size_t)py * W + px;
        if (!(Z < zbuf[k])) continue;                                 /* GL_LESS; earlier id wins ties */
        zbuf[k] = Z;
        index[k] = (uint32_t)id;
        float* o = vertConf + 4 * k; o[0] = X; o[1] = Y; o[2] = Z; o[3] = s[3];
        memcpy(colorTime + 4 * k, s + 4, 16);
        float nx = Ri[0] * s[8] + Ri[1] * s[9] + Ri[2] * s[10];
        float ny = Ri[3] * s[8] + Ri[4] * s[9] + Ri[5] * s[10];
        float nz = Ri[6] * s[8] + Ri[7] * s[9] + Ri[8] * s[10];
        float inv = 1.0f / sqrtf(nx * nx + ny * ny + nz * nz);
        o = normRad + 4 * k; o[0] = nx * inv; o[1] = ny * inv; o[2] = nz * inv; o[3] = s[11];
        memcpy(curvMax + 4 * k, s + 12, 16);
        memcpy(curvMin + 4 * k, s + 16, 16);
    }
    free(zbuf);
}

/* ---- hrbfbase.glsl ---- */
#define MAXNB 100

/* hrbfbase.glsl:20-34 (getWeightD) folded into :126-145 (hrbfvalue) */
static float hrbf_value(const float p[3], const float (*vc)[4], const float (*nr)[4], int n, int* support_count)
{
    float value = 0;
    int cnt = 0;
    for (int i = 0; i < n; ++i) {
        float sx = 10.0f * nr[i][0], sy = 10.0f * nr[i][1], sz = 10.0f * nr[i][2];
        float vx = p[0] - vc[i][0], vy = p[1] - vc[i][1], vz = p[2] - vc[i][2];
        float d2 = vx * vx + vy * vy + vz * vz;
        float support = nr[i][3];
        if (support * support < d2) continue;
        float T2 = support * support;
        float gx = 0, gy = 0, gz = 0;
        if (!(d2 > T2 || d2 == 0.0f)) {
            float invT2 = 1.0f / T2;
            float r = sqrtf(d2 * invT2);
            float s = 1.0f - r;
            float s3 = s * s * s;
            float t = -20 * s3 * invT2;
            gx = vx * t; gy = vy * t; gz = vz * t;
        }
        value -= gx * sx + gy * sy + gz * sz;
        cnt++;
    }
    *support_count = cnt;
    return value;
}

/* hrbfbase.glsl:37-69 (getWeightH) + :147-166 (hrbfgradient) */
static void hrbf_gradient(const float p[3], const float (*vc)[4], const float (*nr)[4], int n, float g[3])
{
    g[0] = g[1] = g[2] = 0;
    for (int i = 0; i < n; ++i) {
        float sx = 10.0f * nr[i][0], sy = 10.0f * nr[i][1], sz = 10.0f * nr[i][2];
        float vx = p[0] - vc[i][0], vy = p[1] - vc[i][1], vz = p[2] - vc[i][2];
        float d2 = vx * vx + vy * vy + vz * vz;
        float support = nr[i][3];
        float T2 = support * support;
        float h[9];
        if (d2 > T2) { for (int k = 0; k < 9; ++k) h[k] = 0; }
        else if (d2 == 0.0f) { for (int k = 0; k < 9; ++k) h[k] = 0; h[0] = h[4] = h[8] = -20.0f / T2; }
        else {
            float r = sqrtf(d2 / T2);
            float s = 1.0f - r;
            float s2 = s * s;
            float t1 = 20.0f * s2 / (T2 * T2 * r);
            float t2 = -r * s * T2;
            h[0] = t1 * (3.0f * vx * vx + t2); h[1] = t1 * 3.0f * vx * vy; h[2] = t1 * 3.0f * vx * vz;
            h[3] = h[1]; h[4] = t1 * (3.0f * vy * vy + t2); h[5] = t1 * 3.0f * vy * vz;
            h[6] = h[2]; h[7] = h[5]; h[8] = t1 * (3.0f * vz * vz + t2);
        }
        g[0] -= sx * h[0] + sy * h[1] + sz * h[2];
        g[1] -= sx * h[3] + sy * h[4] + sz * h[5];
        g[2] -= sx * h[6] + sy * h[7] + sz * h[8];
    }
}

/* predict_hrbf.frag:40-311 */
void orc_predictHRBF(const orc_predict_params* p,
                     const float* vertConf, const float* colorTime, const float* normRad,
                     const float* curvMax, const float* curvMin,
                     unsigned char* image, float* vertex, float* normal,
                     float* ocurvMax, float* ocurvMin, unsigned short* time, float* icp_weight)
{
    const int W = p->cols, H = p->rows;
    /* uniform cam = (cx, cy, 1/fx, 1/fy), IndexMap.cpp:449-452 (1.0 / fx evaluated in double, stored as float) */
    const float icx = (float)(1.0 / (double)p->fx), icy = (float)(1.0 / (double)p->fy);

#pragma omp parallel for schedule(dynamic, 4)
    for (int py = 0; py < H; ++py) {
        for (int px = 0; px < W; ++px) {
            const size_t o = (size_t)py * W + px;
            float x = (float)px + 0.5f, y = (float)py + 0.5f;             /* :42-43 */
            float xl = (x - p->cx) * icx, yl = (y - p->cy) * icy;
            float rl = sqrtf(xl * xl + yl * yl + 1.0f);
            float ray[3] = { xl / rl, yl / rl, 1.0f / rl };

            float vc[MAXNB][4], nr[MAXNB][4];
            size_t src[MAXNB];
            int N = 0;
            /* :74-113 ring-by-ring gather; `break` leaves only the innermost (y) loop */
            for (int i = 0; i <= p->win; ++i)
                for (int dx = -i; dx <= i; ++dx)
                    for (int dy = -i; dy <= i; ++dy) {
                        if (!(dx == -i || dy == -i || dx == i || dy == i)) continue;
                        int qx = px + dx, qy = py + dy;
                        if (qx < 0 || qx >= W || qy < 0 || qy >= H) continue;
                        size_t q = (size_t)qy * W + qx;
                        const float* v = vertConf + 4 * q;
                        const float* n = normRad + 4 * q;
                        float nl = sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
                        if (v[2] < 0.1f || nl < 0.1f || v[3] < p->confThreshold || n[2] < 0.0f) continue;
                        if (N < MAXNB) { memcpy(vc[N], v, 16); memcpy(nr[N], n, 16); src[N] = q; }
                        N++;
                        if (N > p->maxNeighbors) break;
                    }
            if (N > MAXNB) N = MAXNB;   /* cannot happen for win <= 3 (the shader's arrays hold 100) */

            float image_v[4] = { 0, 0, 0, 0 }, p_surface[3] = { 0, 0, 0 }, p_normal[3] = { 0, 0, 0 };
            float cmaxv[4] = { 0, 0, 0, 1000.0f }, cminv[4] = { 0, 0, 0, 1000.0f };
            float icpw = 0, confidence = 0, radius = 0;
            unsigned short tstamp = 0;
            float p_temp[3] = { 0, 0, 0 }, normal_temp[3] = { 0, 0, 0 };
            float start[3] = { 0, 0, 0 }, end[3] = { 0, 0, 0 }, closest[3] = { 0, 0, 0 };
            float projmin = 1000000.0f;
            for (int i = 0; i < N; ++i) {                                  /* :134-142 */
                float pj = fabsf(vc[i][0] * ray[0] + vc[i][1] * ray[1] + vc[i][2] * ray[2]);
                if (pj < projmin) { closest[0] = pj * ray[0]; closest[1] = pj * ray[1]; closest[2] = pj * ray[2]; projmin = pj; }
            }
            int find_interval = 0, found_surface = 0, nsp = 0;
            if (N > p->minNeighbors) {                                     /* :152-230 */
                float v0 = hrbf_value(closest, vc, nr, N, &nsp);
                if (nsp > p->minNeighbors) {
                    int dummy;
                    if (v0 > 0) {
                        int sfound = 0;
                        memcpy(end, closest, sizeof end);
                        for (int i = 0; i < 25; ++i) {
                            float tt = 0.004f * (float)i;
                            float p1[3] = { end[0] - tt * ray[0], end[1] - tt * ray[1], end[2] - tt * ray[2] };
                            float v1 = hrbf_value(p1, vc, nr, N, &dummy);
                            if (v1 < 0) { memcpy(start, p1, sizeof start); sfound = 1; break; }
                        }
                        if (sfound)
                            for (int i = 1; i < 11; ++i) {
                                float tt = 0.0004f * (float)i;
                                float p2[3] = { start[0] + tt * ray[0], start[1] + tt * ray[1], start[2] + tt * ray[2] };
                                float v2 = hrbf_value(p2, vc, nr, N, &dummy);
                                if (v2 > 0) { memcpy(end, p2, sizeof end); find_interval = 1; break; }
                            }
                    } else {
                        int efound = 0;
                        memcpy(start, closest, sizeof start);
                        for (int i = 0; i < 25; ++i) {
                            float tt = 0.004f * (float)i;
                            float p1[3] = { start[0] + tt * ray[0], start[1] + tt * ray[1], start[2] + tt * ray[2] };
                            float v1 = hrbf_value(p1, vc, nr, N, &dummy);
                            if (v1 > 0) { memcpy(end, p1, sizeof end); efound = 1; break; }
                        }
                        if (efound)
                            for (int i = 1; i < 11; ++i) {
                                float tt = 0.0004f * (float)i;
                                float p2[3] = { end[0] - tt * ray[0], end[1] - tt * ray[1], end[2] - tt * ray[2] };
                                float v2 = hrbf_value(p2, vc, nr, N, &dummy);
                                if (v2 < 0) { memcpy(start, p2, sizeof start); find_interval = 1; break; }
                            }
                    }
                }
            }
            if (find_interval) {                                           /* :234-270 */
                int dummy;
                for (int j = 0; j < 10; ++j) {
                    float st[3] = { end[0] - start[0], end[1] - start[1], end[2] - start[2] };
                    if (sqrtf(st[0] * st[0] + st[1] * st[1] + st[2] * st[2]) < 0.00001f) {
                        memcpy(p_surface, p_temp, sizeof p_surface);
                        hrbf_gradient(p_surface, vc, nr, N, normal_temp);
                        found_surface = 1; break;
                    }
                    p_temp[0] = start[0] + 0.5f * st[0]; p_temp[1] = start[1] + 0.5f * st[1]; p_temp[2] = start[2] + 0.5f * st[2];
                    float f = hrbf_value(p_temp, vc, nr, N, &dummy);
                    if (fabsf(f) < 0.00001f) {
                        memcpy(p_surface, p_temp, sizeof p_surface);
                        hrbf_gradient(p_surface, vc, nr, N, normal_temp);
                        found_surface = 1; break;
                    }
                    if (f < 0) memcpy(start, p_temp, sizeof start); else memcpy(end, p_temp, sizeof end);
                }
            }
            if (found_surface) {                                           /* :273-303 */
                memcpy(p_surface, p_temp, sizeof p_surface);
                float nl = sqrtf(normal_temp[0] * normal_temp[0] + normal_temp[1] * normal_temp[1] + normal_temp[2] * normal_temp[2]);
                p_normal[0] = normal_temp[0] / nl; p_normal[1] = normal_temp[1] / nl; p_normal[2] = normal_temp[2] / nl;
                float dmin = 1000000;
                for (int it = 0; it < N; ++it) {
                    float ddx = p_surface[0] - vc[it][0], ddy = p_surface[1] - vc[it][1], ddz = p_surface[2] - vc[it][2];
                    float dist = sqrtf(ddx * ddx + ddy * ddy + ddz * ddz);
                    if (dist < dmin) {
                        memcpy(cmaxv, curvMax + 4 * src[it], 16); memcpy(cminv, curvMin + 4 * src[it], 16);
                        confidence = vc[it][3]; radius = nr[it][3];
                        const float* ct = colorTime + 4 * src[it];
                        int c = (int)ct[0];                                /* color.glsl:27-34 decodeColor */
                        image_v[0] = (float)((c >> 16) & 0xFF) / 255.0f; image_v[1] = (float)((c >> 8) & 0xFF) / 255.0f;
                        image_v[2] = (float)(c & 0xFF) / 255.0f; image_v[3] = 1.0f;
                        tstamp = (unsigned short)(unsigned int)ct[2];
                        dmin = dist;
                    }
                }
                float lambda = p->icpWeightLambda;
                float a1 = fabsf(cmaxv[3]), a2 = fabsf(cminv[3]);
                float cmax = a1 > a2 ? a1 : a2;
                icpw = (1.0f / (p_surface[2] * p_surface[2])) * (confidence / 256.0f + expf(-0.5f * (lambda * lambda) / (cmax * cmax)));
            }
            for (int c = 0; c < 4; ++c) {
                float q = image_v[c] * 255.0f;                              /* RGBA8 unorm store */
                image[4 * o + c] = (unsigned char)(q <= 0 ? 0 : q >= 255.0f ? 255 : (int)(q + 0.5f));
            }
            vertex[4 * o + 0] = p_surface[0]; vertex[4 * o + 1] = p_surface[1]; vertex[4 * o + 2] = p_surface[2]; vertex[4 * o + 3] = confidence;
            normal[4 * o + 0] = p_normal[0]; normal[4 * o + 1] = p_normal[1]; normal[4 * o + 2] = p_normal[2]; normal[4 * o + 3] = radius;
            memcpy(ocurvMax + 4 * o, cmaxv, 16); memcpy(ocurvMin + 4 * o, cminv, 16);
            time[o] = tstamp;
            icp_weight[o] = icpw;
        }
    }
}
