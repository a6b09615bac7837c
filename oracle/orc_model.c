/*
 * oracle/orc_model.c -- CPU ORACLE (test infrastructure, never shipped) for SURVEY.md section 8 rows 8-9:
 * the surfel map's initialise / fuse / clean.
 *
 * Restates the GLSL transform-feedback passes
 *   GlobalModel::initialise  Core/src/GlobalModel.cpp:214-288, Shaders/init_unstableTex.vert:51-89, .geom
 *   GlobalModel::fuse        GlobalModel.cpp:355-549, Shaders/data.vert:63-198, data.geom, data.frag, update.vert:51-115
 *   GlobalModel::clean       GlobalModel.cpp:551-688, Shaders/copy_unstable.vert:62-166, copy_unstable.geom:37-50
 * with color.glsl:19-34, surfels.glsl, geometry.glsl, utils.glsl.  GL behaviour is DEFINED with integer
 * semantics (see orc_prep.c header): GL_NEAREST + clamp-to-edge; the half-pixel search windows
 * {-1,-1/2,0,+1/2}*(win/2) sample texel floor(x + offset); primitives are processed in buffer order, the uv
 * buffer being x outer / y inner (GlobalModel.cpp:89-96); when several input pixels update the same surfel the
 * FIRST in that order wins (all fragments at depth 0 under GL_LESS); round() is half-away-from-zero;
 * mat4*vec4 is evaluated ((m0*x + m1*y) + m2*z) + m3.
 * Parity unpinned: the reference holds no fixture for these passes.
 */
#include "orc.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

void orc_getNormalPCA(const orc_prep_params* p, const float* depth, int px, int py, float vz, float n[3]);
float orc_getRadius(float icx, float icy, float depth, float norm_z);
float orc_confidence(float cx, float cy, float x, float y, float max_dist, float w);

static void apply_pose(const float m[16], const float v[3], float o[3])
{
    for (int i = 0; i < 3; ++i) o[i] = ((m[i * 4] * v[0] + m[i * 4 + 1] * v[1]) + m[i * 4 + 2] * v[2]) + m[i * 4 + 3];
}
static void apply_rot(const float m[16], const float v[3], float o[3])
{
    for (int i = 0; i < 3; ++i) o[i] = (m[i * 4] * v[0] + m[i * 4 + 1] * v[1]) + m[i * 4 + 2] * v[2];
}
static void rigid_inverse16(const float pose[16], float inv[16])
{
    memset(inv, 0, 64);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) inv[i * 4 + j] = pose[j * 4 + i];
    for (int i = 0; i < 3; ++i) inv[i * 4 + 3] = -(inv[i * 4] * pose[3] + inv[i * 4 + 1] * pose[7] + inv[i * 4 + 2] * pose[11]);
    inv[15] = 1.0f;
}
/* color.glsl:19-25 on 8-bit inputs (c = v/255 -> round(c*255) = v) */
static float encode_rgb8(const unsigned char* c) { return (float)((((((int)c[0]) << 8) + (int)c[1]) << 8) + (int)c[2]); }
static float encode_color(const float c[3])
{
    int rgb = (int)roundf(c[0] * 255.0f);
    rgb = (rgb << 8) + (int)roundf(c[1] * 255.0f);
    rgb = (rgb << 8) + (int)roundf(c[2] * 255.0f);
    return (float)rgb;
}
static void decode_color(float c, float col[3])
{
    const int i = (int)c;
    col[0] = (float)(i >> 16 & 0xFF) / 255.0f; col[1] = (float)(i >> 8 & 0xFF) / 255.0f; col[2] = (float)(i & 0xFF) / 255.0f;
}
static int clampi(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }

static orc_prep_params prep_of(const orc_model_params* p)
{
    orc_prep_params q;
    memset(&q, 0, sizeof q);
    q.cx = p->cx; q.cy = p->cy; q.fx = p->fx; q.fy = p->fy; q.cols = p->cols; q.rows = p->rows;
    return q;
}

/* init_unstableTex.vert:51-89 + .geom */
int orc_model_initialise(const orc_model_params* p, const float pose[16],
                         const float* vertexRaw, const float* normal, const unsigned char* rgb,
                         const float* curv1, const float* curv2, const float* gradientMag, int useConfEval, float epsilon,
                         float* surfels_out)
{
    const int W = p->cols, H = p->rows;
    const float max_dist = sqrtf(((float)H * 0.5f) * ((float)H * 0.5f) + ((float)W * 0.5f) * ((float)W * 0.5f));
    int count = 0;
    for (int px = 0; px < W; ++px)
        for (int py = 0; py < H; ++py) {
            const size_t o = (size_t)py * W + px;
            const float* vl = vertexRaw + 4 * o;
            const float* nl = normal + 4 * o;
            const float* k1 = curv1 + 4 * o;
            const float* k2 = curv2 + 4 * o;
            float s[20];
            apply_pose(pose, vl, s);
            float conf = orc_confidence(p->cx, p->cy, (float)px + 0.5f, (float)py + 0.5f, max_dist, 1.0f);
            if (useConfEval > 0) conf = conf * expf(-epsilon / sqrtf(gradientMag[o]));
            s[3] = conf;
            apply_rot(pose, nl, s + 8);
            s[11] = nl[3];
            s[4] = encode_rgb8(rgb + 3 * o); s[5] = 0.0f; s[6] = 1.0f; s[7] = 1.0f;
            memcpy(s + 12, k1, 16); memcpy(s + 16, k2, 16);
            const float len = sqrtf(s[8] * s[8] + s[9] * s[9] + s[10] * s[10]);
            if (len > 0.5f && k1[3] > -p->curvThr && k1[3] < p->curvThr && k2[3] > -p->curvThr && k2[3] < p->curvThr) {
                memcpy(surfels_out + (size_t)count * 20, s, 80);
                ++count;
            }
        }
    return count;
}

/* data.vert:63-198 for one pixel.  Returns updateId (0 none, 1 merge with *best, 2 new); rec = 20 floats */
static int fuse_pixel(const orc_model_params* p, const orc_prep_params* pp, const float pose[16], int time, float indexSubmap,
                      const unsigned char* rgb, const float* depthRaw, const float* depthFiltered,
                      const float* curv1, const float* curv2, const float* confidence,
                      const uint32_t* index, const float* vertConf, const float* normRad,
                      int px, int py, float rec[20], uint32_t* best_out)
{
    const int W = p->cols, H = p->rows;
    const size_t o = (size_t)py * W + px;
    const float icx = (float)(1.0 / (double)p->fx), icy = (float)(1.0 / (double)p->fy);
    const float x = (float)px + 0.5f, y = (float)py + 0.5f;
    const float z = depthRaw[o], zf = depthFiltered[o];
    const float vloc[3] = { (x - p->cx) * z * icx, (y - p->cy) * z * icy, z };
    const float* k1 = curv1 + 4 * o;
    const float* k2 = curv2 + 4 * o;
    if (!(px % 2 == time % 2 && py % 2 == time % 2)) return 0;
    float n[3] = { 0, 0, 0 };
    if (p->pca) { orc_set_uv_vbo_coords(1); orc_getNormalPCA(pp, depthFiltered, px, py, zf, n); orc_set_uv_vbo_coords(0); }
    const float nlen = sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    if (!(nlen > 0.8f && vloc[2] > 0.3f && vloc[2] <= p->maxDepth && k1[3] > -300.0f && k1[3] < 300.0f && k2[3] > -300.0f && k2[3] < 300.0f)) return 0;

    apply_pose(pose, vloc, rec);
    rec[3] = confidence[o];
    rec[4] = encode_rgb8(rgb + 3 * o); rec[5] = indexSubmap; rec[6] = (float)time; rec[7] = 0.0f;
    apply_rot(pose, n, rec + 8);
    rec[11] = p->radiusMultiplier * orc_getRadius(icx, icy, zf, n[2]);
    memcpy(rec + 12, k1, 16); memcpy(rec + 16, k2, 16);

    const float xl = (x - p->cx) * icx, yl = (y - p->cy) * icy;
    const float lambda = sqrtf(xl * xl + yl * yl + 1);
    const float ray[3] = { xl, yl, 1.0f };
    const float raylen = sqrtf(ray[0] * ray[0] + ray[1] * ray[1] + ray[2] * ray[2]);
    const float offs[4] = { -1.0f, -0.5f, 0.0f, 0.5f };
    int counter = 0;
    float bestDist = 1000;
    uint32_t best = 0;
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) {
            const int sx = clampi((int)floorf(x + offs[a]), 0, W - 1), sy = clampi((int)floorf(y + offs[b]), 0, H - 1);
            const size_t q = (size_t)sy * W + sx;
            const uint32_t current = index[q];
            if (current > 0u) {
                const float* vc = vertConf + 4 * q;
                if (fabsf((vc[2] * lambda) - (vloc[2] * lambda)) < 0.05f) {
                    const float cr[3] = { ray[1] * vc[2] - ray[2] * vc[1], ray[2] * vc[0] - ray[0] * vc[2], ray[0] * vc[1] - ray[1] * vc[0] };
                    const float dist = sqrtf(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]) / raylen;
                    const float* nr = normRad + 4 * q;
                    const float la = sqrtf(nr[0] * nr[0] + nr[1] * nr[1] + nr[2] * nr[2]);
                    const float ang = acosf((nr[0] * n[0] + nr[1] * n[1] + nr[2] * n[2]) / (la * nlen));
                    if (dist < bestDist && (fabsf(nr[2]) < 0.75f || fabsf(ang) < 0.5f)) { counter++; bestDist = dist; best = current; }
                }
            }
        }
    if (counter > 0) { rec[7] = -1.0f; *best_out = best; return 1; }
    rec[7] = -2.0f;
    return 2;
}

/* update.vert:51-115 : merge the winning new measurement `nw` into surfel `s` */
static void merge_surfel(const float* s, const float* nw, int time, float* out)
{
    memcpy(out, s, 80);
    const float c_k = s[3], a = nw[3];
    if (nw[11] < (1.0f + 0.5f) * s[11]) {
        for (int k = 0; k < 3; ++k) out[k] = ((c_k * s[k]) + (a * nw[k])) / (c_k + a);
        out[3] = c_k + a;
        float oc[3], nc[3], avg[3];
        decode_color(s[4], oc); decode_color(nw[4], nc);
        for (int k = 0; k < 3; ++k) avg[k] = ((c_k * oc[k]) + (a * nc[k])) / (c_k + a);
        out[4] = encode_color(avg); out[5] = s[5]; out[6] = s[6]; out[7] = (float)time;
        float nr[4];
        for (int k = 0; k < 4; ++k) nr[k] = ((c_k * s[8 + k]) + (a * nw[8 + k])) / (c_k + a);
        const float len = sqrtf(nr[0] * nr[0] + nr[1] * nr[1] + nr[2] * nr[2]);
        out[8] = nr[0] / len; out[9] = nr[1] / len; out[10] = nr[2] / len; out[11] = nr[3];
        for (int k = 0; k < 4; ++k) { out[12 + k] = ((c_k * s[12 + k]) + (a * nw[12 + k])) / (c_k + a); out[16 + k] = ((c_k * s[16 + k]) + (a * nw[16 + k])) / (c_k + a); }
    } else {
        out[3] = c_k + a;
        out[7] = (float)time;
    }
}

int orc_model_fuse(const orc_model_params* p, const float pose[16], int time,
                   const unsigned char* rgb, const float* depthRaw, const float* depthFiltered,
                   const float* curv1, const float* curv2, const float* confidence,
                   const uint32_t* index, const float* vertConf, const float* colorTime, const float* normRad,
                   float indexSubmap,
                   const float* surfels_in, int count, float* surfels_out, float* unstable_out)
{
    (void)colorTime;
    const int W = p->cols, H = p->rows;
    const orc_prep_params pp = prep_of(p);
    int* winner = (int*)malloc(sizeof(int) * (size_t)(count > 0 ? count : 1));
    for (int i = 0; i < count; ++i) winner[i] = -1;
    int n_un = 0;
    for (int px = 0; px < W; ++px)
        for (int py = 0; py < H; ++py) {
            float rec[20];
            uint32_t best = 0;
            const int id = fuse_pixel(p, &pp, pose, time, indexSubmap, rgb, depthRaw, depthFiltered, curv1, curv2, confidence,
                                      index, vertConf, normRad, px, py, rec, &best);
            if (id == 0) continue;
            memcpy(unstable_out + (size_t)n_un * 20, rec, 80);
            if (id == 1 && (int)best < count && winner[best] < 0) winner[best] = n_un;   /* first fragment wins */
            ++n_un;
        }
    for (int i = 0; i < count; ++i) {
        if (winner[i] >= 0) merge_surfel(surfels_in + (size_t)i * 20, unstable_out + (size_t)winner[i] * 20, time, surfels_out + (size_t)i * 20);
        else memcpy(surfels_out + (size_t)i * 20, surfels_in + (size_t)i * 20, 80);
    }
    free(winner);
    return n_un;
}

/* copy_unstable.vert:62-166 : 1 = keep.  s is modified (vColor.w == -2 -> time) */
static int clean_test(const orc_model_params* p, const float inv[16], int time,
                      const uint32_t* index, const float* vertConf, const float* colorTime,
                      const float* active_kf, int kf_dim, float* s)
{
    const int W = p->cols, H = p->rows;
    int test = 1;
    float lp[3], ln[3];
    apply_pose(inv, s, lp);
    const float x = ((p->fx * lp[0]) / lp[2]) + p->cx, y = ((p->fy * lp[1]) / lp[2]) + p->cy;
    apply_rot(inv, s + 8, ln);
    const float nl = sqrtf(ln[0] * ln[0] + ln[1] * ln[1] + ln[2] * ln[2]);
    const float lnz = ln[2] / nl;
    int count = 0, zCount = 0;
    const float sub = s[5];
    const int kf = (sub >= 0.0f && sub < (float)kf_dim) ? (int)sub : -1;
    const float active = kf >= 0 ? active_kf[kf] : 0.0f;
    if (lp[2] < p->maxDepth && lp[2] > 0 && x > 0 && y > 0 && x < (float)W && y < (float)H) {
        /* copy_unstable.vert:106-108 walks the window with FLOAT counters in texture space,
         *     for (float i = x / cols - (scale * indexXStep * windowMultiplier); i < x / cols + (...); i += indexXStep)
         * with indexXStep = (1 / (cols * scale)) * 0.5: nominally 2 * windowMultiplier samples per axis half a pixel apart, but the
         * accumulated counter can stay an ulp below the end value and then a further sample is taken (and a sample within round-off of
         * a texel boundary falls on either side).  Literal mode runs these loops as written with GL_NEAREST texels (orc_texel_of);
         * the intended-window mode keeps round 1's integer offsets. */
        const int literal = orc_get_float_loops();
        const float scale = 1.0f, wm = (float)p->cleanWindow;
        const float stepx = (1.0f / ((float)W * scale)) * 0.5f, stepy = (1.0f / ((float)H * scale)) * 0.5f;
        int sxs[16], sys[16], nx = 0, ny = 0;
        if (literal) {
            const float i0 = x / (float)W - (scale * stepx * wm), i1 = x / (float)W + (scale * stepx * wm);
            const float j0 = y / (float)H - (scale * stepy * wm), j1 = y / (float)H + (scale * stepy * wm);
            for (float i = i0; i < i1 && nx < 16; i += stepx) sxs[nx++] = orc_texel_of(i, W);
            for (float j = j0; j < j1 && ny < 16; j += stepy) sys[ny++] = orc_texel_of(j, H);
        } else {
            const int ns = 2 * p->cleanWindow;
            for (int a = 0; a < ns; ++a) {
                sxs[nx++] = clampi((int)floorf(x + 0.5f * (float)(a - p->cleanWindow)), 0, W - 1);
                sys[ny++] = clampi((int)floorf(y + 0.5f * (float)(a - p->cleanWindow)), 0, H - 1);
            }
        }
        for (int a = 0; a < nx; ++a)
            for (int b = 0; b < ny; ++b) {
                const size_t q = (size_t)sys[b] * W + sxs[a];
                if (index[q] > 0u) {
                    const float* vc = vertConf + 4 * q;
                    const float* ct = colorTime + 4 * q;
                    const float dx = vc[0] - lp[0], dy = vc[1] - lp[1];
                    if (ct[2] < s[6] && vc[3] > p->confThreshold && vc[2] > lp[2] && vc[2] - lp[2] < 0.01f &&
                        sqrtf(dx * dx + dy * dy) < s[11] * 1.4f) count++;
                    if (ct[3] == (float)time && vc[3] > p->confThreshold && vc[2] > lp[2] && vc[2] - lp[2] > 0.01f &&
                        fabsf(lnz) > 0.85f && active > 0.0f) zCount++;
                }
            }
    }
    if (s[15] < -p->curvThr || s[15] > p->curvThr || s[19] < -p->curvThr || s[19] > p->curvThr) test = 0;
    if (count > 8 || zCount > 4) test = 0;
    if (s[7] == -2.0f) s[7] = (float)time;
    if (s[7] == -1.0f || (((float)time - s[7]) > 200.0f && s[3] < p->confThreshold)) test = 0;
    return test;
}

int orc_model_clean(const orc_model_params* p, const float pose[16], int time,
                    const uint32_t* index, const float* vertConf, const float* colorTime, const float* normRad,
                    const float* active_kf, int kf_dim,
                    const float* surfels_in, int count, const float* unstable, int n_unstable,
                    float* surfels_out)
{
    (void)normRad;
    float inv[16];
    rigid_inverse16(pose, inv);
    int n = 0;
    for (int pass = 0; pass < 2; ++pass) {
        const float* src = pass == 0 ? surfels_in : unstable;
        const int m = pass == 0 ? count : n_unstable;
        for (int i = 0; i < m; ++i) {
            float s[20];
            memcpy(s, src + (size_t)i * 20, 80);
            if (clean_test(p, inv, time, index, vertConf, colorTime, active_kf, kf_dim, s)) { memcpy(surfels_out + (size_t)n * 20, s, 80); ++n; }
        }
    }
    return n;
}

/* GlobalModel::updateModel (GlobalModel.cpp:690-767) + Shaders/update_delta_trans.vert:41-91: every surfel is moved by the
 * rigid correction of its sub-map, T = DeltaTransformKF[(int)colour.y]: position <- T * (p, 1) (confidence kept), normal <- R n
 * (radius kept); colour/time and both curvature vectors pass through unchanged (the directions are NOT rotated there).
 * delta: n_delta row-major 4x4 matrices.  In place; order and count unchanged. */
void orc_model_update(float* surfels, int count, const float* delta, int n_delta)
{
    for (int i = 0; i < count; ++i) {
        float* s = surfels + 20 * (size_t)i;
        const unsigned int sub = (unsigned int)s[5];
        if (sub >= (unsigned int)n_delta) continue;      /* texel outside the uploaded range: the texture holds its initial zeros there; restated as "no-op" */
        const float* T = delta + 16 * (size_t)sub;
        const float x = s[0], y = s[1], z = s[2], nx = s[8], ny = s[9], nz = s[10];
        s[0] = ((T[0] * x + T[1] * y) + T[2] * z) + T[3];
        s[1] = ((T[4] * x + T[5] * y) + T[6] * z) + T[7];
        s[2] = ((T[8] * x + T[9] * y) + T[10] * z) + T[11];
        s[8] = (T[0] * nx + T[1] * ny) + T[2] * nz;
        s[9] = (T[4] * nx + T[5] * ny) + T[6] * nz;
        s[10] = (T[8] * nx + T[9] * ny) + T[10] * nz;
    }
}
