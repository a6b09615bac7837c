// indexmap_kernels.cuh -- sm_100a kernels for SURVEY.md section 8 rows 6-7:
//   * surfel -> image index-map splat (replaces the GL point rasteriser + z-buffer pass
//     IndexMap::predictIndices, Core/src/IndexMap.cpp:193-267, Shaders/index_map.vert:34-66):
//     one 64-bit atomicMin(depth_bits << 32 | surfel id) per visible surfel, then a per-pixel gather
//   * per-pixel HRBF ray-cast prediction (replaces the full-screen fragment pass
//     IndexMap::predictHRBF, IndexMap.cpp:413-518, Shaders/predict_hrbf.frag:40-311,
//     hrbfbase.glsl:7-166): 4 lanes cooperate on one pixel, neighbours live in registers.
#pragma once
#include "common.cuh"

namespace hrbf {

struct SplatArgs {
    const float* inv_pose;       // device: Ri[9], ti[3] = pose^-1 (IndexMap.cpp:207)
    float fx, fy, cx, cy;
    int cols, rows;
    float maxDepth;
    const float* active_kf;      // device float[kf_dim] 0/1 (IndexMap.cpp:222-237)
    int kf_dim;
};

constexpr unsigned long long kEmptyKey = ~0ull;

// The products/sums of the projection are written with explicit _rn intrinsics (never contracted
// to FMA) so that the pixel a surfel lands in is bit-identical to the oracle's definition.
__device__ __forceinline__ float3 rigid_apply(const float* R, const float* t, float x, float y, float z)
{
    return make_float3(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[0], x), __fmul_rn(R[1], y)), __fmul_rn(R[2], z)), t[0]),
                       __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[3], x), __fmul_rn(R[4], y)), __fmul_rn(R[5], z)), t[1]),
                       __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[6], x), __fmul_rn(R[7], y)), __fmul_rn(R[8], z)), t[2]));
}
__device__ __forceinline__ float3 rot_apply(const float* R, float x, float y, float z)
{
    return make_float3(__fadd_rn(__fadd_rn(__fmul_rn(R[0], x), __fmul_rn(R[1], y)), __fmul_rn(R[2], z)),
                       __fadd_rn(__fadd_rn(__fmul_rn(R[3], x), __fmul_rn(R[4], y)), __fmul_rn(R[5], z)),
                       __fadd_rn(__fadd_rn(__fmul_rn(R[6], x), __fmul_rn(R[7], y)), __fmul_rn(R[8], z)));
}

// index_map.vert:34-59 for one surfel -> pixel index (or -1) and camera-frame position
__device__ __forceinline__ int splat_project(const SplatArgs& a, const float* Ri, const float* ti, const float4 pos, float submap, float3* pc)
{
    const float3 P = rigid_apply(Ri, ti, pos.x, pos.y, pos.z);
    *pc = P;
    const int kf = (submap >= 0.0f && submap < (float)a.kf_dim) ? (int)submap : -1;
    const float active = kf >= 0 ? __ldg(a.active_kf + kf) : 0.0f;
    if (P.z > a.maxDepth || P.z < 0.f || active == 0.0f) return -1;
    if (!(P.z < a.maxDepth)) return -1;                       // depth 1.0 fails GL_LESS against the clear value
    const float xw = __fadd_rn(__fdiv_rn(__fmul_rn(a.fx, P.x), P.z), a.cx);
    const float yw = __fadd_rn(__fdiv_rn(__fmul_rn(a.fy, P.y), P.z), a.cy);
    if (!(xw >= 0.0f && xw < (float)a.cols && yw >= 0.0f && yw < (float)a.rows)) return -1;
    return (int)floorf(yw) * a.cols + (int)floorf(xw);
}

// pass 1: one thread per surfel, 32 B of the 80-B record are read (position + colour/time)
__global__ void __launch_bounds__(256) splat_keys_kernel(const float4* __restrict__ surfels, const unsigned int* __restrict__ count_dev,
                                                         SplatArgs a, unsigned long long* __restrict__ keys)
{
    const unsigned int count = *count_dev;
    float Ri[9], ti[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) Ri[k] = __ldg(a.inv_pose + k);
#pragma unroll
    for (int k = 0; k < 3; ++k) ti[k] = __ldg(a.inv_pose + 9 + k);
    for (unsigned int id = blockIdx.x * blockDim.x + threadIdx.x; id < count; id += blockDim.x * gridDim.x) {
        const float4 pos = __ldg(surfels + 5 * (size_t)id);
        const float submap = __ldg(reinterpret_cast<const float*>(surfels + 5 * (size_t)id + 1) + 1);
        float3 pc;
        const int k = splat_project(a, Ri, ti, pos, submap, &pc);
        if (k < 0) continue;
        // nearest z wins, ties -> lowest id (GL_LESS + in-order rasterisation); z >= 0 so its bits order like the value
        atomicMin(keys + k, ((unsigned long long)__float_as_uint(pc.z) << 32) | id);
    }
}

// pass 2: one thread per pixel; re-arms the key buffer for the next call
__global__ void __launch_bounds__(256) splat_gather_kernel(const float4* __restrict__ surfels, SplatArgs a, unsigned long long* __restrict__ keys,
                                                           unsigned int* __restrict__ index, float4* __restrict__ vertConf, float4* __restrict__ colorTime,
                                                           float4* __restrict__ normRad, float4* __restrict__ curvMax, float4* __restrict__ curvMin)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.cols * a.rows) return;
    const unsigned long long key = keys[k];
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (key == kEmptyKey) {
        index[k] = 0u; vertConf[k] = z4; colorTime[k] = z4; normRad[k] = z4; curvMax[k] = z4; curvMin[k] = z4;
        return;
    }
    keys[k] = kEmptyKey;
    float Ri[9], ti[3];
#pragma unroll
    for (int q = 0; q < 9; ++q) Ri[q] = __ldg(a.inv_pose + q);
#pragma unroll
    for (int q = 0; q < 3; ++q) ti[q] = __ldg(a.inv_pose + 9 + q);
    const unsigned int id = (unsigned int)(key & 0xffffffffull);
    const float4* s = surfels + 5 * (size_t)id;
    const float4 pos = __ldg(s), ct = __ldg(s + 1), nr = __ldg(s + 2);
    const float3 P = rigid_apply(Ri, ti, pos.x, pos.y, pos.z);
    const float3 n = rot_apply(Ri, nr.x, nr.y, nr.z);
    const float inv = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(n.x, n.x), __fmul_rn(n.y, n.y)), __fmul_rn(n.z, n.z))));
    index[k] = id;
    vertConf[k] = make_float4(P.x, P.y, P.z, pos.w);
    colorTime[k] = ct;
    normRad[k] = make_float4(__fmul_rn(n.x, inv), __fmul_rn(n.y, inv), __fmul_rn(n.z, inv), nr.w);
    curvMax[k] = __ldg(s + 3);
    curvMin[k] = __ldg(s + 4);
}

__global__ void fill_keys_kernel(unsigned long long* keys, int n)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) keys[k] = kEmptyKey;
}

// ------------------------------------------------------------------ row 6 ---
struct PredictArgs {
    const float4 *vertConf, *colorTime, *normRad, *curvMax, *curvMin;     // index maps (camera frame)
    uchar4* image; float4 *vertex, *normal, *ocurvMax, *ocurvMin; unsigned short* time; float* icpw;
    int cols, rows;
    float cx, cy, icx, icy;      // uniform cam = (cx, cy, 1/fx, 1/fy)
    int win, minN, maxN;
    float confThr, lambda;
};

constexpr int kPredTileW = 16, kPredTileH = 4, kPredHalo = 3;
constexpr int kPredSW = kPredTileW + 2 * kPredHalo, kPredSH = kPredTileH + 2 * kPredHalo;
constexpr int kPredLanes = 4, kPredSlots = 8;            // 4 lanes x 8 slots = 32 >= max neighbours (maxN + 16)
constexpr int kPredCand = 49;

// candidate order of predict_hrbf.frag:75-80 (rings i = 0..3, x outer, y inner, perimeter only) and, for each
// candidate, the index of the first candidate of the NEXT x column (where the shader's `break` resumes)
struct PredTable { signed char dx[kPredCand], dy[kPredCand]; unsigned char next_col[kPredCand]; unsigned char ring_end[4]; };
__constant__ PredTable c_pred;

inline PredTable make_pred_table()
{
    PredTable t;
    int c = 0;
    for (int i = 0; i <= 3; ++i) {
        for (int dx = -i; dx <= i; ++dx) {
            const int col_start = c;
            for (int dy = -i; dy <= i; ++dy) {
                if (!(dx == -i || dy == -i || dx == i || dy == i)) continue;
                t.dx[c] = (signed char)dx; t.dy[c] = (signed char)dy; ++c;
            }
            for (int k = col_start; k < c; ++k) t.next_col[k] = (unsigned char)c;
        }
        t.ring_end[i] = (unsigned char)c;
    }
    return t;
}

struct Nb { float cx, cy, cz, sx, sy, sz, T2, invT2; };

// hrbfbase.glsl:126-145 over this lane's slots; group-reduced over the 4 lanes of the pixel
__device__ __forceinline__ float hrbf_value_group(const Nb (&nb)[kPredSlots], int nslots, float px, float py, float pz, unsigned gmask, int* support)
{
    float value = 0.f;
    int cnt = 0;
#pragma unroll
    for (int s = 0; s < kPredSlots; ++s) {
        if (s < nslots) {
            const float vx = px - nb[s].cx, vy = py - nb[s].cy, vz = pz - nb[s].cz;
            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)), __fmul_rn(vz, vz));
            if (!(nb[s].T2 < d2)) {
                if (!(d2 > nb[s].T2 || d2 == 0.0f)) {
                    const float r = sqrtf(d2 * nb[s].invT2);
                    const float q = 1.0f - r;
                    const float t = -20.f * (q * q * q) * nb[s].invT2;
                    value -= (vx * t) * nb[s].sx + (vy * t) * nb[s].sy + (vz * t) * nb[s].sz;
                }
                ++cnt;
            }
        }
    }
    value += __shfl_xor_sync(gmask, value, 1); cnt += __shfl_xor_sync(gmask, cnt, 1);
    value += __shfl_xor_sync(gmask, value, 2); cnt += __shfl_xor_sync(gmask, cnt, 2);
    *support = cnt;
    return value;
}

// hrbfbase.glsl:147-166 (+ getWeightH :37-69)
__device__ __forceinline__ float3 hrbf_gradient_group(const Nb (&nb)[kPredSlots], int nslots, float px, float py, float pz, unsigned gmask)
{
    float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
    for (int s = 0; s < kPredSlots; ++s) {
        if (s < nslots) {
            const float vx = px - nb[s].cx, vy = py - nb[s].cy, vz = pz - nb[s].cz;
            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)), __fmul_rn(vz, vz));
            const float T2 = nb[s].T2;
            if (d2 > T2) continue;
            if (d2 == 0.0f) {
                const float h = -20.0f / T2;
                gx -= nb[s].sx * h; gy -= nb[s].sy * h; gz -= nb[s].sz * h;
                continue;
            }
            const float r = sqrtf(d2 / T2);
            const float q = 1.0f - r;
            const float t1 = 20.0f * (q * q) / (T2 * T2 * r);
            const float t2 = -r * q * T2;
            const float h0 = t1 * (3.0f * vx * vx + t2), h1 = t1 * 3.0f * vx * vy, h2 = t1 * 3.0f * vx * vz;
            const float h4 = t1 * (3.0f * vy * vy + t2), h5 = t1 * 3.0f * vy * vz, h8 = t1 * (3.0f * vz * vz + t2);
            gx -= nb[s].sx * h0 + nb[s].sy * h1 + nb[s].sz * h2;
            gy -= nb[s].sx * h1 + nb[s].sy * h4 + nb[s].sz * h5;
            gz -= nb[s].sx * h2 + nb[s].sy * h5 + nb[s].sz * h8;
        }
    }
    gx += __shfl_xor_sync(gmask, gx, 1); gy += __shfl_xor_sync(gmask, gy, 1); gz += __shfl_xor_sync(gmask, gz, 1);
    gx += __shfl_xor_sync(gmask, gx, 2); gy += __shfl_xor_sync(gmask, gy, 2); gz += __shfl_xor_sync(gmask, gz, 2);
    return make_float3(gx, gy, gz);
}

// 256 threads = 64 pixels (16 x 4 tile) x 4 lanes.  The (16+6) x (4+6) halo tile of the two maps the
// ray march needs (position+confidence, normal+radius) is staged in shared memory once per CTA.
__global__ void __launch_bounds__(256) predict_hrbf_kernel(PredictArgs a)
{
    __shared__ float4 s_v[kPredSH][kPredSW];
    __shared__ float4 s_n[kPredSH][kPredSW];
    __shared__ unsigned char s_sel[64][32];

    const int tx0 = blockIdx.x * kPredTileW, ty0 = blockIdx.y * kPredTileH;
    for (int t = threadIdx.x; t < kPredSH * kPredSW; t += blockDim.x) {
        const int sy = t / kPredSW, sx = t - sy * kPredSW;
        const int gx = tx0 + sx - kPredHalo, gy = ty0 + sy - kPredHalo;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f), n = v;          // outside the image: rejected by z < 0.1
        if (gx >= 0 && gx < a.cols && gy >= 0 && gy < a.rows) {
            v = __ldg(a.vertConf + (size_t)gy * a.cols + gx);
            n = __ldg(a.normRad + (size_t)gy * a.cols + gx);
        }
        s_v[sy][sx] = v; s_n[sy][sx] = n;
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, sub = lane & 3;
    const unsigned gmask = 0xFu << (lane & ~3);
    const int grp = threadIdx.x >> 2;                   // pixel within the tile
    const int lx = grp & (kPredTileW - 1), ly = grp >> 4;
    const int px = tx0 + lx, py = ty0 + ly;
    const bool inside = px < a.cols && py < a.rows;     // uniform within the 4-lane group

    // ---- neighbour gather (predict_hrbf.frag:74-113) ----
    const int ncand = c_pred.ring_end[a.win];
    unsigned long long valid = 0ull;
    for (int c = sub; c < ncand; c += kPredLanes) {
        const int qx = px + c_pred.dx[c], qy = py + c_pred.dy[c];
        if (qx < 0 || qx >= a.cols || qy < 0 || qy >= a.rows) continue;
        const float4 v = s_v[ly + kPredHalo + c_pred.dy[c]][lx + kPredHalo + c_pred.dx[c]];
        const float4 n = s_n[ly + kPredHalo + c_pred.dy[c]][lx + kPredHalo + c_pred.dx[c]];
        const float nl = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(n.x, n.x), __fmul_rn(n.y, n.y)), __fmul_rn(n.z, n.z)));
        if (v.z < 0.1f || nl < 0.1f || v.w < a.confThr || n.z < 0.0f) continue;
        valid |= 1ull << c;
    }
    valid |= __shfl_xor_sync(gmask, valid, 1);
    valid |= __shfl_xor_sync(gmask, valid, 2);
    // sequential emulation of the shader's append / `break` (leaves only the innermost loop)
    int N = 0;
    if (sub == 0 && inside) {
        int c = 0;
        while (c < ncand) {
            if ((valid >> c) & 1ull) {
                if (N < 32) s_sel[grp][N] = (unsigned char)c;
                ++N;
                if (N > a.maxN) { c = c_pred.next_col[c]; continue; }
            }
            ++c;
        }
    }
    N = __shfl_sync(gmask, N, lane & ~3);
    if (N > 32) N = 32;                                 // cannot happen for maxN <= 16, win <= 3 (host-checked)
    __syncwarp(gmask);

    Nb nb[kPredSlots];
    const int nslots = (N - sub + kPredLanes - 1) / kPredLanes;       // slots s with s*4+sub < N
#pragma unroll
    for (int s = 0; s < kPredSlots; ++s) {
        nb[s].cx = nb[s].cy = nb[s].cz = nb[s].sx = nb[s].sy = nb[s].sz = 0.f; nb[s].T2 = -1.f; nb[s].invT2 = 0.f;
        if (s < nslots) {
            const int c = s_sel[grp][s * kPredLanes + sub];
            const float4 v = s_v[ly + kPredHalo + c_pred.dy[c]][lx + kPredHalo + c_pred.dx[c]];
            const float4 n = s_n[ly + kPredHalo + c_pred.dy[c]][lx + kPredHalo + c_pred.dx[c]];
            nb[s].cx = v.x; nb[s].cy = v.y; nb[s].cz = v.z;
            nb[s].sx = 10.0f * n.x; nb[s].sy = 10.0f * n.y; nb[s].sz = 10.0f * n.z;
            nb[s].T2 = __fmul_rn(n.w, n.w);
            nb[s].invT2 = __fdiv_rn(1.0f, nb[s].T2);
        }
    }

    // ---- viewing ray through the pixel centre (:42-50) ----
    const float xl = ((float)px + 0.5f - a.cx) * a.icx, yl = ((float)py + 0.5f - a.cy) * a.icy;
    const float rl = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(xl, xl), __fmul_rn(yl, yl)), 1.0f));
    const float rx = xl / rl, ry = yl / rl, rz = 1.0f / rl;

    // closest projection onto the ray (:134-142)
    float projmin = 1000000.0f;
#pragma unroll
    for (int s = 0; s < kPredSlots; ++s)
        if (s < nslots) {
            const float pj = fabsf(__fadd_rn(__fadd_rn(__fmul_rn(nb[s].cx, rx), __fmul_rn(nb[s].cy, ry)), __fmul_rn(nb[s].cz, rz)));
            projmin = fminf(projmin, pj);
        }
    projmin = fminf(projmin, __shfl_xor_sync(gmask, projmin, 1));
    projmin = fminf(projmin, __shfl_xor_sync(gmask, projmin, 2));
    const float c0x = projmin * rx, c0y = projmin * ry, c0z = projmin * rz;

    // ---- interval search (:152-230) ----
    bool find_interval = false;
    float sx_ = 0.f, sy_ = 0.f, sz_ = 0.f, ex_ = 0.f, ey_ = 0.f, ez_ = 0.f;   // starting / ending point
    int sup = 0;
    if (inside && N > a.minN) {
        const float v0 = hrbf_value_group(nb, nslots, c0x, c0y, c0z, gmask, &sup);
        if (sup > a.minN) {
            const float dir = v0 > 0.f ? -1.0f : 1.0f;           // v0 > 0: search backward for f < 0; else forward for f > 0
            float ax = c0x, ay = c0y, az = c0z;                   // anchor of the coarse march
            bool coarse = false;
            float bx = 0.f, by = 0.f, bz = 0.f;
            for (int i = 0; i < 25; ++i) {
                const float tt = 0.004f * (float)i * dir;
                const float qx = ax + tt * rx, qy = ay + tt * ry, qz = az + tt * rz;
                int dummy;
                const float v1 = hrbf_value_group(nb, nslots, qx, qy, qz, gmask, &dummy);
                if (v0 > 0.f ? (v1 < 0.f) : (v1 > 0.f)) { bx = qx; by = qy; bz = qz; coarse = true; break; }
            }
            if (coarse) {
                for (int i = 1; i < 11; ++i) {
                    const float tt = -0.0004f * (float)i * dir;
                    const float qx = bx + tt * rx, qy = by + tt * ry, qz = bz + tt * rz;
                    int dummy;
                    const float v2 = hrbf_value_group(nb, nslots, qx, qy, qz, gmask, &dummy);
                    if (v0 > 0.f ? (v2 > 0.f) : (v2 < 0.f)) {
                        if (v0 > 0.f) { sx_ = bx; sy_ = by; sz_ = bz; ex_ = qx; ey_ = qy; ez_ = qz; }
                        else { ex_ = bx; ey_ = by; ez_ = bz; sx_ = qx; sy_ = qy; sz_ = qz; }
                        find_interval = true;
                        break;
                    }
                }
            }
        }
    }

    // ---- bisection (:234-270) ----
    bool found = false;
    float tx = 0.f, ty = 0.f, tz = 0.f;                 // p_temp
    float3 g = make_float3(0.f, 0.f, 0.f);
    if (find_interval) {
        for (int j = 0; j < 10; ++j) {
            const float dx = ex_ - sx_, dy = ey_ - sy_, dz = ez_ - sz_;
            if (sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz))) < 0.00001f) { found = true; break; }
            tx = sx_ + 0.5f * dx; ty = sy_ + 0.5f * dy; tz = sz_ + 0.5f * dz;
            int dummy;
            const float f = hrbf_value_group(nb, nslots, tx, ty, tz, gmask, &dummy);
            if (fabsf(f) < 0.00001f) { found = true; break; }
            if (f < 0.f) { sx_ = tx; sy_ = ty; sz_ = tz; } else { ex_ = tx; ey_ = ty; ez_ = tz; }
        }
        if (found) g = hrbf_gradient_group(nb, nslots, tx, ty, tz, gmask);
    }

    // ---- attributes of the nearest neighbour (:273-303) ----
    float best = 1000000.f;
    int besti = 0x7fffffff;
    if (found) {
#pragma unroll
        for (int s = 0; s < kPredSlots; ++s)
            if (s < nslots) {
                const float dx = tx - nb[s].cx, dy = ty - nb[s].cy, dz = tz - nb[s].cz;
                const float d = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
                if (d < best) { best = d; besti = s * kPredLanes + sub; }
            }
#pragma unroll
        for (int m = 1; m <= 2; m <<= 1) {
            const float ob = __shfl_xor_sync(gmask, best, m);
            const int oi = __shfl_xor_sync(gmask, besti, m);
            if (ob < best || (ob == best && oi < besti)) { best = ob; besti = oi; }
        }
    }
    if (sub != 0 || !inside) return;

    const size_t o = (size_t)py * a.cols + px;
    uchar4 img = make_uchar4(0, 0, 0, 0);
    float4 vout = make_float4(0.f, 0.f, 0.f, 0.f), nout = vout;
    float4 kmax = make_float4(0.f, 0.f, 0.f, 1000.0f), kmin = kmax;
    unsigned short tstamp = 0;
    float icpw = 0.f;
    if (found && besti < 32) {
        const int c = s_sel[grp][besti];
        const int qx = px + c_pred.dx[c], qy = py + c_pred.dy[c];
        const size_t q = (size_t)qy * a.cols + qx;
        const float4 v = s_v[ly + kPredHalo + c_pred.dy[c]][lx + kPredHalo + c_pred.dx[c]];
        const float4 n = s_n[ly + kPredHalo + c_pred.dy[c]][lx + kPredHalo + c_pred.dx[c]];
        const float4 ct = __ldg(a.colorTime + q);
        kmax = __ldg(a.curvMax + q); kmin = __ldg(a.curvMin + q);
        const int col = (int)ct.x;                                   // color.glsl:27-34
        img = make_uchar4((unsigned char)((col >> 16) & 0xFF), (unsigned char)((col >> 8) & 0xFF), (unsigned char)(col & 0xFF), 255);
        tstamp = (unsigned short)(unsigned int)ct.z;
        const float nl = sqrtf(g.x * g.x + g.y * g.y + g.z * g.z);
        vout = make_float4(tx, ty, tz, v.w);
        nout = make_float4(g.x / nl, g.y / nl, g.z / nl, n.w);
        const float cm = fmaxf(fabsf(kmax.w), fabsf(kmin.w));
        icpw = (1.0f / (tz * tz)) * (v.w / 256.0f + expf(-0.5f * (a.lambda * a.lambda) / (cm * cm)));
    }
    a.image[o] = img; a.vertex[o] = vout; a.normal[o] = nout; a.ocurvMax[o] = kmax; a.ocurvMin[o] = kmin;
    a.time[o] = tstamp; a.icpw[o] = icpw;
}

}  // namespace hrbf
