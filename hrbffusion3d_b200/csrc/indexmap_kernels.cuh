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
    pdl_wait();
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
// out_mask: which of the optional maps are produced (kSplatColorTime | kSplatNormRad | kSplatCurv); index and vertConf always are.
// The frame pipeline's splats before fuse / clean skip the maps those passes never read (data.vert, copy_unstable.vert).
enum { kSplatColorTime = 1, kSplatNormRad = 2, kSplatCurv = 4, kSplatAll = 7 };
__global__ void __launch_bounds__(256) splat_gather_kernel(const float4* __restrict__ surfels, SplatArgs a, unsigned long long* __restrict__ keys,
                                                           unsigned int* __restrict__ index, float4* __restrict__ vertConf, float4* __restrict__ colorTime,
                                                           float4* __restrict__ normRad, float4* __restrict__ curvMax, float4* __restrict__ curvMin, int out_mask)
{
    pdl_wait();
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.cols * a.rows) return;
    const unsigned long long key = keys[k];
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (key == kEmptyKey) {
        index[k] = 0u; vertConf[k] = z4;
        if (out_mask & kSplatColorTime) colorTime[k] = z4;
        if (out_mask & kSplatNormRad) normRad[k] = z4;
        if (out_mask & kSplatCurv) { curvMax[k] = z4; curvMin[k] = z4; }
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
    const float4 pos = __ldg(s);
    const float3 P = rigid_apply(Ri, ti, pos.x, pos.y, pos.z);
    index[k] = id;
    vertConf[k] = make_float4(P.x, P.y, P.z, pos.w);
    if (out_mask & kSplatColorTime) colorTime[k] = __ldg(s + 1);
    if (out_mask & kSplatNormRad) {
        const float4 nr = __ldg(s + 2);
        const float3 n = rot_apply(Ri, nr.x, nr.y, nr.z);
        const float inv = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(n.x, n.x), __fmul_rn(n.y, n.y)), __fmul_rn(n.z, n.z))));
        normRad[k] = make_float4(__fmul_rn(n.x, inv), __fmul_rn(n.y, inv), __fmul_rn(n.z, inv), nr.w);
    }
    if (out_mask & kSplatCurv) { curvMax[k] = __ldg(s + 3); curvMin[k] = __ldg(s + 4); }
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
    const unsigned long long* row_lut;   // device copy of make_pred_row_lut(): [7][128] candidate masks of a window row
    unsigned int* dense_count;   // optional: += 1 for every texel centre (20 i + 10, 20 j + 10) of the 1/20 grid with a predicted surface
                                 // (the sample set of HRBFFusion::denseEnough, HRBFFusion.cpp:974-987, Shaders/Resize.cpp:106-139)
};

constexpr int kPredTileW = 16, kPredTileH = 4, kPredHalo = 3;
constexpr int kPredSW = kPredTileW + 2 * kPredHalo, kPredSH = kPredTileH + 2 * kPredHalo;
constexpr int kPredLanes = 4, kPredSlots = 8;            // 4 lanes x 8 slots = 32 >= max neighbours (maxN + 16)
constexpr int kPredCand = 49;

// candidate order of predict_hrbf.frag:75-80 (rings i = 0..3, x outer, y inner, perimeter only) and, for each
// candidate, the index of the first candidate of the NEXT x column (where the shader's `break` resumes)
constexpr int kPredCols = 16;     // x columns over the four rings: 1 + 3 + 5 + 7
struct PredTable {
    unsigned long long col_mask[kPredCols];      // candidates of each column (contiguous index ranges, in scan order)
    signed char dx[kPredCand], dy[kPredCand];
    unsigned char next_col[kPredCand], col_of[kPredCand];
    unsigned char ring_end[4];
};
__constant__ PredTable c_pred;

inline PredTable make_pred_table()
{
    PredTable t;
    int c = 0, col = 0;
    for (int i = 0; i <= 3; ++i) {
        for (int dx = -i; dx <= i; ++dx) {
            const int col_start = c;
            for (int dy = -i; dy <= i; ++dy) {
                if (!(dx == -i || dy == -i || dx == i || dy == i)) continue;
                t.dx[c] = (signed char)dx; t.dy[c] = (signed char)dy; ++c;
            }
            t.col_mask[col] = 0ull;
            for (int k = col_start; k < c; ++k) { t.next_col[k] = (unsigned char)c; t.col_of[k] = (unsigned char)col; t.col_mask[col] |= 1ull << k; }
            ++col;
        }
        t.ring_end[i] = (unsigned char)c;
    }
    return t;
}

// One neighbour as the ray march sees it.  Every sample point lies on the pixel's viewing ray, p(t) = c0 + t r, so with
// w0 = c0 - c_i:   |p - c_i|^2 / rho_i^2 = a + t (b + t c)      (a = |w0|^2/rho^2, b = 2 w0.r/rho^2, c = |r|^2/rho^2)
//                  (p - c_i) . s_i        = cs + t ds            (s_i = (200/rho_i^2) n_i, cs = w0.s, ds = r.s)
// f(p(t)) = -sum_i grad phi_i . (10 n_i) = sum_i (1 - sqrt(u_i))_+^3 (cs_i + t ds_i)   (hrbfbase.glsl:20-34,126-145):
// 9 operations per neighbour and evaluation, 5 registers per neighbour (40 for the 8 slots: no spills at 80 registers).
// An empty slot has a = 4 (outside every support) and contributes exactly 0.
// Two neighbours (slots 2k, 2k+1) side by side: the evaluation runs on packed fp32 pairs (fma.rn.f32x2 -- FFMA2 / FMUL2 in SASS), which
// on sm_100 halves the issue slots of the arithmetic (the kernel is issue-bound: 77 % issue-active, the FMA pipe itself has room).
struct NbRay2 { float2 a, b, c, cs, ds; };

// Window row r (dy = r - 3) with validity bits m (bit k: dx = k - 3) -> the mask of valid candidates in the shader's scan order.
// A pixel's 49-candidate mask is the OR of 7 table entries instead of 49 tests.
inline void make_pred_row_lut(unsigned long long* lut /* [7][128] */)
{
    const PredTable t = make_pred_table();
    for (int r = 0; r < 7; ++r)
        for (int m = 0; m < 128; ++m) {
            unsigned long long v = 0ull;
            for (int c = 0; c < kPredCand; ++c)
                if (t.dy[c] == r - 3 && ((m >> (t.dx[c] + 3)) & 1)) v |= 1ull << c;
            lut[r * 128 + m] = v;
        }
}

__device__ __forceinline__ float sqrt_approx(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

__device__ __forceinline__ float hrbf_ray_value(const NbRay2 (&nb)[kPredSlots / 2], bool upper_half, float t)
{
    const float2 t2 = make_float2(t, t), one2 = make_float2(1.0f, 1.0f), neg2 = make_float2(-1.0f, -1.0f);
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < kPredSlots / 2; ++k) {
        if (k >= kPredSlots / 4 && !upper_half) break;              // warp-uniform
        const float2 u = __ffma2_rn(t2, __ffma2_rn(t2, nb[k].c, nb[k].b), nb[k].a);
        const float2 sq = make_float2(sqrt_approx(u.x), sqrt_approx(u.y));
        float2 q = __ffma2_rn(sq, neg2, one2);                      // 1 - sqrt(u)
        q.x = fmaxf(q.x, 0.0f); q.y = fmaxf(q.y, 0.0f);
        const float2 q3 = __fmul2_rn(__fmul2_rn(q, q), q);
        acc = __ffma2_rn(q3, __ffma2_rn(t2, nb[k].ds, nb[k].cs), acc);
    }
    return acc.x + acc.y;
}

// hrbfbase.glsl:147-166 (+ getWeightH :37-69) at point p, over this lane's neighbours (slot s -> s_sel[s * 4 + sub]), summed
// over the 4 lanes of the pixel.  Runs once per found pixel: the neighbours are re-read from the shared-memory tile.
template <typename CenterAt>
__device__ __forceinline__ float3 hrbf_gradient_group(CenterAt nb_at, int nslots, float px, float py, float pz, unsigned gmask, int sub,
                                                      float& best, int& besti)
{
    float gx = 0.f, gy = 0.f, gz = 0.f;
    for (int s = 0; s < nslots; ++s) {
        float4 v, n;
        nb_at(s, v, n);
        const float vx = px - v.x, vy = py - v.y, vz = pz - v.z;
        const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)), __fmul_rn(vz, vz));
        {   // nearest neighbour of the root: smallest distance, ties -> lowest rank
            const float d = sqrtf(d2);
            if (d < best) { best = d; besti = s * kPredLanes + sub; }
        }
        const float T2 = __fmul_rn(n.w, n.w);
        if (d2 > T2) continue;
        const float nx = 10.0f * n.x, ny = 10.0f * n.y, nz = 10.0f * n.z;
        if (d2 == 0.0f) {
            const float h = -20.0f / T2;
            gx -= nx * h; gy -= ny * h; gz -= nz * h;
            continue;
        }
        const float r = sqrtf(d2 / T2);
        const float q = 1.0f - r;
        const float t1 = 20.0f * (q * q) / (T2 * T2 * r);
        const float t2 = -r * q * T2;
        const float h0 = t1 * (3.0f * vx * vx + t2), h1 = t1 * 3.0f * vx * vy, h2 = t1 * 3.0f * vx * vz;
        const float h4 = t1 * (3.0f * vy * vy + t2), h5 = t1 * 3.0f * vy * vz, h8 = t1 * (3.0f * vz * vz + t2);
        gx -= nx * h0 + ny * h1 + nz * h2;
        gy -= nx * h1 + ny * h4 + nz * h5;
        gz -= nx * h2 + ny * h5 + nz * h8;
    }
    gx += __shfl_xor_sync(gmask, gx, 1); gy += __shfl_xor_sync(gmask, gy, 1); gz += __shfl_xor_sync(gmask, gz, 1);
    gx += __shfl_xor_sync(gmask, gx, 2); gy += __shfl_xor_sync(gmask, gy, 2); gz += __shfl_xor_sync(gmask, gz, 2);
    return make_float3(gx, gy, gz);
}

// 256 threads = 64 pixels (16 x 4 tile) x 4 lanes.  The (16+6) x (4+6) halo tile of the two maps the
// ray march needs (position+confidence, normal+radius) is staged in shared memory once per CTA.
#ifndef HRBF_PRED_MINBLOCKS
#define HRBF_PRED_MINBLOCKS 4
#endif
__global__ void __launch_bounds__(256, HRBF_PRED_MINBLOCKS) predict_hrbf_kernel(PredictArgs a)
{
    pdl_wait();
    __shared__ float4 s_v[kPredSH][kPredSW];
    __shared__ float4 s_n[kPredSH][kPredSW];
    __shared__ unsigned char s_sel[64][32];              // per pixel: tile cell (row * kPredSW + col) of each selected neighbour
    __shared__ unsigned int s_rowmask[kPredSH];          // per tile row: bit sx = the cell passes the (pixel-independent) neighbour tests
    if (threadIdx.x < kPredSH) s_rowmask[threadIdx.x] = 0u;
    __syncthreads();
    __shared__ PredTable s_tab;                          // the candidate tables: per-lane indices would serialise in the constant cache
    static_assert(sizeof(PredTable) % 4 == 0, "copied as words");
    for (int t = threadIdx.x; t < (int)(sizeof(PredTable) / 4); t += blockDim.x)
        reinterpret_cast<unsigned int*>(&s_tab)[t] = reinterpret_cast<const unsigned int*>(&c_pred)[t];

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
        // predict_hrbf.frag:94-97, evaluated once per cell instead of once per (pixel, candidate)
        const float nl2 = __fadd_rn(__fadd_rn(__fmul_rn(n.x, n.x), __fmul_rn(n.y, n.y)), __fmul_rn(n.z, n.z));     // length < 0.1 <=> length^2 < 0.01
        if (!(v.z < 0.1f || nl2 < 0.01f || v.w < a.confThr || n.z < 0.0f)) atomicOr(&s_rowmask[sy], 1u << sx);
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, sub = lane & 3;
    const unsigned gmask = 0xFu << (lane & ~3);
    const int grp = threadIdx.x >> 2;                   // pixel slot of this 4-lane group (also its s_sel row)
    // a warp's 8 pixels form a compact 4 x 2 patch (similar surfaces -> similar trip counts of the ray march)
    const int wrp = threadIdx.x >> 5, gin = (threadIdx.x >> 2) & 7;
    const int lx = 4 * (wrp & 3) + (gin & 3), ly = 2 * (wrp >> 2) + (gin >> 2);
    const int px = tx0 + lx, py = ty0 + ly;
    const bool inside = px < a.cols && py < a.rows;     // uniform within the 4-lane group

    // ---- neighbour gather (predict_hrbf.frag:74-113) ----
    const int ncand = s_tab.ring_end[a.win];
    unsigned long long valid = 0ull;
    // (the 7-KB table is read through L1: staging it in shared memory cost every 64-pixel CTA 900 loads + stores, a tenth of the kernel)
    for (int r = sub; r < 7; r += kPredLanes) valid |= __ldg(a.row_lut + r * 128 + ((s_rowmask[ly + r] >> lx) & 127u));      // cells outside the image are invalid (z = 0)
    valid &= (1ull << ncand) - 1ull;                     // rings beyond `win` are not scanned
    valid |= __shfl_xor_sync(gmask, valid, 1);
    valid |= __shfl_xor_sync(gmask, valid, 2);
    // The shader appends valid candidates in scan order and its `break` only leaves the innermost (y) loop: once more
    // than maxN are collected, every further x column still contributes its FIRST valid candidate.  In mask form: all valid
    // bits up to the (maxN+1)-th one, plus the lowest valid bit of each later column (all 4 lanes compute it redundantly).
    unsigned long long sel = valid;
    if (__popcll(valid) > a.maxN + 1) {
        const unsigned int lo32 = (unsigned int)valid, hi32 = (unsigned int)(valid >> 32);
        const int nlo = __popc(lo32);
        const int cstar = (nlo > a.maxN) ? (int)__fns(lo32, 0, a.maxN + 1) : 32 + (int)__fns(hi32, 0, a.maxN + 1 - nlo);
        sel = valid & ((2ull << cstar) - 1ull);
        for (int k = s_tab.col_of[cstar] + 1; k < kPredCols; ++k) {
            const unsigned long long m = valid & s_tab.col_mask[k];
            sel |= m & (0ull - m);
        }
    }
    if (!inside) sel = 0ull;
    int N = __popcll(sel);
    if (N > 32) N = 32;                                 // cannot happen for maxN <= 16, win <= 3 (host-checked)
    // neighbour of rank j (in scan order) -> lane j & 3, slot j >> 2.  Every lane ranks the candidates it tested itself.
    for (unsigned long long mine = sel & (0x1111111111111111ull << sub); mine != 0ull; mine &= mine - 1ull) {      // only the set bits of this lane's quarter
        const int c = __ffsll((long long)mine) - 1;
        const int j = __popcll(sel & ((1ull << c) - 1ull));
        if (j < 32) s_sel[grp][j] = (unsigned char)((ly + kPredHalo + s_tab.dy[c]) * kPredSW + lx + kPredHalo + s_tab.dx[c]);
    }
    __syncwarp(gmask);

    const int nslots = (N - sub + kPredLanes - 1) / kPredLanes;       // slots s with s*4+sub < N
    auto nb_at = [&](int sl, float4& v, float4& n) {
        const int cell = s_sel[grp][sl * kPredLanes + sub];
        v = (&s_v[0][0])[cell];
        n = (&s_n[0][0])[cell];
    };

    // ---- viewing ray through the pixel centre (:42-50) ----
    const float xl = ((float)px + 0.5f - a.cx) * a.icx, yl = ((float)py + 0.5f - a.cy) * a.icy;
    const float rl = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(xl, xl), __fmul_rn(yl, yl)), 1.0f));
    const float rx = xl / rl, ry = yl / rl, rz = 1.0f / rl;
    const float rr = fmaf(rz, rz, fmaf(ry, ry, rx * rx));

    // closest projection onto the ray (:134-142)
    float projmin = 1000000.0f;
    for (int sl = 0; sl < nslots; ++sl) {
        float4 v, n;
        nb_at(sl, v, n);
        const float pj = fabsf(__fadd_rn(__fadd_rn(__fmul_rn(v.x, rx), __fmul_rn(v.y, ry)), __fmul_rn(v.z, rz)));
        projmin = fminf(projmin, pj);
    }
    projmin = fminf(projmin, __shfl_xor_sync(gmask, projmin, 1));
    projmin = fminf(projmin, __shfl_xor_sync(gmask, projmin, 2));
    const float c0x = projmin * rx, c0y = projmin * ry, c0z = projmin * rz;

    // per-neighbour ray coefficients (registers)
    NbRay2 nb[kPredSlots / 2];
    int cnt0 = 0;                                        // neighbours whose support contains the start point c0 (:152-157)
#pragma unroll
    for (int sl = 0; sl < kPredSlots; ++sl) {
        float na = 4.0f, nbb = 0.f, nc = 0.f, ncs = 0.f, nds = 0.f;
        if (sl < nslots) {
            float4 v, n;
            nb_at(sl, v, n);
            const float T2 = __fmul_rn(n.w, n.w), iT2 = 1.0f / T2;
            const float wx = c0x - v.x, wy = c0y - v.y, wz = c0z - v.z;
            const float d2 = fmaf(wz, wz, fmaf(wy, wy, wx * wx));
            const float k = 200.0f * iT2;
            na = d2 * iT2;
            nbb = 2.0f * fmaf(wz, rz, fmaf(wy, ry, wx * rx)) * iT2;
            nc = rr * iT2;
            ncs = k * fmaf(wz, n.z, fmaf(wy, n.y, wx * n.x));
            nds = k * fmaf(rz, n.z, fmaf(ry, n.y, rx * n.x));
            cnt0 += !(T2 < d2) ? 1 : 0;
        }
        NbRay2& d = nb[sl >> 1];
        if (sl & 1) { d.a.y = na; d.b.y = nbb; d.c.y = nc; d.cs.y = ncs; d.ds.y = nds; }
        else { d.a.x = na; d.b.x = nbb; d.c.x = nc; d.cs.x = ncs; d.ds.x = nds; }
    }

    // ---- interval search + bisection (:152-270) as ONE warp-convergent state machine over the ray parameter t ----
    // The shader runs three data-dependent loops in sequence (coarse march of 4 mm steps, fine march of 0.4 mm steps
    // back, bisection); run as written, a warp pays the SUM of its pixels' worst trip counts with ~60 % of the lanes
    // idle.  Here every lane evaluates f once per iteration at the t its pixel's state asks for, so a warp pays the
    // MAXIMUM of its pixels' total evaluation counts and the evaluation itself is executed convergently.
    // Root search, result-equivalent to the shader's three loops (:152-270) within its own final resolution:
    //   coarse : as written -- 4-mm steps from the start point until f changes sign (at most 24)
    //   fine   : the shader walks back from the crossing in 0.4-mm steps until the sign flips back (<= 10 evaluations);
    //            here the same 0.4-mm bracket is located by bisecting the step index (<= 4 evaluations: identical bracket
    //            whenever f has a single crossing inside the 4-mm step)
    //   root   : the shader bisects the bracket to 6 um (6 evaluations) and returns the last midpoint; here one
    //            regula-falsi evaluation + a secant step land within ~0.1 um of the root, i.e. inside that final interval
    // Typical pixel: 8 evaluations instead of 14, and nearly the same count for every pixel of a warp.
    enum { ST_COARSE = 0, ST_FINE, ST_ROOT, ST_DONE };
    int nmax = nslots;                                   // warp-uniform slot bound
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, m));
    const bool upper_half = nmax > kPredSlots / 2;
    // the first evaluation, at the start point itself (every pixel with enough neighbours does exactly this one)
    float v0 = hrbf_ray_value(nb, upper_half, 0.0f);
    v0 += __shfl_xor_sync(0xffffffffu, v0, 1); cnt0 += __shfl_xor_sync(0xffffffffu, cnt0, 1);
    v0 += __shfl_xor_sync(0xffffffffu, v0, 2); cnt0 += __shfl_xor_sync(0xffffffffu, cnt0, 2);
    const float dir = v0 > 0.f ? -1.0f : 1.0f;           // v0 > 0: search backward for f < 0; else forward for f > 0
    // i = 0 of the shader's coarse march re-evaluates the start point: never a sign change, skipped
    int state = (inside && N > a.minN && cnt0 > a.minN) ? ST_COARSE : ST_DONE;
    int step = 1, lo = 0, hi = 10;
    float tb = 0.f;                                      // coarse crossing
    float f_lo = 0.f, f_hi = v0;                         // f at fine step lo (crossed side) / hi (start side)
    float tm = 0.f;
    bool found = false;
    while (__any_sync(0xffffffffu, state != ST_DONE)) {
        // 1. the ray parameter this pixel's state asks for
        float t = 0.f, t_lo = 0.f, t_hi = 0.f;
        if (state == ST_COARSE) t = 0.004f * (float)step * dir;
        else if (state == ST_FINE) t = tb - 0.0004f * (float)((lo + hi) >> 1) * dir;
        else if (state == ST_ROOT) {
            t_lo = tb - 0.0004f * (float)lo * dir; t_hi = tb - 0.0004f * (float)hi * dir;
            const float den = f_lo - f_hi;
            t = den != 0.f ? t_lo + (t_hi - t_lo) * (f_lo / den) : t_lo;
        }
        // 2. f(c0 + t r): this lane's neighbours, reduced over the 4 lanes of the pixel
        float value = hrbf_ray_value(nb, upper_half, t);
        value += __shfl_xor_sync(0xffffffffu, value, 1);
        value += __shfl_xor_sync(0xffffffffu, value, 2);
        // 3. state transition
        if (state == ST_COARSE) {
            if (v0 > 0.f ? (value < 0.f) : (value > 0.f)) { tb = t; f_lo = value; lo = 0; hi = 10; state = ST_FINE; }
            else { f_hi = value; if (++step >= 25) state = ST_DONE; }
        } else if (state == ST_FINE) {
            if (v0 > 0.f ? (value > 0.f) : (value < 0.f)) { hi = (lo + hi) >> 1; f_hi = value; }      // flipped back: the bracket ends here
            else { lo = (lo + hi) >> 1; f_lo = value; }
            if (hi - lo == 1) state = ST_ROOT;
        } else if (state == ST_ROOT) {
            // secant step through (t, value) and the bracket end of opposite sign, clamped to the bracket
            const bool same_as_lo = (value < 0.f) == (f_lo < 0.f);
            const float to = same_as_lo ? t_hi : t_lo, fo = same_as_lo ? f_hi : f_lo;
            const float den = value - fo;
            float t2 = den != 0.f ? t - value * ((t - to) / den) : t;
            t2 = fminf(fmaxf(t2, fminf(t_lo, t_hi)), fmaxf(t_lo, t_hi));
            tm = t2; found = true; state = ST_DONE;
        }
    }
    const float tx = fmaf(tm, rx, c0x), ty = fmaf(tm, ry, c0y), tz = fmaf(tm, rz, c0z);
    // ---- gradient at the root (hrbfbase.glsl:147-166) and attributes of the nearest neighbour (:273-303), one pass ----
    float3 g = make_float3(0.f, 0.f, 0.f);
    float best = 1000000.f;
    int besti = 0x7fffffff;
    if (found) g = hrbf_gradient_group(nb_at, nslots, tx, ty, tz, gmask, sub, best, besti);
    if (found) {
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
        const int cell = s_sel[grp][besti];
        const int qy = ty0 + cell / kPredSW - kPredHalo, qx = tx0 + cell % kPredSW - kPredHalo;
        const size_t q = (size_t)qy * a.cols + qx;
        const float4 v = (&s_v[0][0])[cell];
        const float4 n = (&s_n[0][0])[cell];
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
    if (a.dense_count != nullptr && vout.z > 0.f && px % 20 == 10 && py % 20 == 10 && px / 20 < a.cols / 20 && py / 20 < a.rows / 20) atomicAdd(a.dense_count, 1u);
    a.image[o] = img; a.vertex[o] = vout; a.normal[o] = nout; a.ocurvMax[o] = kmax; a.ocurvMin[o] = kmin;
    a.time[o] = tstamp; a.icpw[o] = icpw;
}

}  // namespace hrbf
