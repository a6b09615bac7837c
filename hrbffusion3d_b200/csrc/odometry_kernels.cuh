// odometry_kernels.cuh -- sm_100a kernels for SURVEY.md section 8 rows 1-3 and 5:
//   * fused AoS->SoA + validity + 3-level pyramid (+ rigid transform) in ONE pass
//     (replaces copyMaps / copyCurvatureMap / copyicpWeightMap / resize*Map / tranformMaps /
//      transformCurvMaps: Core/src/Cuda/cudafuncs.cu:213-743, ~25 launches + syncs)
//   * single-pass ICP / RGB / SO3 Jacobian-product reductions with the Gauss-Newton solve in
//     the last block (replaces Core/src/Cuda/reduce.cu:253-1359 + host solve)
#pragma once
#include "gn_solve.cuh"

namespace hrbf {

// ------------------------------------------------------------------ row 5 ---
struct PyrOut {            // three pyramid levels of one SoA map
    float* p[3];
    int pitch[3];          // elements
    // the tracker's packed copy of the 17 floats ICP consumes (dense, one record per pixel, 16-byte loads):
    //   pk0 = {v.x, v.y, v.z, n.x}   pk1 = {n.y, n.z, k1.w, k2.w}
    // set on the FIRST map of a pair only (vertex map of a vertex/normal pair, k1 map of a curvature pair)
    float4* pk0[3];
    float4* pk1[3];
};

enum { PYR_VN = 0, PYR_K = 1 };

__device__ __forceinline__ void store4(float* base, int pitch, int rows, int y, int x, float a, float b, float c, float d)
{
    base[(size_t)(0 * rows + y) * pitch + x] = a;
    base[(size_t)(1 * rows + y) * pitch + x] = b;
    base[(size_t)(2 * rows + y) * pitch + x] = c;
    base[(size_t)(3 * rows + y) * pitch + x] = d;
}

// One warp = one 8x4 tile of level 0; 2x2 / 4x4 means are taken with xor-shuffles in exactly the
// reference's summation order ((x00 + x01) + x10) + x11 (cudafuncs.cu:555-577).
// KIND == PYR_VN : a = vertex texture, b = normal texture, joint validity  !(v.z==0) && n.w>0
//                  (cudafuncs.cu:367); normals re-normalised on down-sampling (:579-580).
// KIND == PYR_K  : a = k1, b = k2 textures, validity per map |k|<thr && !isnan (cudafuncs.cu:422);
//                  down-sampling validity on plane w only (:642-648).
// pose (device, R[9] t[3]) != nullptr : every level is moved to the global frame after the
//                  pyramid is built (RGBDOdometry.cpp:233-244, 747-757); vertices get +t.
// depth_out != nullptr (PYR_VN) : verticesToDepth of the raw texture (cudafuncs.cu:874-885).
// sel != nullptr && *sel != 0 : read the alternate textures (a_alt, b_alt) instead -- the device-side form of
//                  `shouldFillIn ? &fillIn.xTexture : indexMap.xTexHRBF()` (HRBFFusion.cpp:1073-1086), no host round trip.
template <int KIND>
__device__ __forceinline__ void pyr_pair_tile(const float4* __restrict__ a_aos, const float4* __restrict__ b_aos,
                                              int rows, int cols, float thr, const float* __restrict__ pose,
                                              const PyrOut& oa, const PyrOut& ob, float* __restrict__ depth_out, float depth_cutoff,
                                              const float4* __restrict__ a_alt, const float4* __restrict__ b_alt, const int* __restrict__ sel,
                                              int bx, int by)
{
    if (sel != nullptr && *sel != 0) { a_aos = a_alt; b_aos = b_alt; }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lx = lane & 7, ly = lane >> 3;
    const int x = (bx * 4 + (warp & 3)) * 8 + lx;
    const int y = (by * 2 + (warp >> 2)) * 4 + ly;
    const bool inb = x < cols && y < rows;
    const float qn = qnan();

    float R[9], t[3];
    const bool xf = pose != nullptr;
    if (xf) {
#pragma unroll
        for (int k = 0; k < 9; ++k) R[k] = __ldg(pose + k);
#pragma unroll
        for (int k = 0; k < 3; ++k) t[k] = __ldg(pose + 9 + k);
    }

    float4 a = make_float4(qn, qn, qn, qn), b = a;
    if (inb) {
        a = __ldg(a_aos + (size_t)y * cols + x);
        b = __ldg(b_aos + (size_t)y * cols + x);
        if (KIND == PYR_VN) {
            if (depth_out) depth_out[(size_t)y * cols + x] = (a.z > depth_cutoff || a.z <= 0.f) ? qn : a.z;
            const bool ok = !(a.z == 0.f) && b.w > 0.f;
            if (!ok) { a = make_float4(qn, qn, qn, qn); b = a; }
        } else {
            if (!(a.w < thr && a.w > -thr && !isnan(a.w))) a = make_float4(qn, qn, qn, qn);
            if (!(b.w < thr && b.w > -thr && !isnan(b.w))) b = make_float4(qn, qn, qn, qn);
        }
    }

    float qa[4] = { a.x, a.y, a.z, a.w }, qb[4] = { b.x, b.y, b.z, b.w };
    int lrows = rows, lcols = cols, lxx = x, lyy = y;
    bool owner = inb;          // lane that owns a pixel of the current level
#pragma unroll
    for (int L = 0; L < 3; ++L) {
        if (L > 0) {
            // 2x2 mean of the previous level; partners at xor (1,8,9) for L==1, (2,16,18) for L==2
            const int sx = (L == 1) ? 1 : 2, sy = (L == 1) ? 8 : 16;
            float na[4], nb[4];
            bool nana = false, nanb = false;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float a01 = __shfl_xor_sync(0xffffffffu, qa[c], sx), a10 = __shfl_xor_sync(0xffffffffu, qa[c], sy),
                            a11 = __shfl_xor_sync(0xffffffffu, qa[c], sx | sy);
                const float b01 = __shfl_xor_sync(0xffffffffu, qb[c], sx), b10 = __shfl_xor_sync(0xffffffffu, qb[c], sy),
                            b11 = __shfl_xor_sync(0xffffffffu, qb[c], sx | sy);
                const int vc = (KIND == PYR_VN) ? 0 : 3;   // plane deciding validity
                if (c == vc) {
                    nana = isnan(qa[c]) || isnan(a01) || isnan(a10) || isnan(a11);
                    nanb = isnan(qb[c]) || isnan(b01) || isnan(b10) || isnan(b11);
                }
                na[c] = (qa[c] + a01 + a10 + a11) / 4;
                nb[c] = (qb[c] + b01 + b10 + b11) / 4;
            }
            if (KIND == PYR_VN) {
                const float rn = 1.0f / sqrtf(nb[0] * nb[0] + nb[1] * nb[1] + nb[2] * nb[2]);
                nb[0] *= rn; nb[1] *= rn; nb[2] *= rn;
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) { qa[c] = nana ? qn : na[c]; qb[c] = nanb ? qn : nb[c]; }
            const int m = (1 << L) - 1;
            owner = owner && ((lx & m) == 0) && ((ly & m) == 0);
            lrows >>= 1; lcols >>= 1; lxx >>= 1; lyy >>= 1;
            owner = owner && lxx < lcols && lyy < lrows;
        }
        if (owner && oa.p[L]) {
            float ax = qa[0], ay = qa[1], az = qa[2], bx = qb[0], by = qb[1], bz = qb[2];
            if (xf) {
                // NaN pixels stay NaN (NaN propagates through the products)
                const float3 ra = mul(R, make_float3(ax, ay, az)), rb = mul(R, make_float3(bx, by, bz));
                if (KIND == PYR_VN) { ax = ra.x + t[0]; ay = ra.y + t[1]; az = ra.z + t[2]; }
                else { ax = ra.x; ay = ra.y; az = ra.z; }
                bx = rb.x; by = rb.y; bz = rb.z;
            }
            store4(oa.p[L], oa.pitch[L], lrows, lyy, lxx, ax, ay, az, qa[3]);
            store4(ob.p[L], ob.pitch[L], lrows, lyy, lxx, bx, by, bz, qb[3]);
            if (oa.pk1[L]) {
                const size_t o = (size_t)lyy * lcols + lxx;
                if (KIND == PYR_VN) {
                    oa.pk0[L][o] = make_float4(ax, ay, az, bx);
                    reinterpret_cast<float2*>(oa.pk1[L] + o)[0] = make_float2(by, bz);
                } else reinterpret_cast<float2*>(oa.pk1[L] + o)[1] = make_float2(qa[3], qb[3]);
            }
        }
    }
}

template <int KIND>
__global__ void __launch_bounds__(256) pyr_pair_kernel(const float4* __restrict__ a_aos, const float4* __restrict__ b_aos,
                                                       int rows, int cols, float thr, const float* __restrict__ pose,
                                                       PyrOut oa, PyrOut ob, float* __restrict__ depth_out, float depth_cutoff,
                                                       const float4* __restrict__ a_alt, const float4* __restrict__ b_alt, const int* __restrict__ sel)
{
    pyr_pair_tile<KIND>(a_aos, b_aos, rows, cols, thr, pose, oa, ob, depth_out, depth_cutoff, a_alt, b_alt, sel, blockIdx.x, blockIdx.y);
}

// icp weight: copy (w>0 else NaN, cudafuncs.cu:464) + 2 resize levels (:718-725)
__device__ __forceinline__ void pyr_weight_tile(const float* __restrict__ w_src, int rows, int cols,
                                                float* o0, int p0, float* o1, int p1, float* o2, int p2,
                                                const float* __restrict__ w_alt, const int* __restrict__ sel, int bx, int by)
{
    if (sel != nullptr && *sel != 0) w_src = w_alt;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lx = lane & 7, ly = lane >> 3;
    const int x = (bx * 4 + (warp & 3)) * 8 + lx;
    const int y = (by * 2 + (warp >> 2)) * 4 + ly;
    const bool inb = x < cols && y < rows;
    const float qn = qnan();
    float w = qn;
    if (inb) {
        const float s = __ldg(w_src + (size_t)y * cols + x);
        w = s > 0.f ? s : qn;
        o0[(size_t)y * p0 + x] = w;
    }
    {
        const float a = __shfl_xor_sync(0xffffffffu, w, 1), b = __shfl_xor_sync(0xffffffffu, w, 8), c = __shfl_xor_sync(0xffffffffu, w, 9);
        const bool bad = isnan(w) || isnan(a) || isnan(b) || isnan(c);
        w = bad ? qn : (w + a + b + c) / 4;
        if (inb && !(lx & 1) && !(ly & 1) && (x >> 1) < (cols >> 1) && (y >> 1) < (rows >> 1)) o1[(size_t)(y >> 1) * p1 + (x >> 1)] = w;
    }
    {
        const float a = __shfl_xor_sync(0xffffffffu, w, 2), b = __shfl_xor_sync(0xffffffffu, w, 16), c = __shfl_xor_sync(0xffffffffu, w, 18);
        const bool bad = isnan(w) || isnan(a) || isnan(b) || isnan(c);
        w = bad ? qn : (w + a + b + c) / 4;
        if (inb && !(lx & 3) && !(ly & 3) && (x >> 2) < (cols >> 2) && (y >> 2) < (rows >> 2)) o2[(size_t)(y >> 2) * p2 + (x >> 2)] = w;
    }
}

__global__ void __launch_bounds__(256) pyr_weight_kernel(const float* __restrict__ w_src, int rows, int cols,
                                                         float* o0, int p0, float* o1, int p1, float* o2, int p2,
                                                         const float* __restrict__ w_alt, const int* __restrict__ sel)
{
    pyr_weight_tile(w_src, rows, cols, o0, p0, o1, p1, o2, p2, w_alt, sel, blockIdx.x, blockIdx.y);
}

// ---- single-function mirrors of the reference's per-map kernels (C-ABI row 5 entry points) ----
__global__ void copy_maps_kernel(int rows, int cols, const float4* __restrict__ v, const float4* __restrict__ n,
                                 float* vd, int vp, float* nd, int np)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    float4 a = __ldg(v + (size_t)y * cols + x), b = __ldg(n + (size_t)y * cols + x);
    if (!(!(a.z == 0.f) && b.w > 0.f)) { const float q = qnan(); a = make_float4(q, q, q, q); b = a; }
    store4(vd, vp, rows, y, x, a.x, a.y, a.z, a.w);
    store4(nd, np, rows, y, x, b.x, b.y, b.z, b.w);
}
__global__ void copy_curv_kernel(int rows, int cols, const float4* __restrict__ c, float* d, int p, float thr)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    float4 a = __ldg(c + (size_t)y * cols + x);
    if (!(a.w < thr && a.w > -thr && !isnan(a.w))) { const float q = qnan(); a = make_float4(q, q, q, q); }
    store4(d, p, rows, y, x, a.x, a.y, a.z, a.w);
}
__global__ void copy_weight_kernel(int rows, int cols, const float* __restrict__ s, float* d, int p)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    const float w = __ldg(s + (size_t)y * cols + x);
    d[(size_t)y * p + x] = w > 0.f ? w : qnan();
}
// MODE 0: vmap, 1: nmap (normalise), 2: cmap (validity on w, NaN -> planes x and w), 3: single-plane weight
template <int MODE>
__global__ void resize_kernel(int drows, int dcols, int srows, const float* __restrict__ in, int ip, float* out, int op)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dcols || y >= drows) return;
    const float qn = qnan();
    const int xs = 2 * x, ys = 2 * y;
    auto ld = [&](int plane, int dy, int dx) { return __ldg(in + (size_t)(plane * srows + ys + dy) * ip + xs + dx); };
    if (MODE == 3) {
        const float a = ld(0, 0, 0), b = ld(0, 0, 1), c = ld(0, 1, 0), d = ld(0, 1, 1);
        out[(size_t)y * op + x] = (isnan(a) || isnan(b) || isnan(c) || isnan(d)) ? qn : (a + b + c + d) / 4;
        return;
    }
    const int vpl = (MODE == 2) ? 3 : 0;
    const float v00 = ld(vpl, 0, 0), v01 = ld(vpl, 0, 1), v10 = ld(vpl, 1, 0), v11 = ld(vpl, 1, 1);
    if (isnan(v00) || isnan(v01) || isnan(v10) || isnan(v11)) {
        out[(size_t)(0 * drows + y) * op + x] = qn;
        if (MODE == 2) out[(size_t)(3 * drows + y) * op + x] = qn;
        return;
    }
    float r[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) r[c] = (ld(c, 0, 0) + ld(c, 0, 1) + ld(c, 1, 0) + ld(c, 1, 1)) / 4;
    if (MODE == 1) {
        const float rn = 1.0f / sqrtf(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
        r[0] *= rn; r[1] *= rn; r[2] *= rn;
    }
    store4(out, op, drows, y, x, r[0], r[1], r[2], r[3]);
}
// MODE 0: vertices+normals (cudafuncs.cu:213-257), 1: two curvature maps (:279-322)
template <int MODE>
__global__ void transform_kernel(int rows, int cols, const float* a_src, int asp, const float* b_src, int bsp,
                                 Mat33 Rm, Vec3 tv, float* a_dst, int adp, float* b_dst, int bdp)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    const float qn = qnan();
    auto at = [&](const float* p, int pitch, int plane) { return p[(size_t)(plane * rows + y) * pitch + x]; };
    {
        const float sx = at(a_src, asp, 0);
        if (!isnan(sx)) {
            const float3 s = make_float3(sx, at(a_src, asp, 1), at(a_src, asp, 2));
            const float w = at(a_src, asp, 3);
            float3 d = mul(Rm.m, s);
            if (MODE == 0) d = d + make_float3(tv.x, tv.y, tv.z);
            store4(a_dst, adp, rows, y, x, d.x, d.y, d.z, w);
        } else a_dst[(size_t)y * adp + x] = qn;
    }
    {
        const float sx = at(b_src, bsp, 0);
        if (!isnan(sx)) {
            const float3 s = make_float3(sx, at(b_src, bsp, 1), at(b_src, bsp, 2));
            const float w = at(b_src, bsp, 3);
            const float3 d = mul(Rm.m, s);
            store4(b_dst, bdp, rows, y, x, d.x, d.y, d.z, w);
        } else b_dst[(size_t)y * bdp + x] = qn;
    }
}

// ---- GPUTest path (initICP(depth)): cudafuncs.cu:57-94, 109-136, 154-195 ----
// pyrDown of the raw depth (cudafuncs.cu:57-94): 5x5 binomial taps (6, 4, 1) / 16 per axis around source pixel (2x, 2y), clipped to
// the image, a tap counting only when it is within 3 sigma_color (= 90 raw units) of the centre; src_at(cx, cy) reads the source level
template <typename At>
__device__ __forceinline__ float depth_down_gated(int srows, int scols, int x, int y, At src_at)
{
    const float tap[3] = { 0.375f, 0.25f, 0.0625f };      // by distance from the centre
    const int sx = 2 * x, sy = 2 * y;
    const float mid = src_at(sx, sy), gate = 3.f * 30.f;
    float acc = 0.f, norm = 0.f;
    for (int cy = max(sy - 2, 0); cy < min(sy + 3, srows); ++cy)
        for (int cx = max(sx - 2, 0); cx < min(sx + 3, scols); ++cx) {
            const float d = src_at(cx, cy);
            if (fabsf(d - mid) < gate) {
                const float w = tap[abs(cx - sx)] * tap[abs(cy - sy)];
                acc += d * w;
                norm += w;
            }
        }
    return acc / norm;
}
__global__ void pyrdown_depth_kernel(int srows, int scols, const float* __restrict__ src, float* dst)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    const int drows = srows / 2, dcols = scols / 2;
    if (x >= dcols || y >= drows) return;
    dst[(size_t)y * dcols + x] = depth_down_gated(srows, scols, x, y, [&](int cx, int cy) { return src[(size_t)cy * scols + cx]; });
}
__global__ void create_vmap_kernel(int rows, int cols, const float* __restrict__ depth, float* vmap, int vp,
                                   float fx_inv, float fy_inv, float cx, float cy, float cutoff, float factor)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y * blockDim.y + threadIdx.y;
    if (u >= cols || v >= rows) return;
    const float z = depth[(size_t)v * cols + u] * factor;
    if (z != 0 && z < cutoff) store4(vmap, vp, rows, v, u, z * (u - cx) * fx_inv, z * (v - cy) * fy_inv, z, 1.0f);
    else vmap[(size_t)v * vp + u] = qnan();
}
__global__ void create_nmap_kernel(int rows, int cols, const float* __restrict__ vmap, int vp, float* nmap, int np)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y * blockDim.y + threadIdx.y;
    if (u >= cols || v >= rows) return;
    if (u == cols - 1 || v == rows - 1) { nmap[(size_t)v * np + u] = qnan(); return; }
    auto at = [&](int plane, int dy, int dx) { return vmap[(size_t)(plane * rows + v + dy) * vp + u + dx]; };
    const float a = at(0, 0, 0), b = at(0, 0, 1), c = at(0, 1, 0);
    if (!isnan(a) && !isnan(b) && !isnan(c)) {
        const float3 v00 = make_float3(a, at(1, 0, 0), at(2, 0, 0)), v01 = make_float3(b, at(1, 0, 1), at(2, 0, 1)),
                     v10 = make_float3(c, at(1, 1, 0), at(2, 1, 0));
        const float3 r = cross(v01 - v00, v10 - v00);
        const float rn = 1.0f / sqrtf(dot(r, r));
        store4(nmap, np, rows, v, u, r.x * rn, r.y * rn, r.z * rn, 1.0f);
    } else nmap[(size_t)v * np + u] = qnan();
}

// ---- RGB branch prep: cudafuncs.cu:493-524, 818-848, 898-911, 930-954, 995-1013 ----
__constant__ float c_gauss25[25] = { 1, 4, 6, 4, 1, 4, 16, 24, 16, 4, 6, 24, 36, 24, 6, 4, 16, 24, 16, 4, 1, 4, 6, 4, 1 };

// pyrDownGaussF (cudafuncs.cu:493-524) for destination pixel (x, y); src_at(cx, cy) reads the source level
template <typename At>
__device__ __forceinline__ float gauss_down_f32(int srows, int scols, int x, int y, At src_at)
{
    const int D = 5;
    const int tx = min(2 * x - D / 2 + D, scols - 1), ty = min(2 * y - D / 2 + D, srows - 1);
    float sum = 0; int count = 0;
    if (2 * x >= 2 && 2 * y >= 2 && tx == 2 * x + 3 && ty == 2 * y + 3) {
        // interior: all 25 taps, same order, the (symmetric) kernel as immediates; the weights are integers, so count += g is the reference's
        // count = (int)((float)count + g)
        constexpr int G[25] = { 1, 4, 6, 4, 1, 4, 16, 24, 16, 4, 6, 24, 36, 24, 6, 4, 16, 24, 16, 4, 1, 4, 6, 4, 1 };
#pragma unroll
        for (int j = 0; j < 5; ++j)
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const float s = src_at(2 * x - 2 + i, 2 * y - 2 + j);
                if (!isnan(s)) { sum = fmaf(s, (float)G[j * 5 + i], sum); count += G[j * 5 + i]; }
            }
        return sum / (float)count;
    }
    for (int cy = max(0, 2 * y - D / 2); cy < ty; ++cy)
        for (int cx = max(0, 2 * x - D / 2); cx < tx; ++cx) {
            const float s = src_at(cx, cy);
            if (!isnan(s)) {
                const float g = c_gauss25[(ty - cy - 1) * 5 + (tx - cx - 1)];
                sum = fmaf(s, g, sum);      // what nvcc makes of the reference's `sum += src * gauss` (cudafuncs.cu:516); the oracle does the same
                count = (int)((float)count + g);
            }
        }
    return sum / (float)count;
}
// pyrDownUcharGauss (cudafuncs.cu:818-848)
template <typename At>
__device__ __forceinline__ unsigned char gauss_down_u8(int srows, int scols, int x, int y, At src_at)
{
    const int D = 5;
    const int tx = min(2 * x - D / 2 + D, scols - 1), ty = min(2 * y - D / 2 + D, srows - 1);
    float sum = 0; int count = 0;
    if (2 * x >= 2 && 2 * y >= 2 && tx == 2 * x + 3 && ty == 2 * y + 3) {
        // interior: all 25 taps; products and sums are small integers (<= 256 * 255), exact in fp32 in any order: integer arithmetic
        constexpr int G[25] = { 1, 4, 6, 4, 1, 4, 16, 24, 16, 4, 6, 24, 36, 24, 6, 4, 16, 24, 16, 4, 1, 4, 6, 4, 1 };
        int isum = 0;
#pragma unroll
        for (int j = 0; j < 5; ++j)
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const int v = (int)src_at(2 * x - 2 + i, 2 * y - 2 + j);
                isum += v * G[j * 5 + i]; count += v > 0 ? G[j * 5 + i] : 0;
            }
        const float r = (float)isum / (float)count;
        return isnan(r) ? (unsigned char)0 : (unsigned char)(int)r;
    }
    for (int cy = max(0, 2 * y - D / 2); cy < ty; ++cy)
        for (int cx = max(0, 2 * x - D / 2); cx < tx; ++cx) {
            const unsigned char s = src_at(cx, cy);
            if (s > 0) {
                const float g = c_gauss25[(ty - cy - 1) * 5 + (tx - cx - 1)];
                sum += (float)s * g;
                count = (int)((float)count + g);
            }
        }
    const float r = sum / (float)count;
    return isnan(r) ? (unsigned char)0 : (unsigned char)(int)r;
}
__global__ void pyrdown_gauss_f32_kernel(int srows, int scols, const float* __restrict__ src, float* dst)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    const int drows = srows / 2, dcols = scols / 2;
    if (x >= dcols || y >= drows) return;
    dst[(size_t)y * dcols + x] = gauss_down_f32(srows, scols, x, y, [&](int cx, int cy) { return src[(size_t)cy * scols + cx]; });
}
__global__ void pyrdown_gauss_u8_kernel(int srows, int scols, const unsigned char* __restrict__ src, unsigned char* dst)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    const int drows = srows / 2, dcols = scols / 2;
    if (x >= dcols || y >= drows) return;
    dst[(size_t)y * dcols + x] = gauss_down_u8(srows, scols, x, y, [&](int cx, int cy) { return src[(size_t)cy * scols + cx]; });
}
// imageBGRToIntensity (cudafuncs.cu:898-911)
__device__ __forceinline__ unsigned char bgr_intensity(uchar4 s)
{
    return (unsigned char)(int)((float)s.x * 0.114f + (float)s.y * 0.299f + (float)s.z * 0.587f);
}
__global__ void rgba_to_intensity_kernel(int n, const uchar4* __restrict__ rgba, unsigned char* dst,
                                         const uchar4* __restrict__ rgba_alt, const int* __restrict__ sel)
{
    if (sel != nullptr && *sel != 0) rgba = rgba_alt;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dst[i] = bgr_intensity(__ldg(rgba + i));
}

// ---- the whole per-frame map preparation of RGBDOdometry in ONE launch (the fused frame pipeline's replacement of the
// 7 init* calls = 17 launches: RGBDOdometry.cpp:183-247, 689-794).  blockIdx.x selects a job and a tile:
//   jobs 0-3 : pyr_pair_tile  (model vertex/normal, current vertex/normal, model curvature, current curvature)
//   job  4   : pyr_weight_tile
//   jobs 5-6 : rgbd_pyramid_tile (model = "last", current = "next"): intensity + verticesToDepth + both Gauss pyramids, all
//              three levels from one 41x41 shared-memory tile (the level-1 / level-2 taps are recomputed per tile
//              instead of going through HBM with a dependent launch per level)
struct RgbdJob {
    const uchar4 *rgba, *rgba_alt;               // image texture (+ fill-in alternate)
    const unsigned char* rgb8;                   // RGB8 source used instead of rgba when non-null (the frame's upload buffer)
    const float4 *vertex, *vertex_alt;           // vertex texture verticesToDepth reads (cudafuncs.cu:874-885)
    unsigned char* img[3];
    float* depth[3];
};
struct PrepAllArgs {
    int rows, cols;
    const int* sel;                              // device fill-in decision (model jobs only), or null: decided from dense_count
    const unsigned int* dense_count; unsigned int* dense_count_reset; float dense_thresh;
    const float* pose;                           // device model pose R[9], t[3]
    float curv_thr, depth_cutoff;
    const float4 *vm, *nm, *vm_alt, *nm_alt, *vc, *nc;              // vertex / normal textures: model (+alt), current
    const float4 *k1m, *k2m, *k1m_alt, *k2m_alt, *k1c, *k2c;        // curvature textures
    const float *w, *w_alt;
    PyrOut o_vg, o_ng, o_vc, o_nc, o_k1g, o_k2g, o_k1c, o_k2c;
    float* o_w[3]; int w_pitch[3];
    RgbdJob rgbd[2];                             // [0] model ("last"), [1] current ("next")
    int pyr_bx, pyr_by, rgbd_bx, rgbd_by;        // tiles per job
    int cur_depth_only;                          // job 1: the intensity pyramid of the camera frame exists already (so3_image_kernel)
    int jobs[7];                                 // blockIdx.z -> job (gridDim.z = number of jobs): 0, 1 RGB-D pyramids (model, current), 2..5 map pairs, 6 weight
};

constexpr int kRgbdL0 = 41, kRgbdL1 = 19;        // tile edge at level 0 / 1 for an 8x8 level-2 tile

// kMode 0: intensity + depth pyramids; 1: the three intensity levels only (needs nothing but the RGB upload: the frame pipeline runs it
// first, for the SO3 pre-alignment); 2: the three depth levels only (the other half of a frame whose intensity pyramid exists already)
template <int kMode = 0>
__device__ __forceinline__ void rgbd_pyramid_tile(const RgbdJob& j, int rows, int cols, float depth_cutoff, bool use_alt, int bx, int by)
{
    constexpr bool kImage = kMode != 2, kDepth = kMode != 1;
    __shared__ unsigned char s_g0[kImage ? kRgbdL0 : 1][kImage ? kRgbdL0 + 3 : 1];
    __shared__ float s_d0[kDepth ? kRgbdL0 : 1][kDepth ? kRgbdL0 : 1];
    __shared__ unsigned char s_g1[kImage ? kRgbdL1 : 1][kImage ? kRgbdL1 + 1 : 1];
    __shared__ float s_d1[kDepth ? kRgbdL1 : 1][kDepth ? kRgbdL1 : 1];
    const uchar4* rgba = use_alt ? j.rgba_alt : j.rgba;
    const float4* vert = use_alt ? j.vertex_alt : j.vertex;
    const int rows1 = rows / 2, cols1 = cols / 2, rows2 = rows1 / 2, cols2 = cols1 / 2;
    const int X0 = 32 * bx - 6, Y0 = 32 * by - 6;            // level-0 origin of the tile
    const int X1 = 16 * bx - 2, Y1 = 16 * by - 2;            // level-1 origin
    const float qn = qnan();
    for (int t = threadIdx.x; t < kRgbdL0 * kRgbdL0; t += blockDim.x) {
        const int sy = t / kRgbdL0, sx = t - sy * kRgbdL0;
        const int gx = X0 + sx, gy = Y0 + sy;
        if (gx < 0 || gy < 0 || gx >= cols || gy >= rows) continue;
        const size_t o = (size_t)gy * cols + gx;
        const bool own = sx >= 6 && sx < 38 && sy >= 6 && sy < 38;
        if (kImage) {
            const unsigned char g = j.rgb8 != nullptr ? bgr_intensity(make_uchar4(__ldg(j.rgb8 + 3 * o), __ldg(j.rgb8 + 3 * o + 1), __ldg(j.rgb8 + 3 * o + 2), 255))
                                                      : bgr_intensity(__ldg(rgba + o));
            s_g0[sy][sx] = g;
            if (own) j.img[0][o] = g;
        }
        if (kDepth) {
            const float z = __ldg(reinterpret_cast<const float*>(vert + o) + 2);
            const float d = (z > depth_cutoff || z <= 0.f) ? qn : z;
            s_d0[sy][sx] = d;
            if (own) j.depth[0][o] = d;
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < kRgbdL1 * kRgbdL1; t += blockDim.x) {
        const int sy = t / kRgbdL1, sx = t - sy * kRgbdL1;
        const int x1 = X1 + sx, y1 = Y1 + sy;
        if (x1 < 0 || y1 < 0 || x1 >= cols1 || y1 >= rows1) continue;
        const bool own = sx >= 2 && sx < 18 && sy >= 2 && sy < 18;
        if (kImage) {
            const unsigned char g = gauss_down_u8(rows, cols, x1, y1, [&](int cx, int cy) { return s_g0[cy - Y0][cx - X0]; });
            s_g1[sy][sx] = g;
            if (own) j.img[1][(size_t)y1 * cols1 + x1] = g;
        }
        if (kDepth) {
            const float d = gauss_down_f32(rows, cols, x1, y1, [&](int cx, int cy) { return s_d0[cy - Y0][cx - X0]; });
            s_d1[sy][sx] = d;
            if (own) j.depth[1][(size_t)y1 * cols1 + x1] = d;
        }
    }
    __syncthreads();
    if (threadIdx.x < 64) {
        const int x2 = 8 * bx + (threadIdx.x & 7), y2 = 8 * by + (threadIdx.x >> 3);
        if (x2 < cols2 && y2 < rows2) {
            if (kImage) j.img[2][(size_t)y2 * cols2 + x2] = gauss_down_u8(rows1, cols1, x2, y2, [&](int cx, int cy) { return s_g1[cy - Y1][cx - X1]; });
            if (kDepth) j.depth[2][(size_t)y2 * cols2 + x2] = gauss_down_f32(rows1, cols1, x2, y2, [&](int cx, int cy) { return s_d1[cy - Y1][cx - X1]; });
        }
    }
}

// the intensity pyramid of a camera frame alone, straight from the uploaded RGB8: the input of the SO3 pre-alignment (level 2) and of
// the Sobel images / photometric residual; prep_all_kernel then only adds the depth pyramid (PrepAllArgs::cur_depth_only)
__global__ void __launch_bounds__(256) so3_image_kernel(const RgbdJob j, int rows, int cols)
{
    pdl_wait();
    rgbd_pyramid_tile<1>(j, rows, cols, 0.f, false, blockIdx.x, blockIdx.y);
}

__global__ void __launch_bounds__(256) prep_all_kernel(const PrepAllArgs A)
{
    pdl_wait();
    // grid = (tiles x, tiles y, 7 jobs): z = 0, 1 are the (heavier) RGB-D pyramid jobs -- lowest block indices, scheduled first --
    // z = 2..6 the map jobs; no index arithmetic beyond blockIdx
    const int job_z = A.jobs[blockIdx.z], bx = blockIdx.x, by = blockIdx.y;
    if (job_z < 2 ? (bx >= A.rgbd_bx || by >= A.rgbd_by) : (bx >= A.pyr_bx || by >= A.pyr_by)) return;
    // HRBFFusion::denseEnough (HRBFFusion.cpp:974-987): fill-in textures replace the prediction when <= thresh of the 1/20 samples are set
    bool alt;
    if (A.sel != nullptr) alt = *A.sel != 0;
    else if (A.dense_count != nullptr) {
        const int total = (A.cols / 20) * (A.rows / 20);
        alt = !((float)*A.dense_count / (float)total > A.dense_thresh);
        if (bx == 0 && by == 0 && job_z == 6 && threadIdx.x == 0 && A.dense_count_reset != nullptr) *A.dense_count_reset = 0u;
    } else alt = false;
    if (job_z >= 2) {
        switch (job_z - 2) {
        case 0: pyr_pair_tile<PYR_VN>(alt ? A.vm_alt : A.vm, alt ? A.nm_alt : A.nm, A.rows, A.cols, 0.f, A.pose, A.o_vg, A.o_ng, nullptr, 0.f, nullptr, nullptr, nullptr, bx, by); break;
        case 1: pyr_pair_tile<PYR_VN>(A.vc, A.nc, A.rows, A.cols, 0.f, nullptr, A.o_vc, A.o_nc, nullptr, 0.f, nullptr, nullptr, nullptr, bx, by); break;
        case 2: pyr_pair_tile<PYR_K>(alt ? A.k1m_alt : A.k1m, alt ? A.k2m_alt : A.k2m, A.rows, A.cols, A.curv_thr, A.pose, A.o_k1g, A.o_k2g, nullptr, 0.f, nullptr, nullptr, nullptr, bx, by); break;
        case 3: pyr_pair_tile<PYR_K>(A.k1c, A.k2c, A.rows, A.cols, A.curv_thr, nullptr, A.o_k1c, A.o_k2c, nullptr, 0.f, nullptr, nullptr, nullptr, bx, by); break;
        default: pyr_weight_tile(alt ? A.w_alt : A.w, A.rows, A.cols, A.o_w[0], A.w_pitch[0], A.o_w[1], A.w_pitch[1], A.o_w[2], A.w_pitch[2], nullptr, nullptr, bx, by); break;
        }
        return;
    }
    const int job = job_z;
    const bool use_alt = job == 0 && alt;
    if (job == 1 && A.cur_depth_only) rgbd_pyramid_tile<2>(A.rgbd[job], A.rows, A.cols, A.depth_cutoff, use_alt, bx, by);
    else rgbd_pyramid_tile<0>(A.rgbd[job], A.rows, A.cols, A.depth_cutoff, use_alt, bx, by);
}
__device__ __forceinline__ void sobel_pixel(int rows, int cols, const unsigned char* __restrict__ src, short* dx, short* dy, int x, int y)
{
    const float gx[9] = { 1, 0, -1, 2, 0, -2, 1, 0, -1 }, gy[9] = { 1, 2, 1, 0, 0, 0, -1, -2, -1 };
    float dxv = 0, dyv = 0;
    int k = 8;
    for (int j = max(y - 1, 0); j <= min(y + 1, rows - 1); ++j)
        for (int i = max(x - 1, 0); i <= min(x + 1, cols - 1); ++i) {
            const float s = (float)src[(size_t)j * cols + i];
            dxv += s * gx[k]; dyv += s * gy[k];
            --k;
        }
    dx[(size_t)y * cols + x] = (short)dxv;
    dy[(size_t)y * cols + x] = (short)dyv;
}
__global__ void sobel_kernel(int rows, int cols, const unsigned char* __restrict__ src, short* dx, short* dy)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    sobel_pixel(rows, cols, src, dx, dy, x, y);
}
__device__ __forceinline__ void project_pixel(int rows, int cols, const float* __restrict__ depth, float* cloud3, float invFx, float invFy, float cx, float cy, int x, int y)
{
    const size_t i = (size_t)y * cols + x;
    const float z = depth[i];
    cloud3[3 * i + 0] = (x - cx) * z * invFx;
    cloud3[3 * i + 1] = (y - cy) * z * invFy;
    cloud3[3 * i + 2] = z;
}
__global__ void project_cloud_kernel(int rows, int cols, const float* __restrict__ depth, float* cloud3,
                                     float invFx, float invFy, float cx, float cy)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    project_pixel(rows, cols, depth, cloud3, invFx, invFy, cx, cy, x, y);
}

// -------------------------------------------------------------- rows 1-3 ---
struct IcpArgs {
    const float *vc, *nc, *k1c, *k2c; int cpitch;      // current frame (camera frame)
    const float *vg, *ng, *k1g, *k2g; int gpitch;      // model prediction (global frame)
    const float* w; int wpitch;                        // icp weight map of the model
    int rows, cols;
    float fx, fy, cx, cy;
    float dist_thres, angle_thres;
    int use_search, radius, use_weight;
    int2* corres;                                      // optional output
    const float4 *pc0, *pc1, *pg0, *pg1;               // packed maps (PyrOut::pk0/pk1) of the current frame / the model, or null
};

// acc[0..27] += w*row_i*row_j (i<=j<7), acc[28] += inlier  (reduce.cu:511-545)
__device__ __forceinline__ void accumulate_row7(float (&acc)[32], const float (&row)[7], float weight, bool found)
{
    if (!found) return;      // every caller passes an all-zero row then: 28 FMAs that add exactly 0 (pixels are spatially coherent: little divergence)
    int k = 0;
#pragma unroll
    for (int i = 0; i < 7; ++i)
#pragma unroll
        for (int j = i; j < 7; ++j) acc[k++] += weight * row[i] * row[j];
    acc[28] += 1.f;
}

// Per-pixel projective association + point-to-plane row (reduce.cu:317-509).
template <bool SEARCH>
__device__ __forceinline__ void icp_pixel(const IcpArgs& a, const float* Rc, const float* tc, const float* Rpi, const float* tp,
                                          int i, float (&acc)[32])
{
    const int y = i / a.cols, x = i - y * a.cols;
    const int rows = a.rows;
    auto ldc = [&](const float* p, int plane) { return __ldg(p + (size_t)(plane * rows + y) * a.cpitch + x); };

    float row[7] = { 0, 0, 0, 0, 0, 0, 0 };
    float weight = 1.f;
    bool found = false;
    int bx = -1, by = -1, cnt = 0;
    float3 n_g, d_g, s_g;

    const float3 vcurr = make_float3(ldc(a.vc, 0), ldc(a.vc, 1), ldc(a.vc, 2));
    const float3 ncurr = make_float3(ldc(a.nc, 0), ldc(a.nc, 1), ldc(a.nc, 2));
    const float k1c = ldc(a.k1c, 3), k2c = ldc(a.k2c, 3);

    const float3 vg = mul(Rc, vcurr) + make_float3(tc[0], tc[1], tc[2]);
    const float3 vcp = mul(Rpi, vg - make_float3(tp[0], tp[1], tp[2]));
    const int ux = __float2int_rn(vcp.x * a.fx / vcp.z + a.cx);
    const int uy = __float2int_rn(vcp.y * a.fy / vcp.z + a.cy);

    if (!(ux < 0 || uy < 0 || ux >= a.cols || uy >= rows || vcp.z < 0) &&
        !(isnan(vcurr.x) || isnan(ncurr.x) || isnan(k1c) || isnan(k2c))) {
        const float3 ng = mul(Rc, ncurr);
        auto ldg = [&](const float* p, int plane, int cy, int cx) { return __ldg(p + (size_t)(plane * rows + cy) * a.gpitch + cx); };
        if (!SEARCH) {
            const float3 vp = make_float3(ldg(a.vg, 0, uy, ux), ldg(a.vg, 1, uy, ux), ldg(a.vg, 2, uy, ux));
            const float3 np = make_float3(ldg(a.ng, 0, uy, ux), ldg(a.ng, 1, uy, ux), ldg(a.ng, 2, uy, ux));
            const float k1 = ldg(a.k1g, 3, uy, ux), k2 = ldg(a.k2g, 3, uy, ux);
            const float dist = norm(vp - vg), sine = norm(cross(ng, np));
            if (!(isnan(vp.x) || isnan(np.x) || isnan(k1) || isnan(k2)) && !(sine > a.angle_thres || dist > a.dist_thres)) {
                found = true; bx = ux; by = uy; d_g = vp; n_g = np;
            }
        } else {
            const int R = a.radius, D = 2 * R + 1;
            float DpR = -1e8f, best_p = 1e8f;
            cnt = 0;
            for (int pass = 0; pass < 2; ++pass) {
                for (int cy = uy - D / 2; cy < uy + D / 2 + 1; ++cy)
                    for (int cx = ux - D / 2; cx < ux + D / 2 + 1; ++cx) {
                        if (cx < 0 || cy < 0 || cx >= a.cols || cy >= rows) continue;
                        const float3 vp = make_float3(ldg(a.vg, 0, cy, cx), ldg(a.vg, 1, cy, cx), ldg(a.vg, 2, cy, cx));
                        const float3 np = make_float3(ldg(a.ng, 0, cy, cx), ldg(a.ng, 1, cy, cx), ldg(a.ng, 2, cy, cx));
                        const float k1 = ldg(a.k1g, 3, cy, cx), k2 = ldg(a.k2g, 3, cy, cx);
                        const float dist = norm(vp - vg), sine = norm(cross(ng, np));
                        if (isnan(vp.x) || isnan(np.x) || isnan(k1) || isnan(k2)) continue;
                        if (sine > a.angle_thres || dist > a.dist_thres) continue;
                        if (pass == 0) { if (dist > DpR) DpR = dist; ++cnt; continue; }
                        const float ckmax = fmaxf(fabsf(k1), fabsf(k2));
                        const float D_p = dist / DpR, D_n = 1 - dot(np, ng);
                        const float D_c = 1 - expf(-fabsf(k1 - k1c) / ckmax) * expf(-fabsf(k2 - k2c) / ckmax);
                        const float p = 0.333f * D_p + 0.333f * D_n + 0.333f * D_c;
                        if (p < best_p) { bx = cx; by = cy; d_g = vp; n_g = np; best_p = p; }
                    }
                if (pass == 0 && cnt == 0) break;
            }
            // all scores NaN: undefined behaviour in the reference (uninitialised vectors, reduce.cu:404-434);
            // defined here, as in the oracle, as "no correspondence"
            found = bx >= 0;
        }
        s_g = vg;
    }
    if (a.corres) a.corres[i] = make_int2(bx, by);
    if (found) {
        const float3 tpv = make_float3(tp[0], tp[1], tp[2]);
        const float3 s_cp = mul(Rpi, s_g - tpv), d_cp = mul(Rpi, d_g - tpv), n_cp = mul(Rpi, n_g);
        if (a.use_weight) {
            const float w = __ldg(a.w + (size_t)by * a.wpitch + bx);
            weight = isnan(w) ? 0.f : w;
        }
        const float3 c = cross(s_cp, n_cp);
        row[0] = n_cp.x; row[1] = n_cp.y; row[2] = n_cp.z;
        row[3] = c.x; row[4] = c.y; row[5] = c.z;
        row[6] = dot(n_cp, s_cp - d_cp);
    }
    accumulate_row7(acc, row, weight, found);
}

// ---- the no-search ICP pixel split into load / gather / finish stages so that two pixels per thread are in flight
// (the pass is bound by two dependent L2 round trips per pixel, not by bandwidth or issue slots) ----
struct IcpCurr { float vx, vy, vz, nx, ny, nz, k1, k2; };
struct IcpModel { float vx, vy, vz, nx, ny, nz, k1, k2, w; int ok, ux, uy; float3 vg, ng; };

template <bool PACKED>
__device__ __forceinline__ IcpCurr icp_load_curr(const IcpArgs& a, int i)
{
    if (PACKED) {
        const float4 p0 = __ldg(a.pc0 + i), p1 = __ldg(a.pc1 + i);
        IcpCurr c;
        c.vx = p0.x; c.vy = p0.y; c.vz = p0.z; c.nx = p0.w; c.ny = p1.x; c.nz = p1.y; c.k1 = p1.z; c.k2 = p1.w;
        return c;
    }
    const int y = i / a.cols, x = i - y * a.cols, rows = a.rows;
    auto ld = [&](const float* p, int plane) { return __ldg(p + (size_t)(plane * rows + y) * a.cpitch + x); };
    IcpCurr c;
    c.vx = ld(a.vc, 0); c.vy = ld(a.vc, 1); c.vz = ld(a.vc, 2);
    c.nx = ld(a.nc, 0); c.ny = ld(a.nc, 1); c.nz = ld(a.nc, 2);
    c.k1 = ld(a.k1c, 3); c.k2 = ld(a.k2c, 3);
    return c;
}
template <bool PACKED>
__device__ __forceinline__ IcpModel icp_gather_model(const IcpArgs& a, const IcpCurr& c, const float* Rc, const float* tc, const float* Rpi, const float* tp)
{
    IcpModel m;
    const int rows = a.rows;
    m.vg = mul(Rc, make_float3(c.vx, c.vy, c.vz)) + make_float3(tc[0], tc[1], tc[2]);
    const float3 vcp = mul(Rpi, m.vg - make_float3(tp[0], tp[1], tp[2]));
    // IEEE division: an approximate one (what the reference's own build uses, --prec-div=false) moves ~50 of the 307 200
    // associations to the neighbouring model pixel and costs 2e-5 of pose parity on the GPUTest pair
    m.ux = __float2int_rn(vcp.x * a.fx / vcp.z + a.cx);
    m.uy = __float2int_rn(vcp.y * a.fy / vcp.z + a.cy);
    m.ok = !(m.ux < 0 || m.uy < 0 || m.ux >= a.cols || m.uy >= rows || vcp.z < 0) && !(isnan(c.vx) || isnan(c.nx) || isnan(c.k1) || isnan(c.k2));
    m.ng = mul(Rc, make_float3(c.nx, c.ny, c.nz));
    m.vx = m.vy = m.vz = m.nx = m.ny = m.nz = m.k1 = m.k2 = m.w = 0.f;
    if (m.ok && PACKED) {
        const int q = m.uy * a.cols + m.ux;
        const float4 p0 = __ldg(a.pg0 + q), p1 = __ldg(a.pg1 + q);
        m.vx = p0.x; m.vy = p0.y; m.vz = p0.z; m.nx = p0.w; m.ny = p1.x; m.nz = p1.y; m.k1 = p1.z; m.k2 = p1.w;
        m.w = a.use_weight ? __ldg(a.w + q) : 1.f;
    } else if (m.ok) {
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
        // ||.|| > thres  <=>  ||.||^2 > thres^2 (both sides non-negative): no square roots
        const float3 dv = vp - m.vg, cr = cross(m.ng, np);
        const float dist2 = dot(dv, dv), sine2 = dot(cr, cr);
        found = !(isnan(vp.x) || isnan(np.x) || isnan(m.k1) || isnan(m.k2)) && !(sine2 > a.angle_thres * a.angle_thres || dist2 > a.dist_thres * a.dist_thres);
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
// this CTA's contiguous pixel range [begin, end), kIcpInFlight pixels in flight per thread: all their current-frame loads
// are issued together, then all their model gathers, then the rows are accumulated (2 dependent memory round trips per
// trip of the loop; 640x480 level 0 = 2076 pixels per CTA = one full trip + a 28-pixel tail).
#ifndef HRBF_ICP_INFLIGHT
#define HRBF_ICP_INFLIGHT 2
#endif
constexpr int kIcpInFlight = HRBF_ICP_INFLIGHT;
template <int kThreads, bool PACKED>
__device__ __forceinline__ void icp_pass_nosearch_t(const IcpArgs& a, const float* Rc, const float* tc, const float* Rpi, const float* tp,
                                                    int begin, int end, float (&acc)[32])
{
    constexpr int kTrackThreads = kThreads;
    for (int i0 = begin + (int)threadIdx.x; i0 < end; i0 += kIcpInFlight * kTrackThreads) {
        IcpCurr c[kIcpInFlight];
#pragma unroll
        for (int u = 0; u < kIcpInFlight; ++u) {
            const int i = i0 + u * kTrackThreads;
            c[u] = icp_load_curr<PACKED>(a, i < end ? i : i0);
        }
        IcpModel m[kIcpInFlight];
#pragma unroll
        for (int u = 0; u < kIcpInFlight; ++u) m[u] = icp_gather_model<PACKED>(a, c[u], Rc, tc, Rpi, tp);
#pragma unroll
        for (int u = 0; u < kIcpInFlight; ++u) {
            const int i = i0 + u * kTrackThreads;
            if (i < end) icp_finish(a, m[u], Rpi, tp, i, acc);
        }
    }
}

template <int kThreads>
__device__ __forceinline__ void icp_pass_nosearch(const IcpArgs& a, const float* Rc, const float* tc, const float* Rpi, const float* tp,
                                                  int begin, int end, float (&acc)[32])
{
    if (a.pc0 != nullptr) icp_pass_nosearch_t<kThreads, true>(a, Rc, tc, Rpi, tp, begin, end, acc);      // uniform branch
    else icp_pass_nosearch_t<kThreads, false>(a, Rc, tc, Rpi, tp, begin, end, acc);
}

// SoA planes -> packed records, for the map builders that only write SoA (GPUTest path: createVMap / createNMap, neutral curvature)
__global__ void pack_maps_kernel(int n, const float* __restrict__ v, const float* __restrict__ nm, const float* __restrict__ k1, const float* __restrict__ k2,
                                 float4* __restrict__ pk0, float4* __restrict__ pk1)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    pk0[i] = make_float4(v[i], v[n + i], v[2 * n + i], nm[i]);
    pk1[i] = make_float4(nm[n + i], nm[2 * n + i], k1[3 * n + i], k2[3 * n + i]);
}

// mode 0: store the 29 sums in st->icp_sums only (hrbf_icp_step, or RGB still to come)
// mode 1: store and run the Gauss-Newton update in the last block
template <bool SEARCH>
__global__ void __launch_bounds__(kReduceThreads, 2) icp_reduce_kernel(IcpArgs a, ReduceWork* wk, int mode, int cur_level, int next_level)
{
    TrackState* st = &wk->st;
    __shared__ double s_total[32];
    __shared__ float s_pose[24];
    pdl_wait();      // (no-op unless launched with programmatic stream serialization)
    if (threadIdx.x < 9) { s_pose[threadIdx.x] = st->Rcurr[threadIdx.x]; s_pose[12 + threadIdx.x] = st->Rprev_inv[threadIdx.x]; }
    if (threadIdx.x < 3) { s_pose[9 + threadIdx.x] = st->tcurr[threadIdx.x]; s_pose[21 + threadIdx.x] = st->tprev[threadIdx.x]; }
    __syncthreads();
    float Rc[9], tc[3], Rpi[9], tp[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) { Rc[k] = s_pose[k]; Rpi[k] = s_pose[12 + k]; }
#pragma unroll
    for (int k = 0; k < 3; ++k) { tc[k] = s_pose[9 + k]; tp[k] = s_pose[21 + k]; }

    float acc[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) acc[k] = 0.f;
    const int N = a.rows * a.cols;
    const bool level_done = (st->done_level == cur_level);
    if (!level_done) {
        if (SEARCH) for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += blockDim.x * gridDim.x) icp_pixel<true>(a, Rc, tc, Rpi, tp, i, acc);
        else {
            // contiguous pixel range per CTA, several pixels per thread in flight (same pass as the persistent tracker)
            const int begin = (int)(((long long)N * blockIdx.x) / gridDim.x), end = (int)(((long long)N * (blockIdx.x + 1)) / gridDim.x);
            icp_pass_nosearch<kReduceThreads>(a, Rc, tc, Rpi, tp, begin, end, acc);
        }
    }

    if (grid_reduce32(acc, wk->partials, &st->ticket, s_total)) {
        if (threadIdx.x < 32) st->icp_sums[threadIdx.x] = s_total[threadIdx.x];
        __syncthreads();
        if (threadIdx.x == 0 && mode == 1 && !level_done) gn_update(st, cur_level, next_level);
    }
}

struct RgbResArgs {
    float minScale, maxDepthDelta;
    const short *dIdx, *dIdy;
    const float *lastDepth, *nextDepth;
    const unsigned char *lastImage, *nextImage;
    hrbf_dataterm* corres;
    int rows, cols;
};

// The pose-independent tests of computeRgbResidual (reduce.cu:1000-1023) for pixel k: inside the processed region, no
// zero pixel of the next image in the 4x4 window (:1005-1011), gradient magnitude^2 >= minScale, next depth not NaN.
// All loads are issued up front (clamped addresses, out-of-window taps ignored) instead of a short-circuit chain.
__device__ __forceinline__ bool rgb_static_candidate(const RgbResArgs& a, int k, short valx, short valy)
{
    const int cols = a.cols, rows = a.rows;
    const int i = k / cols, j0 = k - i * cols;
    if (!(j0 < cols - 5 && i < rows - 1)) return false;
    unsigned int zero_seen = 0;
#pragma unroll
    for (int du = -2; du < 2; ++du)
#pragma unroll
        for (int dv = -2; dv < 2; ++dv) {
            const int u = i + du, v = j0 + dv;
            const bool in = u >= 0 && u < rows && v >= 0 && v < cols;
            const unsigned char px = __ldg(a.nextImage + (size_t)min(max(u, 0), rows - 1) * cols + min(max(v, 0), cols - 1));
            zero_seen |= (in && px == 0) ? 1u : 0u;
        }
    if (zero_seen != 0) return false;
    const float mTwo = (float)((valx * valx) + (valy * valy));
    if (!(mTwo >= a.minScale)) return false;
    return !isnan(__ldg(a.nextDepth + k));
}

// row r of the photometric warp d1 * (K.x * x + K.y * y + K.z) + kt (reduce.cu:1029-1032), with the FMA contraction nvcc applies to
// that expression spelled out -- FMUL y K.y, FFMA x K.x + ., FADD + K.z, FFMA d1 . + kt -- so that the rounding of u0 / v0 does
// not depend on the compiler's mood; s_k = krkinv[9], kt[3]
__device__ __forceinline__ float rgb_warp_row(const float* s_k, int r, int x, int y, float d1)
{
    return fmaf(d1, __fadd_rn(fmaf((float)x, s_k[3 * r], __fmul_rn((float)y, s_k[3 * r + 1])), s_k[3 * r + 2]), s_k[9 + r]);
}

// one pixel of computeRgbResidual (reduce.cu:986-1060); s_k = krkinv[9], kt[3]
__device__ __forceinline__ void rgb_residual_pixel(const RgbResArgs& a, const float* s_k, int k, int& cnt, int& sig)
{
    const int cols = a.cols, rows = a.rows;
    const int i = k / cols, j0 = k - i * cols;
    hrbf_dataterm c;
    c.zero_x = c.zero_y = c.one_x = c.one_y = 0; c.diff = 0.f; c.valid = 0; c.pad[0] = c.pad[1] = c.pad[2] = 0;
    if (rgb_static_candidate(a, k, __ldg(a.dIdx + k), __ldg(a.dIdy + k))) {
        const int y = i, x = j0;
        const float d1 = __ldg(a.nextDepth + k);
        const float td1 = rgb_warp_row(s_k, 2, x, y, d1);
        const int u0 = __float2int_rn(rgb_warp_row(s_k, 0, x, y, d1) / td1);
        const int v0 = __float2int_rn(rgb_warp_row(s_k, 1, x, y, d1) / td1);
        if (u0 >= 0 && v0 >= 0 && u0 < cols && v0 < rows) {
            const float d0 = __ldg(a.lastDepth + (size_t)v0 * cols + u0);
            const unsigned char li = __ldg(a.lastImage + (size_t)v0 * cols + u0);
            if (d0 > 0 && fabsf(td1 - d0) <= a.maxDepthDelta && li != 0) {
                c.zero_x = (short)u0; c.zero_y = (short)v0; c.one_x = (short)x; c.one_y = (short)y;
                c.diff = (float)__ldg(a.nextImage + k) - (float)li;
                c.valid = 1;
                cnt += 1;
                sig += (int)(c.diff * c.diff);
            }
        }
    }
    a.corres[k] = c;
}

// reduce.cu:986-1060; the int2 {count, sum diff^2} goes through redux + one atomic per warp.
// When the caller is the tracking loop, the last block also evaluates sigma / the rgbOnly break
// (RGBDOdometry.cpp:1017-1032).
__global__ void __launch_bounds__(256) rgb_residual_kernel(RgbResArgs a, ReduceWork* wk, int finalize, int cur_level, int first_iter, int next_lower)
{
    TrackState* st = &wk->st;
    __shared__ float s_k[12];
    __shared__ unsigned int s_last;
    if (threadIdx.x < 9) s_k[threadIdx.x] = st->krkinv[threadIdx.x];
    if (threadIdx.x < 3) s_k[9 + threadIdx.x] = st->kt[threadIdx.x];
    __syncthreads();
    const int N = a.rows * a.cols;
    int cnt = 0, sig = 0;
    const bool level_done = (st->done_level == cur_level);
    if (!level_done)
        for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < N; k += blockDim.x * gridDim.x) rgb_residual_pixel(a, s_k, k, cnt, sig);
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    sig = __reduce_add_sync(0xffffffffu, sig);
    if ((threadIdx.x & 31) == 0 && (cnt | sig)) { atomicAdd(&st->rgb_count, cnt); atomicAdd(&st->rgb_sigma, sig); }
    if (!finalize) return;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(&st->ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        st->ticket = 0u;
        if (!level_done) {
            const int sigma = *(volatile int*)&st->rgb_sigma, rgbSize = *(volatile int*)&st->rgb_count;
            st->rgb_sigma = 0; st->rgb_count = 0;   // consumed: re-arm for the next iteration
            float sigmaVal = sqrtf(((float)sigma / (float)rgbSize == 0.f) ? 1.f : (float)rgbSize);
            const float rgbError = sqrtf((float)sigma) / (float)(rgbSize == 0 ? 1 : rgbSize);
            const float lastErr = first_iter ? FLT_MAX : st->lastRGBError;   // RGBDOdometry.cpp:962
            if (st->rgbOnly && rgbError > lastErr) {
                st->done_level = cur_level;
                // the next level's first iteration warps with ITS camera matrix (RGBDOdometry.cpp:983-992 rebuilds K R K^-1, K t
                // at the top of every iteration): the last gn_update prepared them for cur_level
                if (next_lower >= 0) {
                    double Rt[16];
                    for (int k = 0; k < 16; ++k) Rt[k] = st->resultRt[k];
                    update_krk(st, Rt, next_lower);
                }
            } else {
                st->lastRGBError = rgbError;
                st->lastRGBCount = (float)rgbSize;
                if (st->rgbOnly) sigmaVal = -1.f;
                st->sigmaVal = sigmaVal;
            }
        }
    }
}

struct RgbStepArgs {
    const hrbf_dataterm* corres;
    const float* cloud3;
    const short *dIdx, *dIdy;
    float fx, fy, sobelScale;
    int use_grad_weight;
    int rows, cols;
};

// one pixel of rgbStep (reduce.cu:718-811)
__device__ __forceinline__ void rgb_step_pixel(const RgbStepArgs& a, float sigma, int i, float (&acc)[32])
{
    const int cols = a.cols;
    const int4 raw = __ldg(reinterpret_cast<const int4*>(a.corres) + i);
    hrbf_dataterm c;
    memcpy(&c, &raw, sizeof c);
    float row[7] = { 0, 0, 0, 0, 0, 0, 0 };
    float rgb_weight = 1.f;
    if (c.valid) {
        float w = sigma + fabsf(c.diff);
        w = w > 1.19209290E-07F ? 1.0f / w : 1.0f;
        if (sigma == -1.f) w = 1.f;
        row[6] = -w * c.diff;
        const float* cp = a.cloud3 + 3 * ((size_t)c.zero_y * cols + c.zero_x);
        const float px = __ldg(cp), py = __ldg(cp + 1), pz = __ldg(cp + 2);
        const float invz = 1.0f / pz;
        const size_t o1 = (size_t)c.one_y * cols + c.one_x;
        const float gx = w * a.sobelScale * (float)__ldg(a.dIdx + o1), gy = w * a.sobelScale * (float)__ldg(a.dIdy + o1);
        const float v0 = gx * a.fx * invz, v1 = gy * a.fy * invz;
        const float v2 = -(v0 * px + v1 * py) * invz;
        row[0] = v0; row[1] = v1; row[2] = v2;
        row[3] = -pz * v1 + py * v2;
        row[4] = pz * v0 - px * v2;
        row[5] = -py * v0 + px * v1;
        if (a.use_grad_weight) {
            const float gm = sqrtf(gx * gx + gy * gy);
            rgb_weight = expf(-0.5f * (10.f / gm) * (10.f / gm));
        }
    }
    accumulate_row7(acc, row, rgb_weight, c.valid != 0);

}

// reduce.cu:718-811.  sigma < -1.5 => take st->sigmaVal (tracking loop).
__global__ void __launch_bounds__(kReduceThreads, 2) rgb_step_kernel(RgbStepArgs a, float sigma_arg, ReduceWork* wk, int mode, int cur_level, int next_level)
{
    TrackState* st = &wk->st;
    __shared__ double s_total[32];
    const float sigma = sigma_arg < -1.5f ? st->sigmaVal : sigma_arg;
    float acc[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) acc[k] = 0.f;
    const int N = a.rows * a.cols;
    const bool level_done = (st->done_level == cur_level);
    if (!level_done)
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += blockDim.x * gridDim.x) rgb_step_pixel(a, sigma, i, acc);
    if (grid_reduce32(acc, wk->partials, &st->ticket, s_total)) {
        if (threadIdx.x < 32) st->rgb_sums[threadIdx.x] = s_total[threadIdx.x];
        __syncthreads();
        if (threadIdx.x == 0 && mode == 1 && !level_done) gn_update(st, cur_level, next_level);
    }
}

// one pixel of so3Step (reduce.cu:1172-1273); s_m = imageBasis[9], kinv[9], krlr[9]
__device__ __forceinline__ void so3_pixel(const unsigned char* __restrict__ lastImage, const unsigned char* __restrict__ nextImage,
                                          int rows, int cols, const float* s_m, int k, float (&acc)[32])
{
    const int y = k / cols, x = k - y * cols;
    const float3 up = make_float3((float)x, (float)y, 1.0f);
    const float3 wp = mul(s_m, up);
    const int wx = __float2int_rn(wp.x / wp.z), wy = __float2int_rn(wp.y / wp.z);
    const bool found = wx >= 1 && wx < cols - 1 && wy >= 1 && wy < rows - 1 && x >= 1 && x < cols - 1 && y >= 1 && y < rows - 1;
    float row[4] = { 0, 0, 0, 0 };
    if (found) {
        auto grad = [&](const unsigned char* img, int px, int py, float& gx, float& gy) {
            const float actu = (float)__ldg(img + (size_t)py * cols + px);
            float back = (float)__ldg(img + (size_t)py * cols + px - 1), fore = (float)__ldg(img + (size_t)py * cols + px + 1);
            gx = ((back + actu) / 2.0f) - ((fore + actu) / 2.0f);
            back = (float)__ldg(img + (size_t)(py - 1) * cols + px); fore = (float)__ldg(img + (size_t)(py + 1) * cols + px);
            gy = ((back + actu) / 2.0f) - ((fore + actu) / 2.0f);
        };
        float gnx, gny, glx, gly;
        grad(nextImage, wx, wy, gnx, gny);
        grad(lastImage, x, y, glx, gly);
        const float gx = (gnx + glx) / 2.0f, gy = (gny + gly) / 2.0f;
        const float3 pt = mul(s_m + 9, up);
        const float z2 = pt.z * pt.z;
        const float* kr = s_m + 18;
        const float3 lp = make_float3(((pt.z * (kr[3] * gy + kr[0] * gx)) - (gy * kr[6] * y) - (gx * kr[6] * x)) / z2,
                                      ((pt.z * (kr[4] * gy + kr[1] * gx)) - (gy * kr[7] * y) - (gx * kr[7] * x)) / z2,
                                      ((pt.z * (kr[5] * gy + kr[2] * gx)) - (gy * kr[8] * y) - (gx * kr[8] * x)) / z2);
        const float3 jr = cross(lp, pt);
        row[0] = jr.x; row[1] = jr.y; row[2] = jr.z;
        row[3] = -((float)__ldg(nextImage + (size_t)wy * cols + wx) - (float)__ldg(lastImage + (size_t)y * cols + x));
    }
    int q = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = i; j < 4; ++j) acc[q++] += row[i] * row[j];
    acc[10] += found ? 1.f : 0.f;

}

// reduce.cu:1172-1273.  mode 1: run the SO3 control flow in the last block.
__global__ void __launch_bounds__(kReduceThreads, 2) so3_reduce_kernel(const unsigned char* __restrict__ lastImage, const unsigned char* __restrict__ nextImage,
                                                                      int rows, int cols, ReduceWork* wk, int mode)
{
    TrackState* st = &wk->st;
    __shared__ double s_total[32];
    __shared__ float s_m[27];
    if (threadIdx.x < 9) { s_m[threadIdx.x] = st->so3_basis[threadIdx.x]; s_m[9 + threadIdx.x] = st->so3_kinv[threadIdx.x]; s_m[18 + threadIdx.x] = st->so3_krlr[threadIdx.x]; }
    __syncthreads();
    float acc[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) acc[k] = 0.f;
    const int N = rows * cols;
    const bool skip = (mode == 1) && st->so3_done;
    if (!skip)
        for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < N; k += blockDim.x * gridDim.x) so3_pixel(lastImage, nextImage, rows, cols, s_m, k, acc);
    if (grid_reduce32(acc, wk->partials, &st->ticket, s_total)) {
        if (!skip) {
            if (threadIdx.x < 16) st->so3_sums[threadIdx.x] = s_total[threadIdx.x];
            __syncthreads();
            if (threadIdx.x == 0 && mode == 1) so3_update(st);
        }
    }
}

// Start of a tracking call: load the previous pose, reset the state (RGBDOdometry.cpp:806-846, 916-935)
__global__ void track_begin_kernel(ReduceWork* wk, const float* __restrict__ prev_pose /* R[9], t[3] */,
                                   int icp, int rgb, int rgbOnly, int so3, float icpWeight, int first_level)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    TrackState* st = &wk->st;
    for (int k = 0; k < 9; ++k) { st->Rprev[k] = prev_pose[k]; st->Rcurr[k] = prev_pose[k]; }
    for (int k = 0; k < 3; ++k) { st->tprev[k] = prev_pose[9 + k]; st->tcurr[k] = prev_pose[9 + k]; }
    inv3f(st->Rprev, st->Rprev_inv);
    for (int k = 0; k < 16; ++k) st->resultRt[k] = (k % 5 == 0) ? 1.0 : 0.0;
    for (int k = 0; k < 9; ++k) { st->resultR[k] = st->lastResultR[k] = (k % 4 == 0) ? 1.0 : 0.0; st->R_lr[k] = (k % 4 == 0) ? 1.f : 0.f; }
    st->so3_lastError = FLT_MAX / 2; st->so3_lastCount = FLT_MAX / 2; st->so3_done = 0;
    st->icp = icp; st->rgb = rgb; st->rgbOnly = rgbOnly; st->so3 = so3; st->icpWeight = icpWeight;
    st->done_level = -1; st->rgb_count = 0; st->rgb_sigma = 0; st->sigmaVal = 0.f;
    st->lastICPError = 0; st->lastICPCount = 0; st->lastRGBError = FLT_MAX; st->lastRGBCount = 0; st->lastSO3Error = 0; st->lastSO3Count = 0;
    st->icp_iterations_run = 0; st->ticket = 0u;
    for (int k = 0; k < 32; ++k) { st->icp_sums[k] = 0; st->rgb_sums[k] = 0; }
    const double I3[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
    const double I4[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
    if (so3) update_so3_mats(st, I3);
    else if (rgb) update_krk(st, I4, first_level);
}
// After the SO3 iterations: seed resultRt with the rotation (RGBDOdometry.cpp:926-935)
__global__ void track_after_so3_kernel(ReduceWork* wk, int first_level)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    TrackState* st = &wk->st;
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) st->resultRt[a * 4 + b] = st->resultR[a * 3 + b];
    if (st->rgb) {
        double Rt[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) Rt[k] = st->resultRt[k];
        update_krk(st, Rt, first_level);
    }
}
// End of a tracking call: 0.3 m guard (RGBDOdometry.cpp:1232-1236), pose out
__global__ void track_end_kernel(ReduceWork* wk, float* __restrict__ pose_out)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    TrackState* st = &wk->st;
    if (st->rgb) {
        const float dx = st->tcurr[0] - st->tprev[0], dy = st->tcurr[1] - st->tprev[1], dz = st->tcurr[2] - st->tprev[2];
        if ((double)sqrtf(dx * dx + dy * dy + dz * dz) > 0.3) {
            for (int k = 0; k < 9; ++k) st->Rcurr[k] = st->Rprev[k];
            for (int k = 0; k < 3; ++k) st->tcurr[k] = st->tprev[k];
        }
    }
    for (int k = 0; k < 9; ++k) pose_out[k] = st->Rcurr[k];
    for (int k = 0; k < 3; ++k) pose_out[9 + k] = st->tcurr[k];
}

}  // namespace hrbf
