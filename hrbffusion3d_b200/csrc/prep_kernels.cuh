// prep_kernels.cuh -- sm_100a kernels for SURVEY.md section 8 row 10: the per-frame preprocessing that
// produces the ICP "curr" maps (replaces the full-screen GLSL passes driven by
// Core/src/HRBFFusion.cpp:1016-1021,1262-1346), plus FillIn (Shaders/FillIn.cpp) and the
// dense-enough test (HRBFFusion.cpp:974-987, Shaders/Resize.cpp).
//   depth_bilateral.frag + depth_metric_{raw,filtered}.frag       -> depth_filter_metric_kernel (one pass, smem tile)
//   depth_vertex_normal_radius.frag (PCA normals, geometry.glsl)  -> vertex_normal_radius_kernel
//   depth_curvature_gradient.frag (HRBF gradient + 3rd-derivative curvature) -> curvature_gradient_kernel
//   depth_confidence_evaluation.frag                              -> confidence_kernel
//   fill_{vertex,normal,curvature,rgb}.frag                       -> fill_in_kernel (4 passes fused)
#pragma once
#include "common.cuh"
#include <type_traits>

namespace hrbf {

// The 7x7 windows of the PCA normal (geometry.glsl:198-212) and of the curvature pass (depth_curvature_gradient.frag:62-75) as
// the shaders' float-counter loops really visit them: `for (float i = tx_min; i <= tx_max; i += step)` overshoots tx_max by an
// ulp for ~40 % of the columns / rows and then never visits the window's last column / row, and the sample coordinate i * cols
// is px + 0.5 only up to that round-off.  Host-built tables (hrbf_window_table) give, per pixel of an axis, the first texel,
// the sample count and the sample coordinates the shader's fp32 expressions produce.
struct WinTab { const int* first; const int* count; const float* coord; };      // coord[p * kWinMax + k]
constexpr int kWinMax = 8;
struct PrepArgs {
    int cols, rows;
    float cx, cy, icx, icy;          // cam = (cx, cy, 1/fx, 1/fy)
    float depthFactor, maxD;         // metres per raw unit, globalDepthCutoff
    float radiusMultiplier;
    int pca, curvWin, bilateral;
    WinTab wx[2], wy[2];             // [0]: texcoords of a full-screen fragment pass, [1]: of the uv VBO (GlobalModel::fuse)
};

// exp() of the bilateral weights as a fixed sequence of IEEE fp32 operations (explicit FMAs in the Horner scheme), identical to the oracle's
// orc_exp_bilateral: the filtered depth must be bit-reproducible because the PCA normal estimation downstream
// amplifies 1-ulp depth differences to ~1e-3 in the normal.
__device__ __forceinline__ float exp_bilateral(float x)
{
    if (!(x > -87.0f)) return 0.0f;
    const float t = __fmul_rn(x, 1.44269504088896341f);
    const float n = rintf(t);
    const float f = __fsub_rn(t, n);
    float p = 1.54035304e-4f;
    p = fmaf(p, f, 1.33335581e-3f);
    p = fmaf(p, f, 9.61812911e-3f);
    p = fmaf(p, f, 5.55041087e-2f);
    p = fmaf(p, f, 2.40226507e-1f);
    p = fmaf(p, f, 6.93147181e-1f);
    p = fmaf(p, f, 1.0f);
    return ldexpf(p, (int)n);
}

// exp_bilateral for the filter loop: same operation sequence, with ldexpf(p, n) done as an exponent add.  n <= 0 here and
// p is in [0.70, 1.42], so p * 2^n is a normal number for n >= -125; for n = -126 (weights < 1.7e-38, which the oracle
// produces as denormals) 0 is returned: such a weight cannot change sums that contain the centre tap's weight 1.
__device__ __forceinline__ float exp_bilateral_fast(float x)
{
    const float t = __fmul_rn(x, 1.44269504088896341f);
    const float tb = __fadd_rn(fmaxf(t, -4194303.0f), 12582912.0f);      // rintf via the 1.5 * 2^23 trick (see exp_bilateral_pair); the clamp keeps it exact for any x
    const float n = __fadd_rn(tb, -12582912.0f);
    const float f = __fsub_rn(t, n);
    float p = 1.54035304e-4f;
    p = fmaf(p, f, 1.33335581e-3f);
    p = fmaf(p, f, 9.61812911e-3f);
    p = fmaf(p, f, 5.55041087e-2f);
    p = fmaf(p, f, 2.40226507e-1f);
    p = fmaf(p, f, 6.93147181e-1f);
    p = fmaf(p, f, 1.0f);
    const int e = __float_as_int(tb) - 0x4B400000;
    return (x > -87.0f && e >= -125) ? __int_as_float(__float_as_int(p) + (e << 23)) : 0.0f;
}

// Two taps at once on packed fp32 pairs (fma.rn.f32x2 / mul.rn.f32x2: FFMA2 / FMUL2 in SASS): the weights exp_bilateral_fast gives for the taps
// t.x and t.y that share the spatial term sp (dx and -dx of a window row).  Every packed operation rounds each half exactly like the
// scalar sequence (value - tmp as fma(tmp, -1, value); (-a) * log2e as a * (-log2e); t - n as fma(n, -1, t)), so the weights are
// bit-identical; the kernel is issue-bound and this halves the issue slots of the arithmetic.  x > -87 is implied by e >= -125
// (x <= -87 gives t <= -125.51 and n = -126).
__device__ __forceinline__ float2 exp_bilateral_pair(float2 value2, float2 t, float sc, float sp)
{
    const float2 neg1 = make_float2(-1.0f, -1.0f);
    const float2 dc = __ffma2_rn(t, neg1, value2);
    const float2 arg = __ffma2_rn(mul2_rn(dc, dc), make_float2(sc, sc), make_float2(sp, sp));
    const float2 tt = mul2_rn(arg, make_float2(-1.44269504088896341f, -1.44269504088896341f));      // (must not fuse with the sum below)
    // rintf(t) as (t + 1.5 * 2^23) - 1.5 * 2^23: exact for |t| < 2^22 (round to nearest even in both), and the integer sits in the low mantissa
    // bits of the first sum -- no FRND / F2I (quarter-rate XU pipe: it was the busiest pipe of the kernel at 68 %)
    const float2 big = make_float2(12582912.0f, 12582912.0f);
    const float2 tb = add2_rn(tt, big);
    const float2 n = add2_rn(tb, make_float2(-12582912.0f, -12582912.0f));
    const float2 f = __ffma2_rn(n, neg1, tt);
    float2 p = make_float2(1.54035304e-4f, 1.54035304e-4f);
    p = __ffma2_rn(p, f, make_float2(1.33335581e-3f, 1.33335581e-3f));
    p = __ffma2_rn(p, f, make_float2(9.61812911e-3f, 9.61812911e-3f));
    p = __ffma2_rn(p, f, make_float2(5.55041087e-2f, 5.55041087e-2f));
    p = __ffma2_rn(p, f, make_float2(2.40226507e-1f, 2.40226507e-1f));
    p = __ffma2_rn(p, f, make_float2(6.93147181e-1f, 6.93147181e-1f));
    p = __ffma2_rn(p, f, make_float2(1.0f, 1.0f));
    const int ex = __float_as_int(tb.x) - 0x4B400000, ey = __float_as_int(tb.y) - 0x4B400000;
    return make_float2(ex >= -125 ? __int_as_float(__float_as_int(p.x) + (ex << 23)) : 0.0f,
                       ey >= -125 ? __int_as_float(__float_as_int(p.y) + (ey << 23)) : 0.0f);
}

// ---- depth_bilateral.frag + depth_metric_raw.frag + depth_metric_filtered.frag -------------------------
constexpr int kBilR = 6, kBilTW = 32, kBilTH = 8;
// the pose-independent spatial term (dx^2 + dy^2) * 0.024691358f of each of the 13 x 13 taps, rounded like the shader does
__constant__ float c_bil_space[(2 * kBilR + 1) * (2 * kBilR + 1)];
inline void make_bilateral_table(float* t)
{
    for (int dy = -kBilR; dy <= kBilR; ++dy)
        for (int dx = -kBilR; dx <= kBilR; ++dx) {
            const float fx = (float)dx, fy = (float)dy;
            volatile float a = fx * fx, b = fy * fy;       // volatile: keep the two products and the sum separately rounded
            volatile float s2 = a + b;
            volatile float r = s2 * 0.024691358f;
            t[(dy + kBilR) * (2 * kBilR + 1) + dx + kBilR] = r;
        }
}
__global__ void __launch_bounds__(256) depth_filter_metric_kernel(PrepArgs a, const unsigned short* __restrict__ raw,
                                                                  float* __restrict__ filtered, float* __restrict__ metric, float* __restrict__ metric_filtered)
{
    pdl_wait();
    __shared__ float s_t[kBilTH + 2 * kBilR][kBilTW + 2 * kBilR];
    const int W = a.cols, H = a.rows;
    const float adj = 1.0f / (a.depthFactor * 1000.0f);
    const int x0 = blockIdx.x * kBilTW - kBilR, y0 = blockIdx.y * kBilTH - kBilR;
    for (int t = threadIdx.x; t < (kBilTH + 2 * kBilR) * (kBilTW + 2 * kBilR); t += 256) {
        const int sy = t / (kBilTW + 2 * kBilR), sx = t - sy * (kBilTW + 2 * kBilR);
        const int gx = x0 + sx, gy = y0 + sy;
        s_t[sy][sx] = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? __fdiv_rn((float)__ldg(raw + (size_t)gy * W + gx), adj) : 0.f;
    }
    __syncthreads();
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    const int x = blockIdx.x * kBilTW + lx, y = blockIdx.y * kBilTH + ly;
    if (x >= W || y >= H) return;
    const size_t o = (size_t)y * W + x;
    const unsigned int rv = __ldg(raw + o);
    const float value = s_t[ly + kBilR][lx + kBilR];
    float out = 0.f;
    if (!(value > a.maxD * 1000.0f || value < 300.0f)) {
        if (a.bilateral) {
            // the shader's loops run over the window clamped to the image (row-major); here every lane walks the full
            // 13 x 13 window in the same order and a tap outside the image gets weight 0 (adds exactly 0 to both sums)
            const float sc = 0.000555556f;
            float sum1 = 0.f, sum2 = 0.f;
            // tiles whose 13 x 13 windows lie inside the image (all but the border tiles) skip the per-tap bounds test
            const bool interior = x0 >= 0 && y0 >= 0 && x0 + kBilTW + 2 * kBilR <= W && y0 + kBilTH + 2 * kBilR <= H;
            const float2 value2 = make_float2(value, value);
            auto window = [&](auto checked) {
                for (int dy = -kBilR; dy <= kBilR; ++dy) {
                    const bool iny = (unsigned)(y + dy) < (unsigned)H;
                    const float* row = &s_t[ly + kBilR + dy][lx];
                    const float* sp = c_bil_space + (dy + kBilR) * (2 * kBilR + 1);
                    constexpr int kTaps = 2 * kBilR + 1;
                    float tmp[kTaps], w[kTaps];
#pragma unroll
                    for (int k = 0; k < kTaps; ++k) tmp[k] = row[k];
                    // the weights of a row, two taps (dx, -dx: the same spatial term) per packed operation; the centre tap alone
#pragma unroll
                    for (int k = 0; k < kBilR; ++k) {
                        const float2 w2 = exp_bilateral_pair(value2, make_float2(tmp[k], tmp[kTaps - 1 - k]), sc, sp[k]);
                        w[k] = w2.x; w[kTaps - 1 - k] = w2.y;
                    }
                    {
                        const float dc = __fsub_rn(value, tmp[kBilR]);
                        w[kBilR] = exp_bilateral_fast(-fmaf(__fmul_rn(dc, dc), sc, sp[kBilR]));
                    }
                    // the sums in the shader's tap order
#pragma unroll
                    for (int k = 0; k < kTaps; ++k) {
                        float weight = w[k];
                        if (decltype(checked)::value && !(iny && (unsigned)(x + k - kBilR) < (unsigned)W)) weight = 0.f;
                        sum1 = fmaf(tmp[k], weight, sum1);
                        sum2 = __fadd_rn(sum2, weight);
                    }
                }
            };
            if (interior) window(std::false_type{}); else window(std::true_type{});
            out = __fmul_rn(__fdiv_rn(sum1, sum2), adj);
        } else out = (float)rv;
    }
    if (filtered) filtered[o] = out;
    const unsigned int hi = (unsigned int)(a.maxD / a.depthFactor), lo = (unsigned int)(0.3f / a.depthFactor);
    metric[o] = (rv > hi || rv < lo) ? 0.f : (float)rv * a.depthFactor;
    metric_filtered[o] = (out > a.maxD / a.depthFactor || out < 0.3f / a.depthFactor) ? 0.f : __fmul_rn(out, a.depthFactor);
}

// ---- surfels.glsl / geometry.glsl helpers -------------------------------------------------------------
__device__ __forceinline__ float get_radius(float icx, float icy, float depth, float norm_z)
{
    const float meanFocal = ((1.0f / fabsf(icx)) + (1.0f / fabsf(icy))) / 2.0f;
    const float radius = (depth / meanFocal) * 1.41421356237f;
    const float radius_n = radius / fabsf(norm_z);
    const float two = 2.0f * radius;
    return two < radius_n ? two : radius_n;
}
__device__ __forceinline__ float confidence_fn(float cx, float cy, float x, float y, float max_dist, float weighting)
{
    const float dx = x - cx, dy = y - cy;
    const float radialDist = sqrtf(dx * dx + dy * dy) / max_dist;
    return expf((-(radialDist * radialDist) / 0.72f)) * weighting;
}
__device__ __forceinline__ void roots2(float b, float c, float (&r)[3])
{
    float d = b * b - 4.0f * c;
    if (d < 0.0f) d = 0.0f;
    const float sd = sqrtf(d);
    r[0] = 0.0f; r[1] = 0.5f * (b + sd); r[2] = 0.5f * (b - sd);
}
// geometry.glsl:86-160 on the symmetric matrix (m00 m10 m20 / m11 m21 / m22)
__device__ __forceinline__ void compute_roots(float m00, float m10, float m20, float m11, float m21, float m22, float (&r)[3])
{
    // left-to-right, uncontracted (see normal_pca)
    const float c0 = __fsub_rn(__fsub_rn(__fsub_rn(__fadd_rn(__fmul_rn(__fmul_rn(m00, m11), m22), __fmul_rn(__fmul_rn(__fmul_rn(2.0f, m10), m20), m21)),
                                                   __fmul_rn(__fmul_rn(m00, m21), m21)), __fmul_rn(__fmul_rn(m11, m20), m20)), __fmul_rn(__fmul_rn(m22, m10), m10));
    const float c1 = __fsub_rn(__fadd_rn(__fsub_rn(__fadd_rn(__fsub_rn(__fmul_rn(m00, m11), __fmul_rn(m10, m10)), __fmul_rn(m00, m22)), __fmul_rn(m20, m20)),
                                         __fmul_rn(m11, m22)), __fmul_rn(m21, m21));
    const float c2 = __fadd_rn(__fadd_rn(m00, m11), m22);
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

// geometry.glsl:190-244, window 3, over the filtered metric depth.  The accumulation runs in the shader's
// order (x outer, y inner); depth(qx,qy) is a callable so callers can serve it from shared memory.
template <typename DepthAt>
__device__ __forceinline__ float3 normal_pca(const PrepArgs& a, DepthAt depth_at, int px, int py, float vz, int coord_variant = 0)
{
    // The covariance is a difference of nearly equal numbers (E[x^2] - E[x]^2 with |x| ~ 1 m and a spread of
    // centimetres), so the normal inherits ~1e-3 of relative round-off: only a bit-identical evaluation order
    // reproduces the oracle.  Hence explicit _rn arithmetic (no FMA contraction) for the sums and the covariance.
    // The nine sums as four packed pairs + one scalar (add2_rn, common.cuh: each half rounded like the scalar operation, and every sum
    // keeps its order): (XX, XY), (Xz, YY), (Yz, zz), (X, Y), z.  The sample coordinates of the window come from the tables once per column
    // (x) and once per pixel (the <= 7 rows).
    float2 s01 = make_float2(0.f, 0.f), s23 = s01, s45 = s01, s67 = s01;
    float a8 = 0.f;
    int N = 0;
    const WinTab tx = a.wx[coord_variant], ty = a.wy[coord_variant];
    const int lx0 = __ldg(tx.first + px), lnx = __ldg(tx.count + px), ly0 = __ldg(ty.first + py), lny = __ldg(ty.count + py);
    float dy[kWinMax - 1];
#pragma unroll
    for (int iy = 0; iy < kWinMax - 1; ++iy) dy[iy] = iy < lny ? __fsub_rn(__ldg(ty.coord + py * kWinMax + iy), a.cy) : 0.f;
    for (int ix = 0; ix < lnx; ++ix) {
        const float dx = __fsub_rn(__ldg(tx.coord + px * kWinMax + ix), a.cx);
        const int qx = lx0 + ix;
#pragma unroll
        for (int iy = 0; iy < kWinMax - 1; ++iy) {
            if (iy >= lny) break;
            const float z = depth_at(qx, ly0 + iy);
            if (z > 0.3f && fabsf(__fsub_rn(z, vz)) < 0.05f) {
                const float X = __fmul_rn(__fmul_rn(dx, z), a.icx), Y = __fmul_rn(__fmul_rn(dy[iy], z), a.icy);
                // scalar products: ptxas fuses a packed product that feeds a packed sum into one FFMA2 even when both carry .rn
                s01 = add2_rn(s01, make_float2(__fmul_rn(X, X), __fmul_rn(X, Y)));
                s23 = add2_rn(s23, make_float2(__fmul_rn(X, z), __fmul_rn(Y, Y)));
                s45 = add2_rn(s45, make_float2(__fmul_rn(Y, z), __fmul_rn(z, z)));
                s67 = add2_rn(s67, make_float2(X, Y));
                a8 = __fadd_rn(a8, z);
                ++N;
            }
        }
    }
    float a0 = s01.x, a1 = s01.y, a2 = s23.x, a3 = s23.y, a4 = s45.x, a5 = s45.y, a6 = s67.x, a7 = s67.y;
    if (N < 8) return make_float3(0.f, 0.f, 0.f);
    const float fn = (float)N;
    a0 = __fdiv_rn(a0, fn); a1 = __fdiv_rn(a1, fn); a2 = __fdiv_rn(a2, fn); a3 = __fdiv_rn(a3, fn); a4 = __fdiv_rn(a4, fn);
    a5 = __fdiv_rn(a5, fn); a6 = __fdiv_rn(a6, fn); a7 = __fdiv_rn(a7, fn); a8 = __fdiv_rn(a8, fn);
    const float c00 = __fsub_rn(a0, __fmul_rn(a6, a6)), c10 = __fsub_rn(a1, __fmul_rn(a6, a7)), c20 = __fsub_rn(a2, __fmul_rn(a6, a8)),
                c11 = __fsub_rn(a3, __fmul_rn(a7, a7)), c21 = __fsub_rn(a4, __fmul_rn(a7, a8)), c22 = __fsub_rn(a5, __fmul_rn(a8, a8));
    const float scale = fmaxf(fmaxf(fmaxf(c00, c10), fmaxf(c20, c11)), fmaxf(c21, c22));
    float ev[3];
    compute_roots(c00, c10, c20, c11, c21, c22, ev);
    const float eigenvalue = ev[0] * scale;
    const float s00 = __fsub_rn(__fdiv_rn(c00, scale), eigenvalue), s10 = __fdiv_rn(c10, scale), s20 = __fdiv_rn(c20, scale),
                s11 = __fsub_rn(__fdiv_rn(c11, scale), eigenvalue), s21 = __fdiv_rn(c21, scale), s22 = __fsub_rn(__fdiv_rn(c22, scale), eigenvalue);
    const float3 r0 = make_float3(s00, s10, s20), r1 = make_float3(s10, s11, s21), r2 = make_float3(s20, s21, s22);
    auto cross_rn = [](float3 p, float3 q) {
        return make_float3(__fsub_rn(__fmul_rn(p.y, q.z), __fmul_rn(p.z, q.y)), __fsub_rn(__fmul_rn(p.z, q.x), __fmul_rn(p.x, q.z)),
                           __fsub_rn(__fmul_rn(p.x, q.y), __fmul_rn(p.y, q.x)));
    };
    const float3 v1 = cross_rn(r0, r1), v2 = cross_rn(r0, r2), v3 = cross_rn(r1, r2);
    const float l1 = norm(v1), l2 = norm(v2), l3 = norm(v3);
    float3 n = (l1 >= l2 && l1 >= l3) ? v1 : (l2 >= l1 && l2 >= l3) ? v2 : v3;
    if (n.z < 0) n = make_float3(-n.x, -n.y, -n.z);
    const float len = norm(n);
    return make_float3(n.x / len, n.y / len, n.z / len);
}

// ---- depth_vertex_normal_radius.frag:23-68 --------------------------------------------------------------
__global__ void __launch_bounds__(256) vertex_normal_radius_kernel(PrepArgs a, const float* __restrict__ metric, const float* __restrict__ metric_filtered,
                                                                   float4* __restrict__ vertex_raw, float4* __restrict__ vertex_filtered,
                                                                   float4* __restrict__ normal, float* __restrict__ radius)
{
    pdl_wait();
    constexpr int TW = 32, TH = 8, R = 3;
    __shared__ float s_d[TH + 2 * R][TW + 2 * R];
    const int W = a.cols, H = a.rows;
    const int x0 = blockIdx.x * TW - R, y0 = blockIdx.y * TH - R;
    for (int t = threadIdx.x; t < (TH + 2 * R) * (TW + 2 * R); t += 256) {
        const int sy = t / (TW + 2 * R), sx = t - sy * (TW + 2 * R);
        const int gx = x0 + sx, gy = y0 + sy;
        s_d[sy][sx] = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? __ldg(metric_filtered + (size_t)gy * W + gx) : 0.f;
    }
    __syncthreads();
    const int px = blockIdx.x * TW + (threadIdx.x & 31), py = blockIdx.y * TH + (threadIdx.x >> 5);
    if (px >= W || py >= H) return;
    const size_t o = (size_t)py * W + px;
    const float z = __ldg(metric + o), zf = s_d[py - y0][px - x0];
    float3 v = make_float3(((float)px - a.cx) * z * a.icx, ((float)py - a.cy) * z * a.icy, z);
    float3 vf = make_float3(((float)px - a.cx) * zf * a.icx, ((float)py - a.cy) * zf * a.icy, zf);
    float3 n = make_float3(0.f, 0.f, 0.f);
    if (a.pca) n = normal_pca(a, [&](int qx, int qy) { return s_d[qy - y0][qx - x0]; }, px, py, zf);
    float rad = a.radiusMultiplier * get_radius(a.icx, a.icy, vf.z, n.z);
    if (norm(n) < 0.3f || v.z < 0.3f || vf.z < 0.3f) { v = vf = n = make_float3(0.f, 0.f, 0.f); rad = 0.f; }
    const float max_dist = sqrtf(((float)H * 0.5f) * ((float)H * 0.5f) + ((float)W * 0.5f) * ((float)W * 0.5f));
    vertex_raw[o] = make_float4(v.x, v.y, v.z, confidence_fn(a.cx, a.cy, (float)px + 0.5f, (float)py + 0.5f, max_dist, 1.0f));
    vertex_filtered[o] = make_float4(vf.x, vf.y, vf.z, 1.0f);
    normal[o] = make_float4(n.x, n.y, n.z, rad);
    if (radius) radius[o] = rad;
}

// ---- depth_curvature_gradient.frag:28-142 ------------------------------------------------------------------
// One thread per pixel; the 7x7 neighbourhood of (filtered vertex, PCA normal + radius) is served from a
// shared-memory halo tile; gradient (hrbfbase.glsl:147-166) and the symmetric third-derivative contraction
// (hrbfHessianMatrix :168-195 -- only g[0,1,2,4,5,8] are accumulated there) are summed in one neighbour loop.
__global__ void __launch_bounds__(128) curvature_gradient_kernel(PrepArgs a, const float4* __restrict__ vertex_filtered, const float4* __restrict__ normal,
                                                                 float4* __restrict__ curv1, float4* __restrict__ curv2, float* __restrict__ gradient_mag,
                                                                 float4* __restrict__ normal_opt)
{
    pdl_wait();
    constexpr int TW = 16, TH = 8, R = 3;
    __shared__ float4 s_v[TH + 2 * R][TW + 2 * R];      // filtered vertex; z = -1000 when the pixel can never be a neighbour
    __shared__ float4 s_n[TH + 2 * R][TW + 2 * R];      // 10 n (the HRBF coefficient), w = 1 / rho^2
    const int W = a.cols, H = a.rows, win = a.curvWin;
    const int x0 = blockIdx.x * TW - R, y0 = blockIdx.y * TH - R;
    const int px = blockIdx.x * TW + (threadIdx.x & 15), py = blockIdx.y * TH + (threadIdx.x >> 4);
    float4 vf = make_float4(0.f, 0.f, 0.f, 0.f), vn = vf;
    for (int t = threadIdx.x; t < (TH + 2 * R) * (TW + 2 * R); t += 128) {
        const int sy = t / (TW + 2 * R), sx = t - sy * (TW + 2 * R);
        const int gx = x0 + sx, gy = y0 + sy;
        float4 v = make_float4(0.f, 0.f, -1000.f, 0.f), n = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gx >= 0 && gx < W && gy >= 0 && gy < H) {
            const float4 v_ = __ldg(vertex_filtered + (size_t)gy * W + gx), n_ = __ldg(normal + (size_t)gy * W + gx);
            // the pose-independent half of the neighbour test (depth_curvature_gradient.frag:60-64)
            if (v_.z > 0.3f && sqrtf(n_.x * n_.x + n_.y * n_.y + n_.z * n_.z) > 0.8f) {
                v = v_;
                n = make_float4(10.0f * n_.x, 10.0f * n_.y, 10.0f * n_.z, 1.0f / (n_.w * n_.w));
            }
        }
        s_v[sy][sx] = v; s_n[sy][sx] = n;
    }
    if (px < W && py < H) { vf = __ldg(vertex_filtered + (size_t)py * W + px); vn = __ldg(normal + (size_t)py * W + px); }
    __syncthreads();
    if (px >= W || py >= H) return;
    const size_t o = (size_t)py * W + px;
    float4 kmax = make_float4(0.f, 0.f, 0.f, 1000.0f), kmin = kmax, nopt = make_float4(0.f, 0.f, 0.f, 0.f);
    float gm = 0.f;
    if (vf.z > 0.3f && sqrtf(vn.x * vn.x + vn.y * vn.y + vn.z * vn.z) > 0.5f) {
        float k1 = 1000.0f, k2 = 1000.0f;
        float3 pmax = make_float3(0.f, 0.f, 0.f), pmin = pmax;
        (void)win;       // the tables are built for curvWin (hrbf_frame_create)
        const int qx0 = __ldg(a.wx[0].first + px), qx1 = qx0 + __ldg(a.wx[0].count + px) - 1, qy0 = __ldg(a.wy[0].first + py), qy1 = qy0 + __ldg(a.wy[0].count + py) - 1;
        int N = 0;
        // Two neighbours per trip on packed fp32 pairs (FFMA2 / FMUL2 / FADD2: half the issue slots for the same arithmetic; the kernel is
        // issue-bound).  A rejected neighbour (depth gap, outside the support, the pixel itself) stays in its lane with a harmless u and a
        // zero mask on everything it adds; the pixel itself (d2 == 0: getWeightH = -20/T2 I, getWeightT = 0) is added on its own below.
        float2 gx2 = make_float2(0.f, 0.f), gy2 = gx2, gz2 = gx2;                                   // gradient, one partial sum per lane
        float2 h0 = gx2, h1 = gx2, h2 = gx2, h4 = gx2, h5 = gx2, h8 = gx2;                          // "Hessian" entries g[0], g[1], g[2], g[4], g[5], g[8]
        const int ny = qy1 - qy0 + 1, ncell = (qx1 - qx0 + 1) * ny;
        const float2 one2 = make_float2(1.0f, 1.0f), neg2 = make_float2(-1.0f, -1.0f);
        const float2 pfx = make_float2(vf.x, vf.x), pfy = make_float2(vf.y, vf.y), pfz = make_float2(vf.z, vf.z);
        auto bc = [](float a) { return make_float2(a, a); };
        int wx_ = qx0 - x0, wy_ = qy0 - y0;                 // tile cell of the next window cell, in the shader's order (x outer, y inner)
        const int wy_end = qy0 - y0 + ny;
        for (int c = 0; c < ncell; c += 2) {
            const int cax = wx_, cay = wy_;
            if (++wy_ == wy_end) { wy_ = qy0 - y0; ++wx_; }
            const bool has_b = c + 1 < ncell;
            const int cbx = has_b ? wx_ : cax, cby = has_b ? wy_ : cay;
            if (++wy_ == wy_end) { wy_ = qy0 - y0; ++wx_; }
            const float4 va = s_v[cay][cax], vb = s_v[cby][cbx];
            const float4 na = s_n[cay][cax], nb = s_n[cby][cbx];      // (sx, sy, sz) = 10 n, w = 1 / rho^2
            const bool ina = fabsf(va.z - vf.z) < 0.10f, inb = has_b && fabsf(vb.z - vf.z) < 0.10f;
            N += (ina ? 1 : 0) + (inb ? 1 : 0);
            const float2 vx = __ffma2_rn(make_float2(va.x, vb.x), neg2, pfx), vy = __ffma2_rn(make_float2(va.y, vb.y), neg2, pfy), vz = __ffma2_rn(make_float2(va.z, vb.z), neg2, pfz);
            const float2 sx = make_float2(na.x, nb.x), sy = make_float2(na.y, nb.y), sz = make_float2(na.z, nb.z), iT2 = make_float2(na.w, nb.w);
            const float2 d2 = __ffma2_rn(vz, vz, __ffma2_rn(vy, vy, __fmul2_rn(vx, vx)));
            const float2 u = __fmul2_rn(d2, iT2);                                                  // (|v| / rho)^2
            const bool oka = ina && !(u.x > 1.0f) && d2.x != 0.0f, okb = inb && !(u.y > 1.0f) && d2.y != 0.0f;
            if (ina && d2.x == 0.0f) { const float h = -20.0f * na.w; gx2.x -= na.x * h; gy2.x -= na.y * h; gz2.x -= na.z * h; }
            if (inb && d2.y == 0.0f) { const float h = -20.0f * nb.w; gx2.y -= nb.x * h; gy2.y -= nb.y * h; gz2.y -= nb.z * h; }
            if (!(oka || okb)) continue;
            // a rejected lane: u = 1/4 and 1 / rho^2 = 0, which zeroes t1 and s3 and with them everything the lane adds
            const float2 us = make_float2(oka ? u.x : 0.25f, okb ? u.y : 0.25f), iT2m = make_float2(oka ? na.w : 0.0f, okb ? nb.w : 0.0f);
            // With s = 10 n, dvs = v.s, r = |v|/rho, q = 1 - r (hrbfbase.glsl:37-69, 72-123, 147-195):
            //   gradient  -= H s,  H = t1 (3 v v^T + t2 I)          ->  t1 (3 dvs v + t2 s)
            //   g[ij]     -= sum_k T_ijk s_k  with the shader's T (its 27 entries are restated in the oracle); collecting
            //   terms:  diagonal ii : A v_i^2 + B + C s_i v_i      off-diagonal ij (i<j) : A v_i v_j + D s_i v_j + E s_j v_i
            const float2 ir = make_float2(rsqrtf(us.x), rsqrtf(us.y)), r = __fmul2_rn(us, ir), q = __ffma2_rn(r, neg2, one2);
            const float2 dvs = __ffma2_rn(vz, sz, __ffma2_rn(vy, sy, __fmul2_rn(vx, sx)));
            const float2 iT4 = __fmul2_rn(iT2m, iT2m), qq = __fmul2_rn(q, q);
            {
                const float2 t1 = __fmul2_rn(__fmul2_rn(bc(20.0f), qq), __fmul2_rn(iT4, ir));
                const float2 t2 = __fmul2_rn(__fmul2_rn(__fmul2_rn(r, q), neg2), __fmul2_rn(d2, __fmul2_rn(ir, ir)));      // -r q T2  (T2 = d2 / u)
                const float2 c3 = __fmul2_rn(__fmul2_rn(bc(3.0f), t1), dvs), ct = __fmul2_rn(t1, t2);
                gx2 = __ffma2_rn(__ffma2_rn(c3, vx, __fmul2_rn(ct, sx)), neg2, gx2);
                gy2 = __ffma2_rn(__ffma2_rn(c3, vy, __fmul2_rn(ct, sy)), neg2, gy2);
                gz2 = __ffma2_rn(__ffma2_rn(c3, vz, __fmul2_rn(ct, sz)), neg2, gz2);
            }
            {
                const float2 s2 = __fadd2_rn(__fadd2_rn(r, bc(-2.0f)), ir);
                const float2 s3 = __fmul2_rn(bc(-60.0f), iT4);                                      // negated: the entries are subtracted
                const float2 kap = __fmul2_rn(__ffma2_rn(__fmul2_rn(ir, ir), neg2, one2), __fmul2_rn(iT2m, ir));
                const float2 tss_pi = __fmul2_rn(qq, ir);                                           // T2 q^2 / (T2 r)
                const float2 A = __fmul2_rn(__fmul2_rn(s3, kap), dvs), B = __fmul2_rn(__fmul2_rn(s3, s2), dvs), Cc = __fmul2_rn(s3, __fadd2_rn(tss_pi, s2)),
                             D = __fmul2_rn(s3, tss_pi), E = __fmul2_rn(s3, s2);
                h0 = __fadd2_rn(h0, __ffma2_rn(A, __fmul2_rn(vx, vx), __ffma2_rn(Cc, __fmul2_rn(sx, vx), B)));
                h4 = __fadd2_rn(h4, __ffma2_rn(A, __fmul2_rn(vy, vy), __ffma2_rn(Cc, __fmul2_rn(sy, vy), B)));
                h8 = __fadd2_rn(h8, __ffma2_rn(A, __fmul2_rn(vz, vz), __ffma2_rn(Cc, __fmul2_rn(sz, vz), B)));
                h1 = __fadd2_rn(h1, __ffma2_rn(A, __fmul2_rn(vx, vy), __ffma2_rn(D, __fmul2_rn(sx, vy), __fmul2_rn(E, __fmul2_rn(sy, vx)))));
                h2 = __fadd2_rn(h2, __ffma2_rn(A, __fmul2_rn(vx, vz), __ffma2_rn(D, __fmul2_rn(sx, vz), __fmul2_rn(E, __fmul2_rn(sz, vx)))));
                h5 = __fadd2_rn(h5, __ffma2_rn(A, __fmul2_rn(vy, vz), __ffma2_rn(D, __fmul2_rn(sy, vz), __fmul2_rn(E, __fmul2_rn(sz, vy)))));
            }
        }
        const float gx = gx2.x + gx2.y, gy = gy2.x + gy2.y, gz = gz2.x + gz2.y;
        const float H0 = h0.x + h0.y, H1 = h1.x + h1.y, H2 = h2.x + h2.y, H4 = h4.x + h4.y, H5 = h5.x + h5.y, H8 = h8.x + h8.y;
        if (N > 15) {
            gm = fabsf(gx * vn.x + gy * vn.y + gz * vn.z);
            const float gl = sqrtf(gx * gx + gy * gy + gz * gz);
            nopt = make_float4(gx / gl, gy / gl, gz / gl, vn.w);
            const float h_x = -gx / gz, h_y = -gy / gz;
            const float gz3 = gz * gz * gz;
            const float h_xx = (2 * gx * gz * H2 - gx * gx * H8 - gz * gz * H0) / gz3;
            const float h_xy = (gx * gz * H5 + gy * gz * H2 - gx * gy * H8 - gz * gz * H1) / gz3;
            const float h_yy = (2 * gy * gz * H5 - gy * gy * H8 - gz * gz * H4) / gz3;
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
                const float3 A = make_float3(__fadd_rn(1.0f, __fmul_rn(lmax, 0.0f)), lmax, h_x + lmax * h_y), B = make_float3(__fadd_rn(1.0f, __fmul_rn(lmin, 0.0f)), lmin, h_x + lmin * h_y);
                const float la = norm(A), lb = norm(B);
                pmax = make_float3(A.x / la, A.y / la, A.z / la);
                pmin = make_float3(B.x / lb, B.y / lb, B.z / lb);
            }
        }
        kmax = make_float4(pmax.x, pmax.y, pmax.z, k1);
        kmin = make_float4(pmin.x, pmin.y, pmin.z, k2);
    }
    curv1[o] = kmax; curv2[o] = kmin; gradient_mag[o] = gm; normal_opt[o] = nopt;
}

// ---- depth_confidence_evaluation.frag ; weighting is read from device memory (no host round trip) ---------
__global__ void confidence_kernel(PrepArgs a, const float* __restrict__ gradient_mag, const float* __restrict__ weighting_dev,
                                  int useConfEval, float epsilon, float* __restrict__ confidence)
{
    const int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y * blockDim.y + threadIdx.y;
    if (px >= a.cols || py >= a.rows) return;
    const size_t o = (size_t)py * a.cols + px;
    const float max_dist = sqrtf(((float)a.rows * 0.5f) * ((float)a.rows * 0.5f) + ((float)a.cols * 0.5f) * ((float)a.cols * 0.5f));
    float c = confidence_fn(a.cx, a.cy, (float)px + 0.5f, (float)py + 0.5f, max_dist, __ldg(weighting_dev));
    if (useConfEval > 0) c = c * expf(-epsilon / sqrtf(__ldg(gradient_mag + o)));
    confidence[o] = c;
}

// ---- fill_vertex / fill_normal / fill_curvature / fill_rgb (FillIn.cpp; every target is cleared to 0 first) ---
struct FillArgs {
    const float4 *eVertex, *eNormal, *eK1, *eK2; const float* eIcpW; const uchar4* eImage;      // existing = HRBF prediction
    const float4 *vertexFiltered, *normal, *k1, *k2; const float* confidence; const unsigned char* rgb;   // raw frame (rgb: RGB8)
    float4 *oVertex, *oNormal, *oK1, *oK2; float* oIcpW; uchar4* oImage;
    int n, passthrough;
    float lambda, curvThr;
    // frame pipeline: when non-null the confidence of depth_confidence_evaluation.frag is evaluated in place from this device scalar
    // instead of being read from the CONFIDENCE texture (same expression, confidence_kernel below)
    const float* weighting; int cols, rows; float cx, cy;
    // frame pipeline: the fill-in maps are only ever read when the prediction is NOT dense enough (HRBFFusion.cpp:1069-1086); when the
    // prediction kernel's sample count says it is, the pass is skipped (non-null dense_count)
    const unsigned int* dense_count; float dense_thresh;
};
__global__ void fill_in_kernel(FillArgs f)
{
    pdl_wait();
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= f.n) return;
    if (f.dense_count != nullptr && (float)*f.dense_count / (float)((f.cols / 20) * (f.rows / 20)) > f.dense_thresh) return;
    const bool pass = f.passthrough == 1;
    const float4 s = __ldg(f.eVertex + o);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    float w = 0.f;
    const float4 rk1 = __ldg(f.k1 + o), rk2 = __ldg(f.k2 + o);
    if (s.z == 0 || pass) {
        if (rk1.w > -f.curvThr && rk1.w < f.curvThr && rk2.w > -f.curvThr && rk2.w < f.curvThr) {
            const float4 fv = __ldg(f.vertexFiltered + o);
            float vConf;
            if (f.weighting != nullptr) {
                const int py = o / f.cols, px = o - py * f.cols;
                const float max_dist = sqrtf(((float)f.rows * 0.5f) * ((float)f.rows * 0.5f) + ((float)f.cols * 0.5f) * ((float)f.cols * 0.5f));
                vConf = confidence_fn(f.cx, f.cy, (float)px + 0.5f, (float)py + 0.5f, max_dist, __ldg(f.weighting));
            } else vConf = __ldg(f.confidence + o);
            const float cmax = fmaxf(fabsf(rk1.w), fabsf(rk2.w));
            w = (1.0f / (fv.z * fv.z)) * (vConf / 256.0f + expf(-0.5f * (f.lambda * f.lambda) / (cmax * cmax)));
            v = make_float4(fv.x, fv.y, fv.z, vConf);
        }
    } else { v = s; w = __ldg(f.eIcpW + o); }
    f.oVertex[o] = v; f.oIcpW[o] = w;
    const float4 en = __ldg(f.eNormal + o);
    f.oNormal[o] = (sqrtf(en.x * en.x + en.y * en.y + en.z * en.z) < 0.8f || pass) ? __ldg(f.normal + o) : en;
    const float4 ek1 = __ldg(f.eK1 + o), ek2 = __ldg(f.eK2 + o);
    const bool rawk = ek1.w > 300 || ek2.w > 300 || pass;
    f.oK1[o] = rawk ? rk1 : ek1; f.oK2[o] = rawk ? rk2 : ek2;
    const uchar4 ei = __ldg(f.eImage + o);
    f.oImage[o] = ((ei.x == 0 && ei.y == 0 && ei.z == 0) || pass) ? make_uchar4(f.rgb[3 * o], f.rgb[3 * o + 1], f.rgb[3 * o + 2], 255) : ei;
}

// ---- Resize (1/20, texel centres) + denseEnough : flag[0] = 1 when the prediction must be filled in -----------
__global__ void should_fill_kernel(const float4* __restrict__ vertex, int rows, int cols, float thresh, int* __restrict__ flag)
{
    const int f = 20, w = cols / f, h = rows / f;
    int sum = 0;
    for (int k = threadIdx.x; k < w * h; k += blockDim.x) {
        const int j = k / w, i = k - j * w;
        sum += __ldg(vertex + (size_t)(f * j + f / 2) * cols + (f * i + f / 2)).z > 0 ? 1 : 0;
    }
    sum = __reduce_add_sync(0xffffffffu, sum);
    __shared__ int s_sum[32];
    if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) tot += s_sum[k];
        const float per = (float)tot / (float)(h * w);
        flag[0] = per > thresh ? 0 : 1;
    }
}

// RGB8 -> RGBA8 (the GL upload of the RGB texture, alpha = 1)
__global__ void rgb_to_rgba_kernel(int n, const unsigned char* __restrict__ rgb, uchar4* __restrict__ rgba)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rgba[i] = make_uchar4(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2], 255);
}

}  // namespace hrbf
