// model_kernels.cuh -- sm_100a kernels for SURVEY.md section 8 rows 8-9: the surfel map.
// Replaces the GL transform-feedback passes of Core/src/GlobalModel.cpp:
//   initialise (:214-288, init_unstableTex.vert/.geom)   -> init_flags_kernel + compacting scan
//   fuse       (:355-549, data.vert/.geom/.frag, update.vert) -> fuse_associate_kernel (1/4 of the pixels: association,
//              PCA normal, new-surfel record, 32-bit atomicMin "first primitive wins") + fuse_merge_kernel (in-place
//              confidence-weighted merge of the winners only: no full-map pass, no 4596^2 update textures)
//   clean      (:551-688, copy_unstable.vert/.geom)      -> clean_flags_kernel + block scan + clean_scatter_kernel
//              (order-preserving stream compaction; the count never leaves the device)
// Surfel record = 5 x float4 (80 B), the reference's VBO layout (Shaders/Vertex.cpp:20-44).
#pragma once
#include "prep_kernels.cuh"

namespace hrbf {

struct ModelArgs {
    int cols, rows;
    float cx, cy, fx, fy, icx, icy;
    float maxDepth, confThreshold, radiusMultiplier, curvThr;
    int pca, cleanWindow;
};

constexpr int kScanBlock = 256;
constexpr unsigned int kNoWinner = 0xffffffffu;

__device__ __forceinline__ float3 pose_apply(const float* P /* R[9], t[3] */, float3 v)
{
    return make_float3(((P[0] * v.x + P[1] * v.y) + P[2] * v.z) + P[9], ((P[3] * v.x + P[4] * v.y) + P[5] * v.z) + P[10],
                       ((P[6] * v.x + P[7] * v.y) + P[8] * v.z) + P[11]);
}
__device__ __forceinline__ float3 pose_rotate(const float* P, float3 v)
{
    return make_float3((P[0] * v.x + P[1] * v.y) + P[2] * v.z, (P[3] * v.x + P[4] * v.y) + P[5] * v.z, (P[6] * v.x + P[7] * v.y) + P[8] * v.z);
}
__device__ __forceinline__ float encode_rgb8(const unsigned char* c) { return (float)((((((int)c[0]) << 8) + (int)c[1]) << 8) + (int)c[2]); }
__device__ __forceinline__ float encode_color(float3 c)
{
    int rgb = (int)roundf(c.x * 255.0f);
    rgb = (rgb << 8) + (int)roundf(c.y * 255.0f);
    rgb = (rgb << 8) + (int)roundf(c.z * 255.0f);
    return (float)rgb;
}
__device__ __forceinline__ float3 decode_color(float c)
{
    const int i = (int)c;
    return make_float3((float)(i >> 16 & 0xFF) / 255.0f, (float)(i >> 8 & 0xFF) / 255.0f, (float)(i & 0xFF) / 255.0f);
}

__global__ void pose_inverse_kernel(const float* __restrict__ pose, float* __restrict__ inv)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    pose_inverse_dev(pose, inv);
}
__global__ void velocity_weighting_kernel(const float* __restrict__ curr, const float* __restrict__ last, float weightMultiplier, float* __restrict__ weighting)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    weighting[0] = velocity_weighting_dev(curr, last, weightMultiplier);
}

// ------------------------------------------------------------------ generic order-preserving compaction ---
// flags (1 byte per item, written by a *_flags_kernel together with block_counts) -> block_offsets, total
// The exclusive scan of the per-block counts by ONE thread block of kThreads threads (all of them call it), + the total (clamped to the
// capacity, overflow flagged).
template <int kThreads>
__device__ __forceinline__ void scan_blocks_cta(const unsigned int* block_counts, unsigned int* block_offsets, unsigned int n_items,
                                                unsigned int capacity, unsigned int* total_out, unsigned int* overflow)
{
    __shared__ unsigned int s_warp[kThreads / 32];
    __shared__ unsigned int s_carry;
    const unsigned int nb = (n_items + kScanBlock - 1) / kScanBlock;
    if (threadIdx.x == 0) s_carry = 0u;
    __syncthreads();
    for (unsigned int base = 0; base < nb; base += kThreads) {
        const unsigned int b = base + threadIdx.x;
        const unsigned int v = b < nb ? __ldcg(block_counts + b) : 0u;
        unsigned int x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned int y = __shfl_up_sync(0xffffffffu, x, d); if ((threadIdx.x & 31) >= d) x += y; }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned int w = threadIdx.x < kThreads / 32 ? s_warp[threadIdx.x] : 0u;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const unsigned int y = __shfl_up_sync(0xffffffffu, w, d); if (threadIdx.x >= d) w += y; }
            if (threadIdx.x < kThreads / 32) s_warp[threadIdx.x] = w;
        }
        __syncthreads();
        const unsigned int incl = x + ((threadIdx.x >> 5) ? s_warp[(threadIdx.x >> 5) - 1] : 0u) + s_carry;
        if (b < nb) block_offsets[b] = incl - v;
        __syncthreads();
        if (threadIdx.x == kThreads - 1) s_carry = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        unsigned int tot = s_carry;
        if (tot > capacity) { tot = capacity; *overflow = 1u; }
        *total_out = tot;
    }
}
__global__ void __launch_bounds__(1024) scan_blocks_kernel(const unsigned int* __restrict__ block_counts, unsigned int* __restrict__ block_offsets,
                                                           const unsigned int* __restrict__ n_items_dev, unsigned int n_items_add,
                                                           unsigned int capacity, unsigned int* __restrict__ total_out, unsigned int* __restrict__ overflow)
{
    pdl_wait();
    scan_blocks_cta<1024>(block_counts, block_offsets, (n_items_dev ? *n_items_dev : 0u) + n_items_add, capacity, total_out, overflow);
}

// exclusive position of a flagged item inside its block (ballot + warp prefix)
__device__ __forceinline__ unsigned int block_rank(bool flag, unsigned int* s_warp /* [kScanBlock/32] */)
{
    const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int bal = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    unsigned int off = 0;
    for (unsigned int w = 0; w < warp; ++w) off += s_warp[w];
    return off + __popc(bal & ((1u << lane) - 1u));
}

// ------------------------------------------------------------------ initialise ---
struct InitArgs {
    const float4 *vertexRaw, *normal, *curv1, *curv2; const unsigned char* rgb; const float* gradientMag;
    const float* pose;      // device R[9], t[3]
    int useConfEval; float epsilon;
};
// item i = px * rows + py (the reference's uv buffer order, GlobalModel.cpp:89-96)
__device__ __forceinline__ bool init_record(const ModelArgs& m, const InitArgs& a, const float* P, int i, float4 (&rec)[5])
{
    const int px = i / m.rows, py = i - px * m.rows;
    const size_t o = (size_t)py * m.cols + px;
    const float4 vl = __ldg(a.vertexRaw + o), nl = __ldg(a.normal + o), k1 = __ldg(a.curv1 + o), k2 = __ldg(a.curv2 + o);
    const float3 g = pose_apply(P, make_float3(vl.x, vl.y, vl.z));
    const float max_dist = sqrtf(((float)m.rows * 0.5f) * ((float)m.rows * 0.5f) + ((float)m.cols * 0.5f) * ((float)m.cols * 0.5f));
    float conf = confidence_fn(m.cx, m.cy, (float)px + 0.5f, (float)py + 0.5f, max_dist, 1.0f);
    if (a.useConfEval > 0) conf = conf * expf(-a.epsilon / sqrtf(__ldg(a.gradientMag + o)));
    const float3 n = pose_rotate(P, make_float3(nl.x, nl.y, nl.z));
    rec[0] = make_float4(g.x, g.y, g.z, conf);
    rec[1] = make_float4(encode_rgb8(a.rgb + 3 * o), 0.0f, 1.0f, 1.0f);
    rec[2] = make_float4(n.x, n.y, n.z, nl.w);
    rec[3] = k1; rec[4] = k2;
    return sqrtf(n.x * n.x + n.y * n.y + n.z * n.z) > 0.5f && k1.w > -m.curvThr && k1.w < m.curvThr && k2.w > -m.curvThr && k2.w < m.curvThr;
}
__global__ void __launch_bounds__(kScanBlock) init_flags_kernel(ModelArgs m, InitArgs a, unsigned char* __restrict__ flags, unsigned int* __restrict__ block_counts)
{
    const int i = blockIdx.x * kScanBlock + threadIdx.x;
    float P[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) P[k] = __ldg(a.pose + k);
    bool f = false;
    if (i < m.cols * m.rows) { float4 rec[5]; f = init_record(m, a, P, i, rec); flags[i] = f ? 1 : 0; }
    const int c = __syncthreads_count(f);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = c;
}
__global__ void __launch_bounds__(kScanBlock) init_scatter_kernel(ModelArgs m, InitArgs a, const unsigned char* __restrict__ flags,
                                                                  const unsigned int* __restrict__ block_offsets, unsigned int capacity, float4* __restrict__ out)
{
    __shared__ unsigned int s_warp[kScanBlock / 32];
    const int i = blockIdx.x * kScanBlock + threadIdx.x;
    float P[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) P[k] = __ldg(a.pose + k);
    const bool f = i < m.cols * m.rows && flags[i];
    const unsigned int pos = block_offsets[blockIdx.x] + block_rank(f, s_warp);
    if (!f || pos >= capacity) return;
    float4 rec[5];
    init_record(m, a, P, i, rec);
#pragma unroll
    for (int k = 0; k < 5; ++k) out[5 * (size_t)pos + k] = rec[k];
}

// ------------------------------------------------------------------ fuse ---
struct FuseArgs {
    const unsigned char* rgb; const float *depthRaw, *depthFiltered; const float4 *curv1, *curv2; const float* confidence;
    const unsigned int* index; const float4 *vertConf, *normRad;
    const float* pose;               // device R[9], t[3]
    int time; float indexSubmap;
    const float* weighting;          // frame pipeline: non-null -> confidence evaluated in place (see FillArgs::weighting)
    const float4* normal_slot;       // frame pipeline: data.vert's PCA normal of every candidate pixel, per slot, computed ahead of the fuse by
                                     // fuse_normals_kernel (it depends on the camera frame alone); null: computed here
    float4* staging;                 // [(cols/2+1)*(rows/2+1)][5] : this frame's candidate records
    unsigned char* update_id;        // per slot: 0 none, 1 merge, 2 new
    unsigned int* best;              // per slot: surfel to merge with
    unsigned int* winner;            // per surfel: lowest slot that wants to update it (kNoWinner = none); self re-arming
};
__host__ __device__ inline int fuse_slots_x(int cols) { return (cols + 1) / 2; }
__host__ __device__ inline int fuse_slots_y(int rows) { return (rows + 1) / 2; }

// data.vert:63-110 recomputes the PCA normal of the filtered depth at each candidate pixel (geometry.glsl with the uv VBO's texture
// coordinates).  That needs the camera frame only, so the frame pipeline runs it ahead of the fuse, on its staging stream: same
// slots, same rejection tests, same function.
// A CTA covers 16 x 8 candidate pixels (a 32 x 16 pixel patch) and serves the 7 x 7 windows from a shared-memory tile of the filtered depth.
constexpr int kFnTX = 16, kFnTY = 8, kFnR = 3;
__global__ void __launch_bounds__(kFnTX * kFnTY) fuse_normals_kernel(ModelArgs m, PrepArgs pa, const float* __restrict__ depthRaw, const float* __restrict__ depthFiltered,
                                                                     const float4* __restrict__ curv1, const float4* __restrict__ curv2, int time, float4* __restrict__ out)
{
    pdl_wait();
    constexpr int SW = 2 * kFnTX + 2 * kFnR, SH = 2 * kFnTY + 2 * kFnR;
    __shared__ float s_d[SH][SW + 1];
    const int W = m.cols, H = m.rows, syn = fuse_slots_y(H), sxn = fuse_slots_x(W);
    const int par = time % 2;
    const int x0 = 2 * kFnTX * blockIdx.x + par - kFnR, y0 = 2 * kFnTY * blockIdx.y + par - kFnR;      // pixel of tile cell (0, 0)
    for (int t = threadIdx.x; t < SH * SW; t += kFnTX * kFnTY) {
        const int sy = t / SW, sx = t - sy * SW;
        const int gx = x0 + sx, gy = y0 + sy;
        s_d[sy][sx] = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? __ldg(depthFiltered + (size_t)gy * W + gx) : 0.f;
    }
    __syncthreads();
    const int cx = threadIdx.x % kFnTX, cy = threadIdx.x / kFnTX;
    const int xs = kFnTX * blockIdx.x + cx, ys = kFnTY * blockIdx.y + cy;          // slot coordinates
    if (xs >= sxn || ys >= syn) return;
    const int px = 2 * xs + par, py = 2 * ys + par;
    float3 n = make_float3(0.f, 0.f, 0.f);
    if (px < W && py < H && m.pca) {
        const size_t o = (size_t)py * W + px;
        const float z = __ldg(depthRaw + o), zf = s_d[py - y0][px - x0];
        const float4 k1 = __ldg(curv1 + o), k2 = __ldg(curv2 + o);
        if (z > 0.3f && z <= m.maxDepth && k1.w > -300.0f && k1.w < 300.0f && k2.w > -300.0f && k2.w < 300.0f)
            n = normal_pca(pa, [&](int qx, int qy) { return s_d[qy - y0][qx - x0]; }, px, py, zf, /* uv-VBO texcoords */ 1);
    }
    out[(size_t)xs * syn + ys] = make_float4(n.x, n.y, n.z, 0.f);
}

// One thread per candidate pixel (x%2 == t%2 && y%2 == t%2, data.vert:113).  Slot order = uv order (x outer, y inner).
__global__ void __launch_bounds__(128) fuse_associate_kernel(ModelArgs m, PrepArgs pa, FuseArgs f, const unsigned int* __restrict__ count_dev)
{
    pdl_wait();
    const int sxn = fuse_slots_x(m.cols), syn = fuse_slots_y(m.rows);
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= sxn * syn) return;
    const int par = f.time % 2;
    const int px = 2 * (slot / syn) + par, py = 2 * (slot % syn) + par;
    f.update_id[slot] = 0;
    if (px >= m.cols || py >= m.rows) return;
    const int W = m.cols, H = m.rows;
    const size_t o = (size_t)py * W + px;
    const float x = (float)px + 0.5f, y = (float)py + 0.5f;
    const float z = __ldg(f.depthRaw + o), zf = __ldg(f.depthFiltered + o);
    const float4 k1 = __ldg(f.curv1 + o), k2 = __ldg(f.curv2 + o);
    if (!(z > 0.3f && z <= m.maxDepth && k1.w > -300.0f && k1.w < 300.0f && k2.w > -300.0f && k2.w < 300.0f)) return;
    float3 n = make_float3(0.f, 0.f, 0.f);
    if (m.pca) {
        if (f.normal_slot != nullptr) { const float4 t = __ldg(f.normal_slot + slot); n = make_float3(t.x, t.y, t.z); }
        else n = normal_pca(pa, [&](int qx, int qy) { return __ldg(f.depthFiltered + (size_t)qy * W + qx); }, px, py, zf, /* uv-VBO texcoords */ 1);
    }
    const float nlen = norm(n);
    if (!(nlen > 0.8f)) return;

    float P[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) P[k] = __ldg(f.pose + k);
    const float3 vloc = make_float3((x - m.cx) * z * m.icx, (y - m.cy) * z * m.icy, z);
    const float3 g = pose_apply(P, vloc), ng = pose_rotate(P, n);

    const float xl = (x - m.cx) * m.icx, yl = (y - m.cy) * m.icy;
    const float lambda = sqrtf(xl * xl + yl * yl + 1);
    const float3 ray = make_float3(xl, yl, 1.0f);
    const float raylen = norm(ray);
    const unsigned int count = *count_dev;
    int counter = 0;
    float bestDist = 1000;
    unsigned int best = 0;
    // 4 x 4 half-pixel offsets around the pixel centre (data.vert:123-138) hit texels {px-1, px, px, px+1} x {py-1, py, py, py+1}.
    // A texel seen again can never replace the best (strict <) and `counter` is only tested for > 0, so a repeated texel is
    // skipped; the scan order of the distinct ones is kept.
    size_t qs[16];
    unsigned int cur[16];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const float ox = -1.0f + 0.5f * (float)a, oy = -1.0f + 0.5f * (float)b;
            const int sx = min(max((int)floorf(x + ox), 0), W - 1), sy = min(max((int)floorf(y + oy), 0), H - 1);
            qs[a * 4 + b] = (size_t)sy * W + sx;
        }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        bool first = true;
#pragma unroll
        for (int j = 0; j < 16; ++j) if (j < k) first = first && qs[j] != qs[k];
        cur[k] = first ? __ldg(f.index + qs[k]) : 0u;
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const size_t q = qs[a * 4 + b];
            const unsigned int current = cur[a * 4 + b];
            if (current > 0u) {
                const float4 vc = __ldg(f.vertConf + q);
                if (fabsf((vc.z * lambda) - (vloc.z * lambda)) < 0.05f) {
                    const float dist = norm(cross(ray, make_float3(vc.x, vc.y, vc.z))) / raylen;
                    const float4 nr = __ldg(f.normRad + q);
                    const float la = sqrtf(nr.x * nr.x + nr.y * nr.y + nr.z * nr.z);
                    const float ang = acosf((nr.x * n.x + nr.y * n.y + nr.z * n.z) / (la * nlen));
                    if (dist < bestDist && (fabsf(nr.z) < 0.75f || fabsf(ang) < 0.5f)) { counter++; bestDist = dist; best = current; }
                }
            }
        }
    const bool merge = counter > 0 && best < count;
    float4* rec = f.staging + 5 * (size_t)slot;
    float conf;
    if (f.weighting != nullptr) {
        const float max_dist = sqrtf(((float)H * 0.5f) * ((float)H * 0.5f) + ((float)W * 0.5f) * ((float)W * 0.5f));
        conf = confidence_fn(m.cx, m.cy, x, y, max_dist, __ldg(f.weighting));
    } else conf = __ldg(f.confidence + o);
    rec[0] = make_float4(g.x, g.y, g.z, conf);
    rec[1] = make_float4(encode_rgb8(f.rgb + 3 * o), f.indexSubmap, (float)f.time, counter > 0 ? -1.0f : -2.0f);
    rec[2] = make_float4(ng.x, ng.y, ng.z, m.radiusMultiplier * get_radius(m.icx, m.icy, zf, n.z));
    rec[3] = k1; rec[4] = k2;
    f.update_id[slot] = counter > 0 ? 1 : 2;
    f.best[slot] = best;
    if (merge) atomicMin(f.winner + best, (unsigned int)slot);     // first primitive in uv order wins (GL_LESS at equal depth)
}

// update.vert:51-115, applied in place to the winners only
__global__ void __launch_bounds__(128) fuse_merge_kernel(ModelArgs m, FuseArgs f, float4* __restrict__ surfels, const unsigned int* __restrict__ count_dev)
{
    pdl_wait();
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= fuse_slots_x(m.cols) * fuse_slots_y(m.rows)) return;
    if (f.update_id[slot] != 1) return;
    const unsigned int best = f.best[slot];
    if (best >= *count_dev || f.winner[best] != (unsigned int)slot) return;
    f.winner[best] = kNoWinner;                                     // re-arm
    float4* s = surfels + 5 * (size_t)best;
    const float4* nw = f.staging + 5 * (size_t)slot;
    const float4 sp = s[0], sc = s[1], sn = s[2], sk1 = s[3], sk2 = s[4];
    const float4 np = nw[0], nc = nw[1], nn = nw[2], nk1 = nw[3], nk2 = nw[4];
    const float c_k = sp.w, a = np.w;
    if (nn.w < (1.0f + 0.5f) * sn.w) {
        const float d = c_k + a;
        s[0] = make_float4(((c_k * sp.x) + (a * np.x)) / d, ((c_k * sp.y) + (a * np.y)) / d, ((c_k * sp.z) + (a * np.z)) / d, d);
        const float3 oc = decode_color(sc.x), ncol = decode_color(nc.x);
        const float3 avg = make_float3(((c_k * oc.x) + (a * ncol.x)) / d, ((c_k * oc.y) + (a * ncol.y)) / d, ((c_k * oc.z) + (a * ncol.z)) / d);
        s[1] = make_float4(encode_color(avg), sc.y, sc.z, (float)f.time);
        const float4 nr = make_float4(((c_k * sn.x) + (a * nn.x)) / d, ((c_k * sn.y) + (a * nn.y)) / d, ((c_k * sn.z) + (a * nn.z)) / d, ((c_k * sn.w) + (a * nn.w)) / d);
        const float len = sqrtf(nr.x * nr.x + nr.y * nr.y + nr.z * nr.z);
        s[2] = make_float4(nr.x / len, nr.y / len, nr.z / len, nr.w);
        s[3] = make_float4(((c_k * sk1.x) + (a * nk1.x)) / d, ((c_k * sk1.y) + (a * nk1.y)) / d, ((c_k * sk1.z) + (a * nk1.z)) / d, ((c_k * sk1.w) + (a * nk1.w)) / d);
        s[4] = make_float4(((c_k * sk2.x) + (a * nk2.x)) / d, ((c_k * sk2.y) + (a * nk2.y)) / d, ((c_k * sk2.z) + (a * nk2.z)) / d, ((c_k * sk2.w) + (a * nk2.w)) / d);
    } else {
        s[0].w = c_k + a;
        s[1].w = (float)f.time;
    }
}

// ------------------------------------------------------------------ clean ---
struct CleanArgs {
    const unsigned int* index; const float4 *vertConf, *colorTime;
    const float* inv_pose;            // device Ri[9], ti[3]
    const float* active_kf; int kf_dim;
    int time;
    const float4* staging; const unsigned char* update_id; int n_slots;
};
// copy_unstable.vert:62-166 on one record; rec[1].w is rewritten (-2 -> time)
__device__ __forceinline__ bool clean_test(const ModelArgs& m, const CleanArgs& c, const float* Pi, float4 (&rec)[5])
{
    const int W = m.cols, H = m.rows;
    bool test = true;
    const float3 lp = pose_apply(Pi, make_float3(rec[0].x, rec[0].y, rec[0].z));
    const float x = ((m.fx * lp.x) / lp.z) + m.cx, y = ((m.fy * lp.y) / lp.z) + m.cy;
    const float3 ln = pose_rotate(Pi, make_float3(rec[2].x, rec[2].y, rec[2].z));
    const float lnz = ln.z / norm(ln);
    int count = 0, zCount = 0;
    const float sub = rec[1].y;
    const int kf = (sub >= 0.0f && sub < (float)c.kf_dim) ? (int)sub : -1;
    const float active = kf >= 0 ? __ldg(c.active_kf + kf) : 0.0f;
    if (lp.z < m.maxDepth && lp.z > 0 && x > 0 && y > 0 && x < (float)W && y < (float)H) {
        auto sample = [&](unsigned int idx, size_t q) {
            if (idx > 0u) {
                const float4 vc = __ldg(c.vertConf + q);
                // both tests need a confident texel strictly behind the surfel: the colour/time texel is only fetched then (the common
                // case -- the texel shows this very surfel, vc.z == lp.z -- costs one gather instead of two)
                if (vc.w > m.confThreshold && vc.z > lp.z) {
                    const float4 ct = __ldg(c.colorTime + q);
                    const float dx = vc.x - lp.x, dy = vc.y - lp.y;
                    if (ct.z < rec[1].z && vc.z - lp.z < 0.01f && sqrtf(dx * dx + dy * dy) < rec[2].w * 1.4f) count++;
                    if (ct.w == (float)c.time && vc.z - lp.z > 0.01f && fabsf(lnz) > 0.85f && active > 0.0f) zCount++;
                }
            }
        };
        // copy_unstable.vert:106-108 walks the window with FLOAT counters in texture space:
        //     for (float i = x / cols - (scale * indexXStep * windowMultiplier); i < x / cols + (...); i += indexXStep),  indexXStep = (1 / (cols * scale)) * 0.5
        // nominally 2 * windowMultiplier samples half a pixel apart, GL_NEAREST; the accumulated counter can stay an ulp below the end
        // value (one more sample) and a sample within round-off of a texel boundary falls on either side, both depending on the
        // texture size.  The loops are run as written, every operation rounded separately like the shader's.
        const float wm = (float)m.cleanWindow;
        const float stepx = __fmul_rn(__fdiv_rn(1.0f, (float)W), 0.5f), stepy = __fmul_rn(__fdiv_rn(1.0f, (float)H), 0.5f);
        const float ux = __fdiv_rn(x, (float)W), uy = __fdiv_rn(y, (float)H);
        const float i0 = __fsub_rn(ux, __fmul_rn(stepx, wm)), i1 = __fadd_rn(ux, __fmul_rn(stepx, wm));
        const float j0 = __fsub_rn(uy, __fmul_rn(stepy, wm)), j1 = __fadd_rn(uy, __fmul_rn(stepy, wm));
        auto texel = [](float u, int n) {      // GL_NEAREST with 8 fractional bits of fixed point
            const float fixed = floorf(__fadd_rn(__fmul_rn(__fmul_rn(u, (float)n), 256.0f), 0.5f));
            return min(max((int)floorf(fixed * 0.00390625f), 0), n - 1);      // / 256: a power of two, the product is the exact quotient
        };
        if (m.cleanWindow == 2) {
            // the reference default: per axis the (4, rarely 5) samples hit at most 3 distinct texels: each DISTINCT texel is fetched once
            // and counted with its multiplicity (the tests only count)
            constexpr int kS = 5;     // 2 * windowMultiplier nominal samples + the one round-off can add
            int tx[kS], ty[kS], nx = 0, ny = 0;
            {
                float i = i0, j = j0;
#pragma unroll
                for (int a = 0; a < kS; ++a) {
                    const bool inx = i < i1, iny = j < j1;
                    tx[a] = inx ? texel(i, W) : -1; ty[a] = iny ? texel(j, H) : -1;
                    nx += inx ? 1 : 0; ny += iny ? 1 : 0;
                    i = __fadd_rn(i, stepx); j = __fadd_rn(j, stepy);
                }
            }
            (void)nx; (void)ny;
            int mx[kS], my[kS];        // multiplicity of sample a if it is the first of a run of equal texels, else 0 (the samples ascend: runs are contiguous)
#pragma unroll
            for (int a = 0; a < kS; ++a) {
                int cx_ = 0, cy_ = 0;
#pragma unroll
                for (int b = 0; b < kS; ++b) if (b >= a) { cx_ += (tx[b] == tx[a]) ? 1 : 0; cy_ += (ty[b] == ty[a]) ? 1 : 0; }
                const bool fx = tx[a] >= 0 && (a == 0 || tx[a] != tx[a > 0 ? a - 1 : 0]), fy = ty[a] >= 0 && (a == 0 || ty[a] != ty[a > 0 ? a - 1 : 0]);
                mx[a] = fx ? cx_ : 0; my[a] = fy ? cy_ : 0;
            }
#pragma unroll
            for (int a = 0; a < kS; ++a) {
                if (mx[a] == 0) continue;
                unsigned int idx[kS];          // one window column: its index loads are issued before the first dependent read
#pragma unroll
                for (int b = 0; b < kS; ++b) idx[b] = (my[b] > 0) ? __ldg(c.index + (size_t)ty[b] * W + tx[a]) : 0u;
#pragma unroll
                for (int b = 0; b < kS; ++b) {
                    const int mult = mx[a] * my[b];
                    if (mult > 0 && idx[b] > 0u) {
                        const int c0 = count, z0 = zCount;
                        sample(idx[b], (size_t)ty[b] * W + tx[a]);
                        count = c0 + (count - c0) * mult; zCount = z0 + (zCount - z0) * mult;
                    }
                }
            }
        } else {
            for (float i = i0; i < i1; i = __fadd_rn(i, stepx))
                for (float j = j0; j < j1; j = __fadd_rn(j, stepy)) {
                    const size_t q = (size_t)texel(j, H) * W + texel(i, W);
                    sample(__ldg(c.index + q), q);
                }
        }
    }
    if (rec[3].w < -m.curvThr || rec[3].w > m.curvThr || rec[4].w < -m.curvThr || rec[4].w > m.curvThr) test = false;
    if (count > 8 || zCount > 4) test = false;
    if (rec[1].w == -2.0f) rec[1].w = (float)c.time;
    if (rec[1].w == -1.0f || (((float)c.time - rec[1].w) > 200.0f && rec[0].w < m.confThreshold)) test = false;
    return test;
}
// items [0, count) = existing surfels, [count, count + n_slots) = this frame's staging slots (update_id 2 = live)
__device__ __forceinline__ bool clean_load(const CleanArgs& c, const float4* __restrict__ surfels, unsigned int count, unsigned int i, float4 (&rec)[5])
{
    const float4* src;
    if (i < count) src = surfels + 5 * (size_t)i;
    else {
        const unsigned int slot = i - count;
        if (c.update_id[slot] != 2) return false;        // 0: nothing emitted; 1: merged (vColor.w == -1 -> dropped)
        src = c.staging + 5 * (size_t)slot;
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) rec[k] = src[k];
    return true;
}
struct ScanTail { unsigned int* ticket; unsigned int* block_offsets; unsigned int capacity; unsigned int* total_out; unsigned int* overflow; };
__global__ void __launch_bounds__(kScanBlock) clean_flags_kernel(ModelArgs m, CleanArgs c, const float4* __restrict__ surfels, const unsigned int* __restrict__ count_dev,
                                                                 unsigned char* __restrict__ flags, unsigned int* __restrict__ block_counts, ScanTail s)
{
    pdl_wait();
    const unsigned int count = *count_dev, n = count + (unsigned int)c.n_slots;
    float Pi[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) Pi[k] = __ldg(c.inv_pose + k);
    for (unsigned int blk = blockIdx.x; blk * kScanBlock < n; blk += gridDim.x) {
        const unsigned int i = blk * kScanBlock + threadIdx.x;
        bool f = false;
        if (i < n) {
            float4 rec[5];
            f = clean_load(c, surfels, count, i, rec) && clean_test(m, c, Pi, rec);
            flags[i] = f ? 1 : 0;
        }
        const int cnt = __syncthreads_count(f);
        if (threadIdx.x == 0) block_counts[blk] = cnt;
    }
    // the block that finishes last scans the counts (no separate single-block launch between the two passes); the ticket re-arms itself
    __shared__ bool s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        s_last = atomicAdd(s.ticket, 1u) == gridDim.x - 1;
        if (s_last) *s.ticket = 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    scan_blocks_cta<kScanBlock>(block_counts, s.block_offsets, n, s.capacity, s.total_out, s.overflow);
}
__global__ void __launch_bounds__(kScanBlock) clean_scatter_kernel(CleanArgs c, const float4* __restrict__ surfels, const unsigned int* __restrict__ count_dev,
                                                                   const unsigned char* __restrict__ flags, const unsigned int* __restrict__ block_offsets,
                                                                   unsigned int capacity, float4* __restrict__ out)
{
    pdl_wait();
    __shared__ unsigned int s_warp[kScanBlock / 32];
    const unsigned int count = *count_dev, n = count + (unsigned int)c.n_slots;
    for (unsigned int blk = blockIdx.x; blk * kScanBlock < n; blk += gridDim.x) {
        const unsigned int i = blk * kScanBlock + threadIdx.x;
        const bool f = i < n && flags[i];
        const unsigned int pos = block_offsets[blk] + block_rank(f, s_warp);
        if (f && pos < capacity) {
            float4 rec[5];
            clean_load(c, surfels, count, i, rec);
            if (rec[1].w == -2.0f) rec[1].w = (float)c.time;
#pragma unroll
            for (int k = 0; k < 5; ++k) out[5 * (size_t)pos + k] = rec[k];
        }
        __syncthreads();       // s_warp is reused by the next tile
    }
}

// ------------------------------------------------------------------ updateModel ---
// GlobalModel::updateModel (GlobalModel.cpp:690-767, update_delta_trans.vert:41-91): rigid correction per sub-map, in place.
// Streams 32 of the 80 bytes of every surfel (position and normal records); delta = n_delta row-major 4x4 matrices.
__global__ void __launch_bounds__(256) update_model_kernel(float4* __restrict__ surfels, const unsigned int* __restrict__ count_dev,
                                                           const float* __restrict__ delta, int n_delta)
{
    const unsigned int count = *count_dev;
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += blockDim.x * gridDim.x) {
        float4* s = surfels + 5 * (size_t)i;
        const unsigned int sub = (unsigned int)reinterpret_cast<const float*>(s + 1)[1];
        if (sub >= (unsigned int)n_delta) continue;
        const float* T = delta + 16 * (size_t)sub;
        const float4 p = s[0], n = s[2];
        s[0] = make_float4(((T[0] * p.x + T[1] * p.y) + T[2] * p.z) + T[3], ((T[4] * p.x + T[5] * p.y) + T[6] * p.z) + T[7],
                           ((T[8] * p.x + T[9] * p.y) + T[10] * p.z) + T[11], p.w);
        s[2] = make_float4((T[0] * n.x + T[1] * n.y) + T[2] * n.z, (T[4] * n.x + T[5] * n.y) + T[6] * n.z, (T[8] * n.x + T[9] * n.y) + T[10] * n.z, n.w);
    }
}

// ------------------------------------------------------------------ savePly ---
// HRBFFusion::savePly (HRBFFusion.cpp:1737-1853): every surfel with confidence > threshold becomes one 43-byte binary PLY vertex
// {x y z | r g b | -nx -ny -nz | curvature_max curvature_min | radius | submapIndex}, in map order.  The reference downloads the whole
// map (80 B per surfel) and filters on the host; here the filter + packing is an order-preserving compaction on the device and only
// the packed records cross PCIe.
constexpr int kPlyVertexBytes = 43;
__global__ void __launch_bounds__(kScanBlock) ply_flags_kernel(const float4* __restrict__ surfels, const unsigned int* __restrict__ count_dev, float confThreshold,
                                                               unsigned char* __restrict__ flags, unsigned int* __restrict__ block_counts)
{
    const unsigned int n = *count_dev;
    for (unsigned int blk = blockIdx.x; blk * kScanBlock < n; blk += gridDim.x) {
        const unsigned int i = blk * kScanBlock + threadIdx.x;
        bool f = false;
        if (i < n) {
            f = __ldg(&surfels[5 * (size_t)i].w) > confThreshold;
            flags[i] = f ? 1 : 0;
        }
        const int cnt = __syncthreads_count(f);
        if (threadIdx.x == 0) block_counts[blk] = cnt;
    }
}
__global__ void __launch_bounds__(kScanBlock) ply_scatter_kernel(const float4* __restrict__ surfels, const unsigned int* __restrict__ count_dev,
                                                                 const unsigned char* __restrict__ flags, const unsigned int* __restrict__ block_offsets,
                                                                 unsigned char* __restrict__ out)
{
    __shared__ unsigned int s_warp[kScanBlock / 32];
    __shared__ unsigned int s_n;
    __shared__ __align__(4) unsigned char s_rec[kScanBlock * kPlyVertexBytes + 1];
    const unsigned int n = *count_dev;
    for (unsigned int blk = blockIdx.x; blk * kScanBlock < n; blk += gridDim.x) {
        const unsigned int i = blk * kScanBlock + threadIdx.x;
        const bool f = i < n && flags[i];
        const unsigned int r = block_rank(f, s_warp);
        if (threadIdx.x == kScanBlock - 1) s_n = r + (f ? 1u : 0u);
        if (f) {
            const float4* sp = surfels + 5 * (size_t)i;
            const float4 pos = __ldg(sp), col = __ldg(sp + 1), nor = __ldg(sp + 2);
            const float cmax = __ldg(&sp[3].w), cmin = __ldg(&sp[4].w);
            const int c = (int)col.x;
            const float fl[10] = { pos.x, pos.y, pos.z, -nor.x, -nor.y, -nor.z, cmax, cmin, nor.w, col.y };
            unsigned char* d = s_rec + r * kPlyVertexBytes;
            auto put = [&](int off, float v) { const unsigned int b = __float_as_uint(v); d[off] = b & 0xff; d[off + 1] = (b >> 8) & 0xff; d[off + 2] = (b >> 16) & 0xff; d[off + 3] = b >> 24; };
            put(0, fl[0]); put(4, fl[1]); put(8, fl[2]);
            d[12] = (unsigned char)(c >> 16 & 0xFF); d[13] = (unsigned char)(c >> 8 & 0xFF); d[14] = (unsigned char)(c & 0xFF);
#pragma unroll
            for (int k = 3; k < 10; ++k) put(15 + 4 * (k - 3), fl[k]);
        }
        __syncthreads();
        // the block's records are one contiguous byte range of the output: coalesced byte stores
        const size_t base = (size_t)block_offsets[blk] * kPlyVertexBytes;
        const unsigned int bytes = s_n * kPlyVertexBytes;
        for (unsigned int b = threadIdx.x; b < bytes; b += kScanBlock) out[base + b] = s_rec[b];
        __syncthreads();       // s_warp / s_rec are reused by the next tile
    }
}

__global__ void fill_u32_kernel(unsigned int* p, size_t n, unsigned int v)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace hrbf
