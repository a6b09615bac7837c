// icp_tile.cuh -- the ICP JtJ / Jtr reduction (Core/src/Cuda/reduce.cu:317-573) over TMA-staged tiles.
//
// The image is cut into tiles (about 80 x 13 pixels, one or a few per CTA).  A tile's packed current-frame records are
// a 2-D box of the image: they are brought into shared memory by the TMA unit (cp.async.bulk.tensor.2d through a tensor map
// per array and pyramid level, one request per box, completion on an mbarrier) without passing through registers; parts of a
// box that fall outside the image are zero-filled by the unit and never read.  Under the inter-frame motions the tracker sees (<= 1 cm,
// <= 0.5 degrees) the projective association is a near-constant shift over a tile, so the model records a tile needs lie in a
// window only a few pixels larger than the tile: its bounding box is computed from the staged current-frame records, and the
// window (model pk0 / pk1 / icp-weight rows) is staged by a second round of bulk copies.  The dependent gather of the
// reference (two dependent global round trips PER PIXEL, reduce.cu:330-381) becomes two bulk stages PER TILE with the whole
// tile in flight, and the per-pixel gather reads shared memory.  An association that leaves the window (depth edges, large
// motion) is served by the same __ldg path as before: same arithmetic, same result.
#pragma once
#include <cuda.h>      // CUtensorMap (type only: the encoder is fetched through the runtime, nothing links against libcuda)
#include "odometry_kernels.cuh"

namespace hrbf {

struct IcpTileGeom {
    int ncol, nrow;      // tile grid over the image
    int tw, th;          // largest tile (pixels); tw is a multiple of 4
    int mw, mh;          // model window (pixels); mw is a multiple of 4
    int cnb, cbx;        // a tile row is fetched as cnb boxes of cbx pixels (a TMA box is at most 256 elements = 128 float4 wide) ...
    int mnb, mbx;        // ... a window row as mnb boxes of mbx pixels; shared-memory layout [box][row][pixel in box]
    int cbs, mbs, wbs;   // distance between consecutive boxes in PIXELS of the array (every box starts 128-byte aligned): records of the
                         // tile, records of the window, weights of the window
    int ctas, threads;   // launch shape
};
inline void icp_tile_boxes(IcpTileGeom& g)      // tw, mw <= 256
{
    g.cnb = div_up(g.tw, 128); g.cbx = (div_up(g.tw, g.cnb) + 3) & ~3;
    g.mnb = div_up(g.mw, 128); g.mbx = (div_up(g.mw, g.mnb) + 3) & ~3;
    g.mw = g.mnb * g.mbx;
    g.cbs = (g.cbx * g.th + 7) & ~7;        // x 16 B = multiple of 128 B
    g.mbs = (g.mbx * g.mh + 7) & ~7;
    g.wbs = (g.mbx * g.mh + 31) & ~31;      // x 4 B
}
constexpr int kTileHaloX = 4, kTileHaloY = 4;      // window = tile + 2 x halo (+ 4 pixels of alignment slack in x)

// tiles ~ (cols / 8a) x (rows / 37b) so that 296 = 8 x 37 CTAs (2 per SM) get 1 (640x480), 4 (1280x960) ... tiles each
inline IcpTileGeom icp_tile_geom(int rows, int cols, int num_sms)
{
    IcpTileGeom g;
    const int per_row = 8, per_col = num_sms * 2 / per_row;      // 8 x 37 on a 148-SM part
    int a = 1, b = 1;
    while (div_up(cols, per_row * a) > 80) ++a;
    while (div_up(rows, per_col * b) > 13) ++b;
    g.tw = (div_up(cols, per_row * a) + 3) & ~3;
    if (g.tw < 8) g.tw = 8;
    g.ncol = div_up(cols, g.tw);
    g.nrow = per_col * b;
    if (g.nrow > rows) g.nrow = rows;
    g.th = div_up(rows, g.nrow);
    g.mw = g.tw + 2 * kTileHaloX + 4;
    g.mh = g.th + 2 * kTileHaloY;
    icp_tile_boxes(g);
    g.threads = kReduceThreads;
    const int tiles = g.ncol * g.nrow;
    g.ctas = tiles < num_sms * 2 ? tiles : num_sms * 2;
    if (g.ctas > kMaxReduceBlocks) g.ctas = kMaxReduceBlocks;
    return g;
}
// every array's boxes start 128-byte aligned
__host__ __device__ inline size_t icp_tile_curr_bytes(const IcpTileGeom& g) { return (size_t)g.cnb * g.cbs * sizeof(float4); }
__host__ __device__ inline size_t icp_tile_model_bytes(const IcpTileGeom& g) { return (size_t)g.mnb * g.mbs * sizeof(float4); }
__host__ __device__ inline size_t icp_tile_weight_bytes(const IcpTileGeom& g) { return (size_t)g.mnb * g.wbs * sizeof(float); }
inline size_t icp_tile_smem_bytes(const IcpTileGeom& g)
{
    return 2 * icp_tile_curr_bytes(g) + 2 * icp_tile_model_bytes(g) + icp_tile_weight_bytes(g) + 128;
}
// the five tensor maps of one pyramid level and geometry, passed BY VALUE inside __grid_constant__ kernel parameters (the TMA unit
// reads the descriptor from parameter space): packed records of the current frame (pk0, pk1) and of the model (pk0, pk1) as
// [rows][cols] arrays of 16-byte pixels (encoded as 2 x 64-bit elements), and the model's icp-weight map as [rows][cols] floats
struct alignas(64) IcpTileMaps { CUtensorMap pc0, pc1, pg0, pg1, w; };

// ---- mbarrier / bulk-copy primitives (PTX ISA: mbarrier, cp.async.bulk) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy by the TMA unit; bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)), "l"(src_gmem),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// one box of a 2-D tensor (tensor map in global memory) -> shared memory; (cx, cy) = element coordinates of the box's low corner,
// may lie outside the tensor (zero fill); dst 128-byte aligned
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const void* tmap, int cx, int cy, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst_smem)),
                 "l"(tmap), "r"(smem_u32(bar)), "r"(cx), "r"(cy)
                 : "memory");
}

// A staged tile: where its pieces sit in shared memory and which part of the image they hold
struct IcpTileView {
    const float4 *c0, *c1;       // current-frame records, [cnb][th][cbx]
    const float4 *g0, *g1;       // model window, [mnb][mh][mbx]
    const float* gw;             // icp-weight window, same layout
    int x0, y0, w, h;            // tile rectangle in the image
    int mx0, my0, mwa, mha;      // model window rectangle staged (mwa == 0: nothing staged); may stick out of the image (zero fill)
    int cbx, cbs, mbx, mbs, wbs; // box widths and box-to-box distances of the layouts above
    // at most two boxes per row (tiles and windows are at most 256 pixels wide): no integer division
    __device__ __forceinline__ int cidx(int x, int y) const { return x >= cbx ? cbs + y * cbx + (x - cbx) : y * cbx + x; }
    __device__ __forceinline__ int midx(int x, int y) const { return x >= mbx ? mbs + y * mbx + (x - mbx) : y * mbx + x; }
    __device__ __forceinline__ int widx(int x, int y) const { return x >= mbx ? wbs + y * mbx + (x - mbx) : y * mbx + x; }
};

// projection half of icp_gather_model (odometry_kernels.cuh): everything but the loads
__device__ __forceinline__ void icp_project(const IcpArgs& a, const IcpCurr& c, const float* Rc, const float* tc, const float* Rpi, const float* tp, IcpModel& m)
{
    m.vg = mul(Rc, make_float3(c.vx, c.vy, c.vz)) + make_float3(tc[0], tc[1], tc[2]);
    const float3 vcp = mul(Rpi, m.vg - make_float3(tp[0], tp[1], tp[2]));
    m.ux = __float2int_rn(vcp.x * a.fx / vcp.z + a.cx);
    m.uy = __float2int_rn(vcp.y * a.fy / vcp.z + a.cy);
    m.ok = !(m.ux < 0 || m.uy < 0 || m.ux >= a.cols || m.uy >= a.rows || vcp.z < 0) && !(isnan(c.vx) || isnan(c.nx) || isnan(c.k1) || isnan(c.k2));
    m.ng = mul(Rc, make_float3(c.nx, c.ny, c.nz));
    m.vx = m.vy = m.vz = m.nx = m.ny = m.nz = m.k1 = m.k2 = m.w = 0.f;
}
__device__ __forceinline__ IcpCurr icp_curr_from(const float4 p0, const float4 p1)
{
    IcpCurr c;
    c.vx = p0.x; c.vy = p0.y; c.vz = p0.z; c.nx = p0.w; c.ny = p1.x; c.nz = p1.y; c.k1 = p1.z; c.k2 = p1.w;
    return c;
}
__device__ __forceinline__ void icp_model_from(IcpModel& m, const float4 p0, const float4 p1, float w)
{
    m.vx = p0.x; m.vy = p0.y; m.vz = p0.z; m.nx = p0.w; m.ny = p1.x; m.nz = p1.y; m.k1 = p1.z; m.k2 = p1.w; m.w = w;
}

// stage 1: the tile's current-frame records, one TMA request per array and box (issued by one thread, which also arms the barrier)
__device__ __forceinline__ void icp_tile_issue_curr(const IcpTileMaps& m, const IcpTileGeom& g, float4* s_c0, float4* s_c1, int x0, int y0, uint64_t* bar)
{
    mbar_expect_tx(bar, 2u * (uint32_t)(g.cnb * g.cbx * g.th) * (uint32_t)sizeof(float4));
    for (int b = 0; b < g.cnb; ++b) {
        tma_load_2d(s_c0 + b * g.cbs, &m.pc0, 2 * (x0 + b * g.cbx), y0, bar);      // x in 64-bit elements: 2 per pixel
        tma_load_2d(s_c1 + b * g.cbs, &m.pc1, 2 * (x0 + b * g.cbx), y0, bar);
    }
}
// stage 2: the model window
__device__ __forceinline__ void icp_tile_issue_model(const IcpTileMaps& m, const IcpTileGeom& g, bool use_weight, float4* s_g0, float4* s_g1, float* s_gw,
                                                     int mx0, int my0, uint64_t* bar)
{
    mbar_expect_tx(bar, (uint32_t)(g.mnb * g.mbx * g.mh) * (use_weight ? 36u : 32u));
    for (int b = 0; b < g.mnb; ++b) {
        tma_load_2d(s_g0 + b * g.mbs, &m.pg0, 2 * (mx0 + b * g.mbx), my0, bar);
        tma_load_2d(s_g1 + b * g.mbs, &m.pg1, 2 * (mx0 + b * g.mbx), my0, bar);
        if (use_weight) tma_load_2d(s_gw + b * g.wbs, &m.w, mx0 + b * g.mbx, my0, bar);
    }
}

// bounding box of the associations of a staged tile -> s_box[4] = {min ux, min uy, max ux, max uy} (shared, pre-set to +-big)
template <int kThreads>
__device__ __forceinline__ void icp_tile_bbox(const IcpArgs& a, const IcpTileView& t, const float* Rc, const float* tc, const float* Rpi, const float* tp, int* s_box)
{
    int lo_x = 1 << 30, lo_y = 1 << 30, hi_x = -1, hi_y = -1;
    const int w = t.w, n = w * t.h;
    const int dy = kThreads / w, dx = kThreads - dy * w;      // pixel i + kThreads from pixel i without a division per pixel
    int y = (int)threadIdx.x / w, x = (int)threadIdx.x - y * w;
    for (int i = threadIdx.x; i < n; i += kThreads, x += dx, y += dy) {
        if (x >= w) { x -= w; ++y; }
        const int ci = t.cidx(x, y);
        const IcpCurr c = icp_curr_from(t.c0[ci], t.c1[ci]);
        IcpModel m;
        icp_project(a, c, Rc, tc, Rpi, tp, m);
        if (m.ok) { lo_x = min(lo_x, m.ux); lo_y = min(lo_y, m.uy); hi_x = max(hi_x, m.ux); hi_y = max(hi_y, m.uy); }
    }
    lo_x = __reduce_min_sync(0xffffffffu, lo_x); lo_y = __reduce_min_sync(0xffffffffu, lo_y);
    hi_x = __reduce_max_sync(0xffffffffu, hi_x); hi_y = __reduce_max_sync(0xffffffffu, hi_y);
    if ((threadIdx.x & 31) == 0 && hi_x >= 0) { atomicMin(&s_box[0], lo_x); atomicMin(&s_box[1], lo_y); atomicMax(&s_box[2], hi_x); atomicMax(&s_box[3], hi_y); }
}
// the window to stage for a bounding box: anchored `margin` pixels (or, when the box is small, half of the slack) below its low
// corner; it may stick out of the image, the TMA unit zero-fills what is not there
__device__ __forceinline__ void icp_tile_window(const int* s_box, int mw, int mh, int margin, int& mx0, int& my0, bool& any)
{
    any = s_box[2] >= 0;
    const int bw = s_box[2] - s_box[0] + 1, bh = s_box[3] - s_box[1] + 1;
    // x is aligned down to 4 pixels: the TMA unit wants a box to start on a 16-byte boundary in global memory, and the icp-weight
    // map has 4-byte pixels (an unaligned start traps as "illegal instruction")
    mx0 = (s_box[0] - min(margin, max(0, (mw - bw) / 2))) & ~3;
    my0 = s_box[1] - min(margin, max(0, (mh - bh) / 2));
}

// stage 3: the pass over a staged tile: same arithmetic as icp_pass_nosearch_t (odometry_kernels.cuh), gathers from shared memory.
// Two pixels per thread are worked on together (pixel i and i + kThreads): the pass is bound by instruction issue and dependent
// fp32 chains, not by memory, and two independent chains per thread keep the schedulers busy.
__device__ __forceinline__ void icp_tile_fetch(const IcpArgs& a, const IcpTileView& t, IcpModel& m)
{
    if (!m.ok) return;
    const int lx = m.ux - t.mx0, ly = m.uy - t.my0;
    if ((unsigned)lx < (unsigned)t.mwa && (unsigned)ly < (unsigned)t.mha) {
        const int q = t.midx(lx, ly);
        icp_model_from(m, t.g0[q], t.g1[q], a.use_weight ? t.gw[t.widx(lx, ly)] : 1.f);
    } else {
        const int q = m.uy * a.cols + m.ux;
        icp_model_from(m, __ldg(a.pg0 + q), __ldg(a.pg1 + q), a.use_weight ? __ldg(a.w + q) : 1.f);
    }
}
template <int kThreads>
__device__ __forceinline__ void icp_tile_pass(const IcpArgs& a, const IcpTileView& t, const float* Rc, const float* tc, const float* Rpi, const float* tp, float (&acc)[32])
{
    const int w = t.w, n = w * t.h;
    const int dy = kThreads / w, dx = kThreads - dy * w;      // pixel i + kThreads from pixel i without a division per pixel
    int y0 = (int)threadIdx.x / w, x0 = (int)threadIdx.x - y0 * w;
    for (int i = threadIdx.x; i < n; i += 2 * kThreads) {
        int x1 = x0 + dx, y1 = y0 + dy;
        if (x1 >= w) { x1 -= w; ++y1; }
        const bool two = i + kThreads < n;
        const int c0i = t.cidx(x0, y0), c1i = two ? t.cidx(x1, y1) : c0i;
        const IcpCurr ca = icp_curr_from(t.c0[c0i], t.c1[c0i]), cb = icp_curr_from(t.c0[c1i], t.c1[c1i]);
        IcpModel ma, mb;
        icp_project(a, ca, Rc, tc, Rpi, tp, ma);
        icp_project(a, cb, Rc, tc, Rpi, tp, mb);
        mb.ok = mb.ok && two;
        icp_tile_fetch(a, t, ma);
        icp_tile_fetch(a, t, mb);
        icp_finish(a, ma, Rpi, tp, (t.y0 + y0) * a.cols + t.x0 + x0, acc);
        if (two) icp_finish(a, mb, Rpi, tp, (t.y0 + y1) * a.cols + t.x0 + x1, acc);
        x0 = x1 + dx; y0 = y1 + dy;
        if (x0 >= w) { x0 -= w; ++y0; }
    }
}

// mode as icp_reduce_kernel: 0 = store the 29 sums in st->icp_sums, 1 = store and run the Gauss-Newton update in the last block
__global__ void __launch_bounds__(kReduceThreads, 2) icp_tile_reduce_kernel(IcpArgs a, IcpTileGeom g, const __grid_constant__ IcpTileMaps maps, ReduceWork* wk, int mode, int cur_level, int next_level)
{
    extern __shared__ __align__(128) unsigned char s_dyn[];
    float4* s_c0 = reinterpret_cast<float4*>(s_dyn);
    float4* s_c1 = reinterpret_cast<float4*>(s_dyn + icp_tile_curr_bytes(g));
    float4* s_g0 = reinterpret_cast<float4*>(s_dyn + 2 * icp_tile_curr_bytes(g));
    float4* s_g1 = reinterpret_cast<float4*>(s_dyn + 2 * icp_tile_curr_bytes(g) + icp_tile_model_bytes(g));
    float* s_gw = reinterpret_cast<float*>(s_dyn + 2 * icp_tile_curr_bytes(g) + 2 * icp_tile_model_bytes(g));
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ double s_total[32];
    __shared__ float s_pose[24];
    __shared__ int s_box[4];
    __shared__ int s_win[3];

    pdl_launch_dependents();
    if (threadIdx.x == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); mbar_fence_init(); }
    pdl_wait();      // the maps and the pose may come from the previous kernel of the stream
    TrackState* st = &wk->st;
    if (threadIdx.x < 9) { s_pose[threadIdx.x] = st->Rcurr[threadIdx.x]; s_pose[12 + threadIdx.x] = st->Rprev_inv[threadIdx.x]; }
    if (threadIdx.x < 3) { s_pose[9 + threadIdx.x] = st->tcurr[threadIdx.x]; s_pose[21 + threadIdx.x] = st->tprev[threadIdx.x]; }
    const bool level_done = (st->done_level == cur_level);
    const int tiles = g.ncol * g.nrow;
    // the first tile's current-frame records are requested before anything else
    if (threadIdx.x == 0) { s_box[0] = s_box[1] = 1 << 30; s_box[2] = s_box[3] = -1; }
    __syncthreads();
    if (!level_done && (int)blockIdx.x < tiles && threadIdx.x == 0) {
        const int tcx = (int)blockIdx.x % g.ncol, try_ = (int)blockIdx.x / g.ncol;
        icp_tile_issue_curr(maps, g, s_c0, s_c1, tcx * g.tw, (int)(((long long)a.rows * try_) / g.nrow), &s_bar[0]);
    }
    float Rc[9], tc[3], Rpi[9], tp[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) { Rc[k] = s_pose[k]; Rpi[k] = s_pose[12 + k]; }
#pragma unroll
    for (int k = 0; k < 3; ++k) { tc[k] = s_pose[9 + k]; tp[k] = s_pose[21 + k]; }

    float acc[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) acc[k] = 0.f;
    uint32_t par0 = 0, par1 = 0;
    if (!level_done) {
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            IcpTileView t;
            const int tcx = tile % g.ncol, try_ = tile / g.ncol;
            t.x0 = tcx * g.tw; t.w = min(g.tw, a.cols - t.x0);
            t.y0 = (int)(((long long)a.rows * try_) / g.nrow); t.h = (int)(((long long)a.rows * (try_ + 1)) / g.nrow) - t.y0;
            t.c0 = s_c0; t.c1 = s_c1; t.g0 = s_g0; t.g1 = s_g1; t.gw = s_gw; t.cbx = g.cbx; t.cbs = g.cbs; t.mbx = g.mbx; t.mbs = g.mbs; t.wbs = g.wbs;
            mbar_wait(&s_bar[0], par0); par0 ^= 1u;
            icp_tile_bbox<kReduceThreads>(a, t, Rc, tc, Rpi, tp, s_box);
            __syncthreads();
            if (threadIdx.x == 0) {
                int mx0, my0; bool any;
                icp_tile_window(s_box, g.mw, g.mh, kTileHaloX, mx0, my0, any);
                s_win[0] = mx0; s_win[1] = my0; s_win[2] = any ? 1 : 0;
                if (any) icp_tile_issue_model(maps, g, a.use_weight != 0, s_g0, s_g1, s_gw, mx0, my0, &s_bar[1]);
                s_box[0] = s_box[1] = 1 << 30; s_box[2] = s_box[3] = -1;      // for the next tile
            }
            __syncthreads();
            t.mx0 = s_win[0]; t.my0 = s_win[1];
            if (s_win[2]) { t.mwa = g.mw; t.mha = g.mh; mbar_wait(&s_bar[1], par1); par1 ^= 1u; }
            else t.mwa = t.mha = 0;
            icp_tile_pass<kReduceThreads>(a, t, Rc, tc, Rpi, tp, acc);
            __syncthreads();      // the tile's buffers are free: request the next tile's records
            const int next = tile + gridDim.x;
            if (next < tiles && threadIdx.x == 0) {
                const int ncx = next % g.ncol, nry = next / g.ncol;
                icp_tile_issue_curr(maps, g, s_c0, s_c1, ncx * g.tw, (int)(((long long)a.rows * nry) / g.nrow), &s_bar[0]);
            }
        }
    }
    if (grid_reduce32(acc, wk->partials, &st->ticket, s_total)) {
        if (threadIdx.x < 32) st->icp_sums[threadIdx.x] = s_total[threadIdx.x];
        __syncthreads();
        if (threadIdx.x == 0 && mode == 1 && !level_done) gn_update(st, cur_level, next_level);
    }
}

}  // namespace hrbf
