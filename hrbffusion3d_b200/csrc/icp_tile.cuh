// icp_tile.cuh -- the ICP JtJ / Jtr reduction (Core/src/Cuda/reduce.cu:317-573) over TMA-staged tiles.
//
// The image is cut into tiles (about 80 x 13 pixels, one or a few per CTA).  A tile's packed current-frame records are
// contiguous rows: they are brought into shared memory by the TMA unit (cp.async.bulk, one bulk copy per row and array,
// completion on an mbarrier) without passing through registers.  Under the inter-frame motions the tracker sees (<= 1 cm,
// <= 0.5 degrees) the projective association is a near-constant shift over a tile, so the model records a tile needs lie in a
// window only a few pixels larger than the tile: its bounding box is computed from the staged current-frame records, and the
// window (model pk0 / pk1 / icp-weight rows) is staged by a second round of bulk copies.  The dependent gather of the
// reference (two dependent global round trips PER PIXEL, reduce.cu:330-381) becomes two bulk stages PER TILE with the whole
// tile in flight, and the per-pixel gather reads shared memory.  An association that leaves the window (depth edges, large
// motion) is served by the same __ldg path as before: same arithmetic, same result.
#pragma once
#include "odometry_kernels.cuh"

namespace hrbf {

struct IcpTileGeom {
    int ncol, nrow;      // tile grid over the image
    int tw, th;          // largest tile (pixels); tw is a multiple of 4
    int mw, mh;          // model window capacity (pixels); mw is a multiple of 4
    int ctas, threads;   // launch shape
};
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
    g.threads = kReduceThreads;
    const int tiles = g.ncol * g.nrow;
    g.ctas = tiles < num_sms * 2 ? tiles : num_sms * 2;
    if (g.ctas > kMaxReduceBlocks) g.ctas = kMaxReduceBlocks;
    return g;
}
inline size_t icp_tile_smem_bytes(const IcpTileGeom& g)
{
    return (size_t)g.tw * g.th * 2 * sizeof(float4) + (size_t)g.mw * g.mh * (2 * sizeof(float4) + sizeof(float)) + 16;
}

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

// A staged tile: where its pieces sit in shared memory and which part of the image they hold
struct IcpTileView {
    const float4 *c0, *c1;       // current-frame records, [th][tw] (row pitch tw)
    const float4 *g0, *g1;       // model window, [mh][mw] (row pitch mw)
    const float* gw;             // icp-weight window
    int x0, y0, w, h;            // tile rectangle in the image
    int mx0, my0, mwa, mha;      // model window rectangle actually staged (mwa == 0: nothing staged)
    int tw, mw;                  // row pitches
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

// stage 1: the tile's current-frame rows (issued by warp 0; one thread arms the barrier with the byte count)
__device__ __forceinline__ void icp_tile_issue_curr(const IcpArgs& a, float4* s_c0, float4* s_c1, int tw, int x0, int y0, int w, int h, uint64_t* bar)
{
    const int lane = threadIdx.x & 31;
    const uint32_t row_bytes = (uint32_t)w * sizeof(float4);
    if (lane == 0) mbar_expect_tx(bar, 2u * row_bytes * (uint32_t)h);
    __syncwarp();
    for (int r = lane; r < 2 * h; r += 32) {
        const int y = r >> 1;
        const size_t src = (size_t)(y0 + y) * a.cols + x0;
        if (r & 1) bulk_g2s(s_c1 + y * tw, a.pc1 + src, row_bytes, bar);
        else bulk_g2s(s_c0 + y * tw, a.pc0 + src, row_bytes, bar);
    }
}
// stage 2: the model window rows
__device__ __forceinline__ void icp_tile_issue_model(const IcpArgs& a, float4* s_g0, float4* s_g1, float* s_gw, int mw, int mx0, int my0, int mwa, int mha, uint64_t* bar)
{
    const int lane = threadIdx.x & 31;
    const int per_row = a.use_weight ? 3 : 2;
    if (lane == 0) mbar_expect_tx(bar, (uint32_t)mha * (uint32_t)mwa * (a.use_weight ? 36u : 32u));
    __syncwarp();
    for (int r = lane; r < per_row * mha; r += 32) {
        const int y = r / per_row, k = r - y * per_row;
        const size_t src = (size_t)(my0 + y) * a.cols + mx0;
        if (k == 0) bulk_g2s(s_g0 + y * mw, a.pg0 + src, (uint32_t)mwa * 16u, bar);
        else if (k == 1) bulk_g2s(s_g1 + y * mw, a.pg1 + src, (uint32_t)mwa * 16u, bar);
        else bulk_g2s(s_gw + y * mw, a.w + src, (uint32_t)mwa * 4u, bar);
    }
}

// bounding box of the associations of a staged tile -> s_box[4] = {min ux, min uy, max ux, max uy} (shared, pre-set to +-big)
template <int kThreads>
__device__ __forceinline__ void icp_tile_bbox(const IcpArgs& a, const IcpTileView& t, const float* Rc, const float* tc, const float* Rpi, const float* tp, int* s_box)
{
    int lo_x = 1 << 30, lo_y = 1 << 30, hi_x = -1, hi_y = -1;
    const int n = t.w * t.h;
    for (int i = threadIdx.x; i < n; i += kThreads) {
        const int y = i / t.w, x = i - y * t.w;
        const IcpCurr c = icp_curr_from(t.c0[y * t.tw + x], t.c1[y * t.tw + x]);
        IcpModel m;
        icp_project(a, c, Rc, tc, Rpi, tp, m);
        if (m.ok) { lo_x = min(lo_x, m.ux); lo_y = min(lo_y, m.uy); hi_x = max(hi_x, m.ux); hi_y = max(hi_y, m.uy); }
    }
    lo_x = __reduce_min_sync(0xffffffffu, lo_x); lo_y = __reduce_min_sync(0xffffffffu, lo_y);
    hi_x = __reduce_max_sync(0xffffffffu, hi_x); hi_y = __reduce_max_sync(0xffffffffu, hi_y);
    if ((threadIdx.x & 31) == 0 && hi_x >= 0) { atomicMin(&s_box[0], lo_x); atomicMin(&s_box[1], lo_y); atomicMax(&s_box[2], hi_x); atomicMax(&s_box[3], hi_y); }
}
// the window to stage for a bounding box: anchored at its low corner (x aligned down to 4 pixels), clipped to the image and to the
// capacity; `margin` pixels are left free on the low side for the pose to move during the iterations that reuse the window
__device__ __forceinline__ void icp_tile_window(const int* s_box, int rows, int cols, int mw, int mh, int margin, int& mx0, int& my0, int& mwa, int& mha)
{
    mwa = mha = 0; mx0 = my0 = 0;
    if (s_box[2] < 0) return;
    const int bw = s_box[2] - s_box[0] + 1, bh = s_box[3] - s_box[1] + 1;
    int sx = min(margin, max(0, (mw - 4 - bw) / 2)), sy = min(margin, max(0, (mh - bh) / 2));      // slack split evenly when the box is small
    mx0 = max(0, (s_box[0] - sx) & ~3);
    my0 = max(0, s_box[1] - sy);
    mwa = min(mw, cols - mx0);
    mha = min(mh, rows - my0);
    mwa &= ~3;
}

// stage 3: the pass over a staged tile: same arithmetic as icp_pass_nosearch_t (odometry_kernels.cuh), gathers from shared memory
template <int kThreads>
__device__ __forceinline__ void icp_tile_pass(const IcpArgs& a, const IcpTileView& t, const float* Rc, const float* tc, const float* Rpi, const float* tp, float (&acc)[32])
{
    const int n = t.w * t.h;
    for (int i = threadIdx.x; i < n; i += kThreads) {
        const int y = i / t.w, x = i - y * t.w;
        const IcpCurr c = icp_curr_from(t.c0[y * t.tw + x], t.c1[y * t.tw + x]);
        IcpModel m;
        icp_project(a, c, Rc, tc, Rpi, tp, m);
        if (m.ok) {
            const int lx = m.ux - t.mx0, ly = m.uy - t.my0;
            if ((unsigned)lx < (unsigned)t.mwa && (unsigned)ly < (unsigned)t.mha) {
                const int q = ly * t.mw + lx;
                icp_model_from(m, t.g0[q], t.g1[q], a.use_weight ? t.gw[q] : 1.f);
            } else {
                const int q = m.uy * a.cols + m.ux;
                icp_model_from(m, __ldg(a.pg0 + q), __ldg(a.pg1 + q), a.use_weight ? __ldg(a.w + q) : 1.f);
            }
        }
        icp_finish(a, m, Rpi, tp, (t.y0 + y) * a.cols + t.x0 + x, acc);
    }
}

// mode as icp_reduce_kernel: 0 = store the 29 sums in st->icp_sums, 1 = store and run the Gauss-Newton update in the last block
__global__ void __launch_bounds__(kReduceThreads, 2) icp_tile_reduce_kernel(IcpArgs a, IcpTileGeom g, ReduceWork* wk, int mode, int cur_level, int next_level)
{
    extern __shared__ __align__(128) unsigned char s_dyn[];
    float4* s_c0 = reinterpret_cast<float4*>(s_dyn);
    float4* s_c1 = s_c0 + g.tw * g.th;
    float4* s_g0 = s_c1 + g.tw * g.th;
    float4* s_g1 = s_g0 + g.mw * g.mh;
    float* s_gw = reinterpret_cast<float*>(s_g1 + g.mw * g.mh);
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ double s_total[32];
    __shared__ float s_pose[24];
    __shared__ int s_box[4];
    __shared__ int s_win[4];

    if (threadIdx.x == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); mbar_fence_init(); }
    pdl_wait();      // the maps and the pose may come from the previous kernel of the stream
    TrackState* st = &wk->st;
    if (threadIdx.x < 9) { s_pose[threadIdx.x] = st->Rcurr[threadIdx.x]; s_pose[12 + threadIdx.x] = st->Rprev_inv[threadIdx.x]; }
    if (threadIdx.x < 3) { s_pose[9 + threadIdx.x] = st->tcurr[threadIdx.x]; s_pose[21 + threadIdx.x] = st->tprev[threadIdx.x]; }
    const bool level_done = (st->done_level == cur_level);
    __syncthreads();
    float Rc[9], tc[3], Rpi[9], tp[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) { Rc[k] = s_pose[k]; Rpi[k] = s_pose[12 + k]; }
#pragma unroll
    for (int k = 0; k < 3; ++k) { tc[k] = s_pose[9 + k]; tp[k] = s_pose[21 + k]; }

    float acc[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) acc[k] = 0.f;
    const int tiles = g.ncol * g.nrow;
    uint32_t parity = 0;
    if (!level_done) {
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            IcpTileView t;
            const int tcx = tile % g.ncol, try_ = tile / g.ncol;
            t.x0 = tcx * g.tw; t.w = min(g.tw, a.cols - t.x0);
            t.y0 = (int)(((long long)a.rows * try_) / g.nrow); t.h = (int)(((long long)a.rows * (try_ + 1)) / g.nrow) - t.y0;
            t.c0 = s_c0; t.c1 = s_c1; t.g0 = s_g0; t.g1 = s_g1; t.gw = s_gw; t.tw = g.tw; t.mw = g.mw;
            if (threadIdx.x < 32) icp_tile_issue_curr(a, s_c0, s_c1, g.tw, t.x0, t.y0, t.w, t.h, &s_bar[0]);
            if (threadIdx.x == 32) { s_box[0] = s_box[1] = 1 << 30; s_box[2] = s_box[3] = -1; }
            __syncthreads();
            mbar_wait(&s_bar[0], parity);
            icp_tile_bbox<kReduceThreads>(a, t, Rc, tc, Rpi, tp, s_box);
            __syncthreads();
            if (threadIdx.x < 32) {
                int mx0, my0, mwa, mha;
                icp_tile_window(s_box, a.rows, a.cols, g.mw, g.mh, kTileHaloX, mx0, my0, mwa, mha);
                if (threadIdx.x == 0) { s_win[0] = mx0; s_win[1] = my0; s_win[2] = mwa; s_win[3] = mha; }
                if (mwa > 0 && mha > 0) icp_tile_issue_model(a, s_g0, s_g1, s_gw, g.mw, mx0, my0, mwa, mha, &s_bar[1]);
            }
            __syncthreads();
            t.mx0 = s_win[0]; t.my0 = s_win[1]; t.mwa = s_win[2]; t.mha = s_win[3];
            if (t.mwa > 0 && t.mha > 0) mbar_wait(&s_bar[1], parity);
            else t.mwa = t.mha = 0;
            icp_tile_pass<kReduceThreads>(a, t, Rc, tc, Rpi, tp, acc);
            // both barriers complete one phase per tile only if the model stage ran: re-arm by re-initialising when it did not
            if (!(t.mwa > 0 && t.mha > 0)) {
                __syncthreads();
                if (threadIdx.x == 0) { mbar_init(&s_bar[1], 1); mbar_fence_init(); }
                // keep the parities of the two barriers in step: barrier 1 restarts at phase 0, so must barrier 0
                if (threadIdx.x == 0) { mbar_init(&s_bar[0], 1); mbar_fence_init(); }
                parity = 0;
                __syncthreads();
            } else {
                parity ^= 1u;
                __syncthreads();      // the tile's buffers are free for the next one
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
