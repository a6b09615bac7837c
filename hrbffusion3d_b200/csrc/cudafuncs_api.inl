// cudafuncs_api.cu -- the remaining free functions of the reference's host<->CUDA seam (Core/src/Cuda/cudafuncs.cuh:65-239) as C-ABI
// entry points on PITCHED device arrays (the reference's DeviceArray2D is cudaMallocPitch memory): pyrDown, createVMap, createNMap,
// verticesToDepth, pyrDownGaussF, pyrDownUcharGauss, imageBGRToIntensity, computeDerivativeImages, projectToPointCloud.
// The frame pipeline does all of this inside prep_all_kernel / the persistent tracker; these one-to-one forms exist so that the
// reference's RGBDOdometry.cpp links against this library unchanged (include/hrbf_cudafuncs_compat.hpp).
// Included at the end of odometry.cu (same translation unit: it shares the kernels of odometry_kernels.cuh).

namespace {
inline dim3 grid2(int cols, int rows, dim3 b) { return dim3(div_up(cols, b.x), div_up(rows, b.y)); }
#define EL(step, T) ((int)((step) / sizeof(T)))

__global__ void k_pyr_down(int srows, int scols, const float* __restrict__ src, int sp, float* dst, int dp)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= scols / 2 || y >= srows / 2) return;
    dst[(size_t)y * dp + x] = depth_down_gated(srows, scols, x, y, [&](int cx, int cy) { return src[(size_t)cy * sp + cx]; });
}
__global__ void k_vertices_to_depth(int rows, int cols, const float4* __restrict__ v, float* depth, int dp, float cutoff)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    const float z = __ldg(v + (size_t)y * cols + x).z;      // cudafuncs.cu:874-885: NaN beyond the cut-off or at / behind the camera
    depth[(size_t)y * dp + x] = (z <= cutoff && z > 0.f) ? z : qnan();
}
__global__ void k_gauss_f(int srows, int scols, const float* __restrict__ src, int sp, float* dst, int dp)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= scols / 2 || y >= srows / 2) return;
    dst[(size_t)y * dp + x] = gauss_down_f32(srows, scols, x, y, [&](int cx, int cy) { return src[(size_t)cy * sp + cx]; });
}
__global__ void k_gauss_u8(int srows, int scols, const unsigned char* __restrict__ src, int sp, unsigned char* dst, int dp)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= scols / 2 || y >= srows / 2) return;
    dst[(size_t)y * dp + x] = gauss_down_u8(srows, scols, x, y, [&](int cx, int cy) { return src[(size_t)cy * sp + cx]; });
}
__global__ void k_intensity(int rows, int cols, const uchar4* __restrict__ rgba, unsigned char* dst, int dp)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    dst[(size_t)y * dp + x] = bgr_intensity(__ldg(rgba + (size_t)y * cols + x));
}
// computeDerivativeImages = applyKernel with the 3x3 Sobel pair (cudafuncs.cu:930-993): same tap order as sobel_pixel
__global__ void k_sobel(int rows, int cols, const unsigned char* __restrict__ src, int sp, short* dx, int xp, short* dy, int yp)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    const float gx[9] = { 1, 0, -1, 2, 0, -2, 1, 0, -1 }, gy[9] = { 1, 2, 1, 0, 0, 0, -1, -2, -1 };
    float dxv = 0, dyv = 0;
    int k = 8;
    for (int j = max(y - 1, 0); j <= min(y + 1, rows - 1); ++j)
        for (int i = max(x - 1, 0); i <= min(x + 1, cols - 1); ++i) {
            const float s = (float)src[(size_t)j * sp + i];
            dxv += s * gx[k]; dyv += s * gy[k];
            --k;
        }
    dx[(size_t)y * xp + x] = (short)dxv;
    dy[(size_t)y * yp + x] = (short)dyv;
}
__global__ void k_project(int rows, int cols, const float* __restrict__ depth, int dp, float* cloud3, int cp /* float3 elements per row */,
                          float invFx, float invFy, float cx, float cy)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    const float z = depth[(size_t)y * dp + x];
    float* o = cloud3 + 3 * ((size_t)y * cp + x);
    o[0] = (x - cx) * z * invFx; o[1] = (y - cy) * z * invFy; o[2] = z;
}
}  // namespace

extern "C" {

int hrbf_pyr_down(const float* src, size_t sstep, float* dst, size_t dstep, int srows, int scols, void* stream)
{
    HRBF_CHECK_ARG(src && dst && srows > 1 && scols > 1 && sstep >= scols * sizeof(float) && dstep >= (scols / 2) * sizeof(float));
    const dim3 b(32, 8);
    k_pyr_down<<<grid2(scols / 2, srows / 2, b), b, 0, (cudaStream_t)stream>>>(srows, scols, src, EL(sstep, float), dst, EL(dstep, float));
    HRBF_KERNEL_CHECK();
    return HRBF_OK;
}
int hrbf_create_vmap(hrbf_camera intr, const float* depth, size_t dstep, float* vmap, size_t vstep, int rows, int cols, float depthCutoff, float depthMapFactor, void* stream)
{
    HRBF_CHECK_ARG(depth && vmap && rows > 0 && cols > 0 && dstep == cols * sizeof(float) && vstep >= cols * sizeof(float));
    const dim3 b(32, 8);
    create_vmap_kernel<<<grid2(cols, rows, b), b, 0, (cudaStream_t)stream>>>(rows, cols, depth, vmap, EL(vstep, float), 1.f / intr.fx, 1.f / intr.fy, intr.cx, intr.cy, depthCutoff, depthMapFactor);
    HRBF_KERNEL_CHECK();
    return HRBF_OK;
}
int hrbf_create_nmap(const float* vmap, size_t vstep, float* nmap, size_t nstep, int rows, int cols, void* stream)
{
    HRBF_CHECK_ARG(vmap && nmap && rows > 0 && cols > 0 && vstep >= cols * sizeof(float) && nstep >= cols * sizeof(float));
    const dim3 b(32, 8);
    create_nmap_kernel<<<grid2(cols, rows, b), b, 0, (cudaStream_t)stream>>>(rows, cols, vmap, EL(vstep, float), nmap, EL(nstep, float));
    HRBF_KERNEL_CHECK();
    return HRBF_OK;
}
int hrbf_vertices_to_depth(const float* v_aos, float* depth, size_t dstep, int rows, int cols, float maxDepth, void* stream)
{
    HRBF_CHECK_ARG(v_aos && depth && rows > 0 && cols > 0 && dstep >= cols * sizeof(float));
    const dim3 b(32, 8);
    k_vertices_to_depth<<<grid2(cols, rows, b), b, 0, (cudaStream_t)stream>>>(rows, cols, (const float4*)v_aos, depth, EL(dstep, float), maxDepth);
    HRBF_KERNEL_CHECK();
    return HRBF_OK;
}
int hrbf_pyr_down_gauss_f(const float* src, size_t sstep, float* dst, size_t dstep, int srows, int scols, void* stream)
{
    HRBF_CHECK_ARG(src && dst && srows > 1 && scols > 1 && sstep >= scols * sizeof(float) && dstep >= (scols / 2) * sizeof(float));
    const dim3 b(32, 8);
    k_gauss_f<<<grid2(scols / 2, srows / 2, b), b, 0, (cudaStream_t)stream>>>(srows, scols, src, EL(sstep, float), dst, EL(dstep, float));
    HRBF_KERNEL_CHECK();
    return HRBF_OK;
}
int hrbf_pyr_down_uchar_gauss(const unsigned char* src, size_t sstep, unsigned char* dst, size_t dstep, int srows, int scols, void* stream)
{
    HRBF_CHECK_ARG(src && dst && srows > 1 && scols > 1 && sstep >= (size_t)scols && dstep >= (size_t)(scols / 2));
    const dim3 b(32, 8);
    k_gauss_u8<<<grid2(scols / 2, srows / 2, b), b, 0, (cudaStream_t)stream>>>(srows, scols, src, (int)sstep, dst, (int)dstep);
    HRBF_KERNEL_CHECK();
    return HRBF_OK;
}
int hrbf_image_bgr_to_intensity(const unsigned char* rgba8, unsigned char* dst, size_t dstep, int rows, int cols, void* stream)
{
    HRBF_CHECK_ARG(rgba8 && dst && rows > 0 && cols > 0 && dstep >= (size_t)cols);
    const dim3 b(32, 8);
    k_intensity<<<grid2(cols, rows, b), b, 0, (cudaStream_t)stream>>>(rows, cols, (const uchar4*)rgba8, dst, (int)dstep);
    HRBF_KERNEL_CHECK();
    return HRBF_OK;
}
int hrbf_compute_derivative_images(const unsigned char* src, size_t sstep, short* dx, size_t xstep, short* dy, size_t ystep, int rows, int cols, void* stream)
{
    HRBF_CHECK_ARG(src && dx && dy && rows > 0 && cols > 0 && sstep >= (size_t)cols && xstep >= cols * sizeof(short) && ystep >= cols * sizeof(short));
    const dim3 b(32, 8);
    k_sobel<<<grid2(cols, rows, b), b, 0, (cudaStream_t)stream>>>(rows, cols, src, (int)sstep, dx, EL(xstep, short), dy, EL(ystep, short));
    HRBF_KERNEL_CHECK();
    return HRBF_OK;
}
int hrbf_project_to_point_cloud(const float* depth, size_t dstep, float* cloud3, size_t cstep, hrbf_camera intr, int level, int rows, int cols, void* stream)
{
    HRBF_CHECK_ARG(depth && cloud3 && rows > 0 && cols > 0 && level >= 0 && level < 8 && dstep >= cols * sizeof(float) && cstep >= cols * 3 * sizeof(float));
    const int div = 1 << level;      // CameraModel::operator()(level), Cuda/types.cuh:93-97
    const float fx = intr.fx / div, fy = intr.fy / div, cx = intr.cx / div, cy = intr.cy / div;
    const dim3 b(32, 8);
    k_project<<<grid2(cols, rows, b), b, 0, (cudaStream_t)stream>>>(rows, cols, depth, EL(dstep, float), cloud3, (int)(cstep / (3 * sizeof(float))), 1.0f / fx, 1.0f / fy, cx, cy);
    HRBF_KERNEL_CHECK();
    return HRBF_OK;
}

}  // extern "C"
