/*
 * oracle/ref_shim.cu -- TEST INFRASTRUCTURE ONLY.
 *
 * A thin extern "C" wrapper (our code) around the REFERENCE's own, unmodified CUDA reduction
 * kernels: Core/src/Cuda/reduce.cu (icpStep / rgbStep / computeRgbResidual / so3Step) and
 * Core/src/Cuda/containers/device_memory.cpp, compiled where they lie under /root/reference by
 * oracle/build_ref.sh into oracle/_ref/libref_reduce.so (git-ignored; travels to the GPU box).
 * No reference source is copied into this repository.
 *
 * It is used by the `-m gpu` tests to pin the CPU oracle (oracle/orc_odometry.c) against the real
 * reference kernels on the same inputs, and by oracle/gen_ref_golden.py to produce the golden
 * vectors under tests/golden/ref_reduce_*.npz that the CPU-only tests check the oracle against.
 * All arguments are HOST pointers; maps are dense SoA float[4*rows][cols].
 */
#include "cudafuncs.cuh"
#include <cstring>

namespace {
mat33 to_mat33(const float* m)
{
    mat33 r;
    std::memcpy(r.data, m, sizeof(float) * 9);
    return r;
}
template <typename T>
void up(DeviceArray2D<T>& d, const void* host, int rows, int cols) { d.upload(host, (size_t)cols * sizeof(T), rows, cols); }
}  // namespace

extern "C" {

int ref_icpStep(int rows, int cols, const float* Rcurr, const float* tcurr,
                const float* vmap_curr, const float* nmap_curr, const float* ck1_curr, const float* ck2_curr,
                const float* Rprev_inv, const float* tprev, float fx, float fy, float cx, float cy,
                const float* vmap_g_prev, const float* nmap_g_prev, const float* ck1_g_prev, const float* ck2_g_prev,
                const float* icpw_g_prev, float distThres, float angleThres, int use_search, int radius, int use_weight,
                int threads, int blocks, float* A, float* b, float* residual, int* corres_out)
{
    DeviceArray2D<float> vc, nc, k1c, k2c, vg, ng, k1g, k2g, w;
    up(vc, vmap_curr, 4 * rows, cols); up(nc, nmap_curr, 4 * rows, cols); up(k1c, ck1_curr, 4 * rows, cols); up(k2c, ck2_curr, 4 * rows, cols);
    up(vg, vmap_g_prev, 4 * rows, cols); up(ng, nmap_g_prev, 4 * rows, cols); up(k1g, ck1_g_prev, 4 * rows, cols); up(k2g, ck2_g_prev, 4 * rows, cols);
    up(w, icpw_g_prev, rows, cols);
    DeviceArray2D<unsigned short> pm_c(rows, cols), pm_g(rows, cols);
    DeviceArray2D<int2> corres(rows, cols);
    DeviceArray2D<float4> cuda_out(rows, cols);
    DeviceArray2D<float3> zmap(rows, cols), lambda(rows, cols);
    DeviceArray<JtJJtrSE3> sum(blocks), out(1);
    const float3 tc = make_float3(tcurr[0], tcurr[1], tcurr[2]), tp = make_float3(tprev[0], tprev[1], tprev[2]);
    icpStep(to_mat33(Rcurr), tc, vc, nc, k1c, k2c, pm_c, 0, to_mat33(Rprev_inv), tp, CameraModel(fx, fy, cx, cy),
            vg, ng, k1g, k2g, w, pm_g, corres, cuda_out, zmap, lambda, distThres, angleThres, 0.0f,
            use_search != 0, radius, use_weight != 0, false, sum, out, A, b, residual, threads, blocks);
    if (corres_out) corres.download(corres_out, (size_t)cols * sizeof(int2));
    return 0;
}

int ref_computeRgbResidual(int rows, int cols, float minScale, const short* dIdx, const short* dIdy,
                           const float* lastDepth, const float* nextDepth,
                           const unsigned char* lastImage, const unsigned char* nextImage,
                           void* corresImg_out /* 16 B per pixel */, float maxDepthDelta,
                           const float* kt, const float* krkinv, int threads, int blocks, int* sigmaSum, int* count)
{
    DeviceArray2D<short> dx, dy;
    DeviceArray2D<float> ld, nd;
    DeviceArray2D<unsigned char> li, ni;
    up(dx, dIdx, rows, cols); up(dy, dIdy, rows, cols); up(ld, lastDepth, rows, cols); up(nd, nextDepth, rows, cols);
    up(li, lastImage, rows, cols); up(ni, nextImage, rows, cols);
    DeviceArray2D<DataTerm> corr(rows, cols);
    cudaMemset2D(corr.ptr(), corr.step(), 0, (size_t)cols * sizeof(DataTerm), rows);
    DeviceArray<int2> sumRes(MAX_THREADS);
    int s = 0, c = 0;
    computeRgbResidual(minScale, dx, dy, ld, nd, li, ni, corr, sumRes, maxDepthDelta,
                       make_float3(kt[0], kt[1], kt[2]), to_mat33(krkinv), s, c, threads, blocks);
    *sigmaSum = s; *count = c;
    /* the reference addresses corresImg LINEARLY (corresImg.data[k], reduce.cu:721,1057), ignoring the pitch */
    if (corresImg_out) cudaMemcpy(corresImg_out, corr.ptr(), (size_t)rows * cols * sizeof(DataTerm), cudaMemcpyDeviceToHost);
    return 0;
}

int ref_rgbStep(int rows, int cols, const void* corresImg /* 16 B per pixel */, float sigma, const float* cloud3,
                float fx, float fy, const short* dIdx, const short* dIdy, int use_grad_weight, float sobelScale,
                int threads, int blocks, float* A, float* b)
{
    DeviceArray2D<DataTerm> corr;
    DeviceArray2D<float3> cloud;
    DeviceArray2D<short> dx, dy;
    corr.create(rows, cols);   /* linear fill, see ref_computeRgbResidual */
    cudaMemcpy(corr.ptr(), corresImg, (size_t)rows * cols * sizeof(DataTerm), cudaMemcpyHostToDevice);
    up(cloud, cloud3, rows, cols); up(dx, dIdx, rows, cols); up(dy, dIdy, rows, cols);
    DeviceArray<JtJJtrSE3> sum(blocks), out(1);
    rgbStep(corr, sigma, cloud, fx, fy, dx, dy, use_grad_weight != 0, sobelScale, sum, out, A, b, threads, blocks);
    return 0;
}

int ref_so3Step(int rows, int cols, const unsigned char* lastImage, const unsigned char* nextImage,
                const float* imageBasis, const float* kinv, const float* krlr, int threads, int blocks,
                float* A, float* b, float* residual)
{
    DeviceArray2D<unsigned char> li, ni;
    up(li, lastImage, rows, cols); up(ni, nextImage, rows, cols);
    DeviceArray<JtJJtrSO3> sum(blocks), out(1);
    so3Step(li, ni, to_mat33(imageBasis), to_mat33(kinv), to_mat33(krlr), sum, out, A, b, residual, threads, blocks);
    return 0;
}

}  // extern "C"
