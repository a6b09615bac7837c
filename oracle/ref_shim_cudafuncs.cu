/*
 * oracle/ref_shim_cudafuncs.cu -- TEST INFRASTRUCTURE ONLY.
 *
 * A thin extern "C" wrapper (our code) around the REFERENCE's own, unmodified map / pyramid kernels of
 * Core/src/Cuda/cudafuncs.cu (copyMaps, copyCurvatureMap, copyicpWeightMap, resize{V,N,C,icpWeight}Map,
 * tranformMaps, transformCurvMaps, createVMap, createNMap, pyrDown, verticesToDepth, pyrDownGaussF, pyrDownUcharGauss,
 * imageBGRToIntensity, computeDerivativeImages, projectToPointCloud), compiled where the file lies under /root/reference
 * by oracle/build_ref.sh into oracle/_ref/libref_cudafuncs.so (git-ignored; travels to the GPU box).  cudafuncs.cu uses one
 * texture REFERENCE (removed in CUDA 12): oracle/ref_texshim.h, force-included in front of it, restores that API on top of
 * texture objects, so the file itself is compiled unmodified.  No reference source is copied into this repository.
 *
 * Every function takes HOST pointers with the argument meaning of the oracle function of the same name (oracle/orc.h);
 * maps are dense SoA float[4*rows][cols], AoS sources float[rows][cols][4].
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
template <typename T>
void down(const DeviceArray2D<T>& d, void* host, int cols) { d.download(host, (size_t)cols * sizeof(T)); }
}  // namespace

extern "C" {

int ref5_copyMaps(int rows, int cols, const float* v_aos, const float* n_aos, float* vmap, float* nmap)
{
    DeviceArray<float> vs, ns;
    vs.upload(v_aos, (size_t)rows * cols * 4); ns.upload(n_aos, (size_t)rows * cols * 4);
    DeviceArray2D<float> vd(4 * rows, cols), nd(4 * rows, cols);
    copyMaps(vs, ns, vd, nd);
    cudaDeviceSynchronize();
    down(vd, vmap, cols); down(nd, nmap, cols);
    return 0;
}
int ref5_copyCurvatureMap(int rows, int cols, const float* c_aos, float* cmap, float thr)
{
    DeviceArray<float> cs;
    cs.upload(c_aos, (size_t)rows * cols * 4);
    DeviceArray2D<float> cd(4 * rows, cols);
    copyCurvatureMap(cs, cd, thr);
    cudaDeviceSynchronize();
    down(cd, cmap, cols);
    return 0;
}
int ref5_copyicpWeightMap(int rows, int cols, const float* w_src, float* w_dst)
{
    DeviceArray<float> ws;
    ws.upload(w_src, (size_t)rows * cols);
    DeviceArray2D<float> wd(rows, cols);
    copyicpWeightMap(ws, wd);
    cudaDeviceSynchronize();
    down(wd, w_dst, cols);
    return 0;
}
/* dst is IN/OUT: the reference kernel leaves some planes of invalid pixels unwritten (stale contents survive) */
int ref5_resizeMap(int drows, int dcols, const float* src, float* dst, int normalize)
{
    DeviceArray2D<float> s, d;
    up(s, src, 8 * drows, 2 * dcols); up(d, dst, 4 * drows, dcols);
    if (normalize) resizeNMap(s, d); else resizeVMap(s, d);
    down(d, dst, dcols);
    return 0;
}
int ref5_resizeCMap(int drows, int dcols, const float* src, float* dst)
{
    DeviceArray2D<float> s, d;
    up(s, src, 8 * drows, 2 * dcols); up(d, dst, 4 * drows, dcols);
    resizeCMap(s, d);
    cudaDeviceSynchronize();
    down(d, dst, dcols);
    return 0;
}
int ref5_resizeicpWeightMap(int drows, int dcols, const float* src, float* dst)
{
    DeviceArray2D<float> s, d;
    up(s, src, 2 * drows, 2 * dcols); up(d, dst, drows, dcols);
    resizeicpWeightMap(s, d);
    cudaDeviceSynchronize();
    down(d, dst, dcols);
    return 0;
}
/* vdst / ndst are IN/OUT for the same reason */
int ref5_tranformMaps(int rows, int cols, const float* vsrc, const float* nsrc, const float* R, const float* t, float* vdst, float* ndst)
{
    DeviceArray2D<float> vs, ns, vd, nd;
    up(vs, vsrc, 4 * rows, cols); up(ns, nsrc, 4 * rows, cols); up(vd, vdst, 4 * rows, cols); up(nd, ndst, 4 * rows, cols);
    tranformMaps(vs, ns, to_mat33(R), make_float3(t[0], t[1], t[2]), vd, nd);
    cudaDeviceSynchronize();
    down(vd, vdst, cols); down(nd, ndst, cols);
    return 0;
}
int ref5_transformCurvMaps(int rows, int cols, const float* k1src, const float* k2src, const float* R, const float* t, float* k1dst, float* k2dst)
{
    DeviceArray2D<float> as, bs, ad, bd;
    up(as, k1src, 4 * rows, cols); up(bs, k2src, 4 * rows, cols); up(ad, k1dst, 4 * rows, cols); up(bd, k2dst, 4 * rows, cols);
    transformCurvMaps(as, bs, to_mat33(R), make_float3(t[0], t[1], t[2]), ad, bd);
    cudaDeviceSynchronize();
    down(ad, k1dst, cols); down(bd, k2dst, cols);
    return 0;
}
int ref5_pyrDownDepth(int srows, int scols, const float* src, float* dst)
{
    DeviceArray2D<float> s, d;
    up(s, src, srows, scols);
    pyrDown(s, d);
    cudaDeviceSynchronize();
    down(d, dst, scols / 2);
    return 0;
}
int ref5_createVMap(float fx, float fy, float cx, float cy, int rows, int cols, const float* depth, float* vmap, float cutoff, float factor)
{
    DeviceArray2D<float> dp, v;
    up(dp, depth, rows, cols);
    // computeVmapKernel writes only plane 0 (NaN) of an invalid pixel: the other planes keep what the allocation held.  Defined here
    // (and in the oracle) as 0: create + clear first, createVMap's own create() of the same size is then a no-op.
    v.create(4 * rows, cols);
    cudaMemset2D(v.ptr(), v.step(), 0, (size_t)cols * sizeof(float), 4 * rows);
    createVMap(CameraModel(fx, fy, cx, cy), dp, v, cutoff, factor);
    cudaDeviceSynchronize();
    down(v, vmap, cols);
    return 0;
}
int ref5_createNMap(int rows, int cols, const float* vmap, float* nmap)
{
    DeviceArray2D<float> v, n;
    up(v, vmap, 4 * rows, cols);
    n.create(4 * rows, cols);      // same convention: planes the kernel leaves unwritten read 0
    cudaMemset2D(n.ptr(), n.step(), 0, (size_t)cols * sizeof(float), 4 * rows);
    createNMap(v, n);
    cudaDeviceSynchronize();
    down(n, nmap, cols);
    return 0;
}
int ref5_verticesToDepth(int rows, int cols, const float* v_aos, float* depth, float cutoff)
{
    DeviceArray<float> vs;
    vs.upload(v_aos, (size_t)rows * cols * 4);
    DeviceArray2D<float> d(rows, cols);
    verticesToDepth(vs, d, cutoff);
    cudaDeviceSynchronize();
    down(d, depth, cols);
    return 0;
}
int ref5_pyrDownGaussF(int srows, int scols, const float* src, float* dst)
{
    DeviceArray2D<float> s, d;
    up(s, src, srows, scols);
    pyrDownGaussF(s, d);
    cudaDeviceSynchronize();
    down(d, dst, scols / 2);
    return 0;
}
int ref5_pyrDownUcharGauss(int srows, int scols, const unsigned char* src, unsigned char* dst)
{
    DeviceArray2D<unsigned char> s, d;
    up(s, src, srows, scols);
    pyrDownUcharGauss(s, d);
    cudaDeviceSynchronize();
    down(d, dst, scols / 2);
    return 0;
}
/* imageBGRToIntensity reads the GL RGBA8 texture through a cudaArray (RGBDOdometry.cpp:701-718) */
int ref5_rgbaToIntensity(int rows, int cols, const unsigned char* rgba, unsigned char* dst)
{
    cudaArray* arr = nullptr;
    const cudaChannelFormatDesc desc = cudaCreateChannelDesc<uchar4>();
    if (cudaMallocArray(&arr, &desc, cols, rows) != cudaSuccess) return 1;
    if (cudaMemcpy2DToArray(arr, 0, 0, rgba, (size_t)cols * 4, (size_t)cols * 4, rows, cudaMemcpyHostToDevice) != cudaSuccess) { cudaFreeArray(arr); return 2; }
    DeviceArray2D<unsigned char> d(rows, cols);
    imageBGRToIntensity(arr, d);
    cudaDeviceSynchronize();
    down(d, dst, cols);
    cudaFreeArray(arr);
    return 0;
}
int ref5_sobel(int rows, int cols, const unsigned char* src, short* dx, short* dy)
{
    DeviceArray2D<unsigned char> s;
    up(s, src, rows, cols);
    DeviceArray2D<short> gx(rows, cols), gy(rows, cols);
    computeDerivativeImages(s, gx, gy);
    cudaDeviceSynchronize();
    down(gx, dx, cols); down(gy, dy, cols);
    return 0;
}
/* k = the intrinsics of level 0; the reference scales them by 2^level itself (CameraModel::operator()) */
int ref5_projectToPointCloud(int rows, int cols, const float* depth, float* cloud3, float fx, float fy, float cx, float cy, int level)
{
    DeviceArray2D<float> d;
    up(d, depth, rows, cols);
    DeviceArray2D<float3> c(rows, cols);
    CameraModel k(fx, fy, cx, cy);
    projectToPointCloud(d, c, k, level);
    down(c, cloud3, cols);
    return 0;
}

}  // extern "C"
