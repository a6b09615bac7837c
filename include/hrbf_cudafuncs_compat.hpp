/*
 * hrbf_cudafuncs_compat.hpp -- source-level drop-in for the reference's host<->CUDA seam
 * (Core/src/Cuda/cudafuncs.cuh:82-214): functions with the reference's NAMES and ARGUMENT ORDER
 * that forward to the C ABI of hrbf_b200.h.  Header-only, duck-typed on the reference's own
 * containers/types so that it does not include them:
 *     Map2D  : DeviceArray2D<T>   -- .ptr(), .step() [bytes], .rows(), .cols()
 *              (Cuda/containers/device_array.hpp:230-252)
 *     Map1D  : DeviceArray<T>     -- .ptr()
 *     Mat33  : mat33              -- float3 data[3], row-major      (Cuda/types.cuh:62-72)
 *     Vec3   : float3
 *     Cam    : CameraModel        -- fx, fy, cx, cy                 (Cuda/types.cuh:82-98)
 *
 * Usage in the reference tree: in Core/src/Utils/RGBDOdometry.cpp replace
 *     #include "../Cuda/cudafuncs.cuh"
 * by
 *     #include "../Cuda/containers/device_array.hpp"
 *     #include "../Cuda/types.cuh"
 *     #include <hrbf_cudafuncs_compat.hpp>
 * drop Cuda/reduce.cu and the listed cudafuncs.cu kernels from the build and link libhrbf_b200.so.
 * Differences from the reference, on purpose:
 *   - errors throw std::runtime_error (the reference prints and calls exit(0), convenience.cuh:64-71)
 *   - `sum` / `out` scratch arrays, `threads`, `blocks` are accepted and ignored (single-pass
 *     reduction sized for 148 SMs); scratch comes from hrbf_compat_workspace()
 *   - arguments the reference kernel never reads (plane_match maps, icp_weight_layers, cuda_out,
 *     z_thrinkMap, lambdaMap, curvatureThres, use_sparse_icp: reduce.cu:317-573) are ignored
 * tests/test_abi.py compiles this header against the reference's own container / type headers.
 */
#ifndef HRBF_CUDAFUNCS_COMPAT_HPP_
#define HRBF_CUDAFUNCS_COMPAT_HPP_

#include <cuda_runtime.h>
#include <stdexcept>
#include <string>

#include "hrbf_b200.h"

namespace hrbf_compat {

inline void check(int rc, const char* what)
{
    if (rc != HRBF_OK) throw std::runtime_error(std::string(what) + ": " + hrbf_last_error());
}

/* one lazily allocated reduction workspace per host thread (the reference's callers are single-threaded) */
inline void* workspace()
{
    static thread_local void* w = nullptr;
    if (!w && cudaMalloc(&w, hrbf_reduce_workspace_bytes()) != cudaSuccess) throw std::runtime_error("hrbf_compat: cudaMalloc(workspace) failed");
    return w;
}

template <class Mat33>
inline void mat_to_rows(const Mat33& m, float* r9)
{
    for (int i = 0; i < 3; ++i) { r9[3 * i] = m.data[i].x; r9[3 * i + 1] = m.data[i].y; r9[3 * i + 2] = m.data[i].z; }
}

}  // namespace hrbf_compat

/* icpStep, Cuda/cudafuncs.cuh:82-116 (definition Cuda/reduce.cu:580-693) */
template <class Mat33, class Vec3, class MapF, class MapU16, class Cam, class MapI2, class MapF4, class MapF3, class Sums>
inline void icpStep(const Mat33& Rcurr, const Vec3& tcurr, const MapF& vmap_curr, const MapF& nmap_curr, const MapF& ck1maps_curr,
                    const MapF& ck2maps_curr, const MapU16& /*plane_match_map_curr*/, const int /*icp_weight_layers*/,
                    const Mat33& Rprev_inv, const Vec3& tprev, const Cam& intr, const MapF& vmap_g_prev, const MapF& nmap_g_prev,
                    const MapF& ck1maps_g_prev, const MapF& ck2maps_g_prev, const MapF& icpWeightmap_g_prev,
                    const MapU16& /*plane_match_map_g*/, MapI2& corresICP, MapF4& /*cuda_out*/, MapF3& /*z_thrinkMap*/,
                    const MapF3& /*lambdaMap*/, float distThres, float angleThres, float /*curvatureThres*/,
                    bool icp_if_use_coorespondence_search, int icp_search_radius, bool icp_if_use_weight, bool /*use_sparse_icp*/,
                    Sums& /*sum*/, Sums& /*out*/, float* matrixA_host, float* vectorB_host, float* residual_host,
                    int /*threads*/, int /*blocks*/)
{
    float Rc[9], Rpi[9];
    hrbf_compat::mat_to_rows(Rcurr, Rc);
    hrbf_compat::mat_to_rows(Rprev_inv, Rpi);
    const float tc[3] = { tcurr.x, tcurr.y, tcurr.z }, tp[3] = { tprev.x, tprev.y, tprev.z };
    const hrbf_camera cam = { intr.fx, intr.fy, intr.cx, intr.cy };
    const hrbf_icp_options opt = { icp_if_use_coorespondence_search ? 1 : 0, icp_search_radius, icp_if_use_weight ? 1 : 0, distThres, angleThres };
    const int cols = vmap_curr.cols(), rows = vmap_curr.rows() / 4;      /* 4 planes stacked row-wise, RGBDOdometry.cpp:128-136 */
    /* corresICP is dense int2[rows][cols] in the reference (RGBDOdometry.cpp:140); only passed through when unpitched */
    int* corres = (corresICP.step() == (size_t)cols * 2 * sizeof(int)) ? (int*)corresICP.ptr() : nullptr;
    hrbf_compat::check(hrbf_icp_step(Rc, tc, vmap_curr.ptr(), nmap_curr.ptr(), ck1maps_curr.ptr(), ck2maps_curr.ptr(), vmap_curr.step(),
                                     Rpi, tp, cam, vmap_g_prev.ptr(), nmap_g_prev.ptr(), ck1maps_g_prev.ptr(), ck2maps_g_prev.ptr(),
                                     vmap_g_prev.step(), icpWeightmap_g_prev.ptr(), icpWeightmap_g_prev.step(), rows, cols, &opt, corres,
                                     hrbf_compat::workspace(), matrixA_host, vectorB_host, residual_host, nullptr, nullptr),
                       "icpStep");
}

/* rgbStep, Cuda/cudafuncs.cuh:118-132 (reduce.cu:842-896) */
template <class MapDT, class MapF3, class MapS, class Sums>
inline void rgbStep(const MapDT& corresImg, const float& sigma, const MapF3& cloud, const float& fx, const float& fy, const MapS& dIdx,
                    const MapS& dIdy, bool rgb_use_RGBGradient_weight, const float& sobelScale, Sums& /*sum*/, Sums& /*out*/,
                    float* matrixA_host, float* vectorB_host, int /*threads*/, int /*blocks*/)
{
    const int rows = dIdx.rows(), cols = dIdx.cols();
    if (dIdx.step() != (size_t)cols * sizeof(short) || cloud.step() != (size_t)cols * 3 * sizeof(float))
        throw std::runtime_error("rgbStep: hrbf_rgb_step takes dense (unpitched) images");
    hrbf_compat::check(hrbf_rgb_step((const hrbf_dataterm*)corresImg.ptr(), sigma, (const float*)cloud.ptr(), fx, fy, dIdx.ptr(), dIdy.ptr(),
                                     rgb_use_RGBGradient_weight ? 1 : 0, sobelScale, rows, cols, hrbf_compat::workspace(), matrixA_host,
                                     vectorB_host, nullptr, nullptr),
                       "rgbStep");
}

/* so3Step, Cuda/cudafuncs.cuh:134-145 (reduce.cu:1301-1359) */
template <class MapU8, class Mat33, class Sums>
inline void so3Step(const MapU8& lastImage, const MapU8& nextImage, const Mat33& imageBasis, const Mat33& kinv, const Mat33& krlr,
                    Sums& /*sum*/, Sums& /*out*/, float* matrixA_host, float* vectorB_host, float* residual_host, int /*threads*/, int /*blocks*/)
{
    float B[9], Ki[9], Kr[9];
    hrbf_compat::mat_to_rows(imageBasis, B);
    hrbf_compat::mat_to_rows(kinv, Ki);
    hrbf_compat::mat_to_rows(krlr, Kr);
    if (lastImage.step() != (size_t)lastImage.cols()) throw std::runtime_error("so3Step: hrbf_so3_step takes dense (unpitched) images");
    hrbf_compat::check(hrbf_so3_step(lastImage.ptr(), nextImage.ptr(), B, Ki, Kr, lastImage.rows(), lastImage.cols(), hrbf_compat::workspace(),
                                     matrixA_host, vectorB_host, residual_host, nullptr, nullptr),
                       "so3Step");
}

/* computeRgbResidual, Cuda/cudafuncs.cuh:147-163 (reduce.cu:1088-1154) */
template <class MapS, class MapF, class MapU8, class MapDT, class SumI2, class Vec3, class Mat33>
inline void computeRgbResidual(const float& minScale, const MapS& dIdx, const MapS& dIdy, const MapF& lastDepth, const MapF& nextDepth,
                               const MapU8& lastImage, const MapU8& nextImage, MapDT& corresImg, SumI2& /*sumResidual*/,
                               const float maxDepthDelta, const Vec3& kt, const Mat33& krkinv, int& sigmaSum, int& count,
                               int /*threads*/, int /*blocks*/)
{
    float K[9];
    hrbf_compat::mat_to_rows(krkinv, K);
    const float t[3] = { kt.x, kt.y, kt.z };
    if (nextImage.step() != (size_t)nextImage.cols()) throw std::runtime_error("computeRgbResidual: dense (unpitched) images expected");
    hrbf_compat::check(hrbf_compute_rgb_residual(minScale, dIdx.ptr(), dIdy.ptr(), lastDepth.ptr(), nextDepth.ptr(), lastImage.ptr(),
                                                 nextImage.ptr(), (hrbf_dataterm*)corresImg.ptr(), maxDepthDelta, t, K, nextImage.rows(),
                                                 nextImage.cols(), hrbf_compat::workspace(), &sigmaSum, &count, nullptr),
                       "computeRgbResidual");
}

/* tranformMaps [sic], Cuda/cudafuncs.cuh:171-176 (cudafuncs.cu:213-277) */
template <class MapF, class Mat33, class Vec3>
inline void tranformMaps(const MapF& vmap_src, const MapF& nmap_src, const Mat33& Rmat, const Vec3& tvec, MapF& vmap_dst, MapF& nmap_dst)
{
    float R[9];
    hrbf_compat::mat_to_rows(Rmat, R);
    const float t[3] = { tvec.x, tvec.y, tvec.z };
    hrbf_compat::check(hrbf_transform_maps(vmap_src.ptr(), vmap_src.step(), nmap_src.ptr(), nmap_src.step(), R, t, vmap_dst.ptr(), vmap_dst.step(),
                                           nmap_dst.ptr(), nmap_dst.step(), vmap_src.rows() / 4, vmap_src.cols(), nullptr),
                       "tranformMaps");
}

/* transformCurvMaps, Cuda/cudafuncs.cuh:178-182 (cudafuncs.cu:279-342) */
template <class MapF, class Mat33, class Vec3>
inline void transformCurvMaps(const MapF& curvk1_src, const MapF& curvk2_src, const Mat33& Rmat, const Vec3& tvec, MapF& curvk1_dst, MapF& curvk2_dst)
{
    float R[9];
    hrbf_compat::mat_to_rows(Rmat, R);
    const float t[3] = { tvec.x, tvec.y, tvec.z };
    hrbf_compat::check(hrbf_transform_curv_maps(curvk1_src.ptr(), curvk1_src.step(), curvk2_src.ptr(), curvk2_src.step(), R, t, curvk1_dst.ptr(),
                                                curvk1_dst.step(), curvk2_dst.ptr(), curvk2_dst.step(), curvk1_src.rows() / 4, curvk1_src.cols(), nullptr),
                       "transformCurvMaps");
}

/* copyMaps, Cuda/cudafuncs.cuh:184-187 (cudafuncs.cu:344-403): dst must already be created (4*rows x cols) */
template <class Map1F, class MapF>
inline void copyMaps(const Map1F& vmap_src, const Map1F& nmap_src, MapF& vmap_dst, MapF& nmap_dst)
{
    hrbf_compat::check(hrbf_copy_maps(vmap_src.ptr(), nmap_src.ptr(), vmap_dst.ptr(), vmap_dst.step(), nmap_dst.ptr(), nmap_dst.step(),
                                      vmap_dst.rows() / 4, vmap_dst.cols(), nullptr),
                       "copyMaps");
}
/* copyCurvatureMap, Cuda/cudafuncs.cuh:189-191 (cudafuncs.cu:405-449) */
template <class Map1F, class MapF>
inline void copyCurvatureMap(const Map1F& cmap_src, MapF& cmap_dst, const float curvatureThreshold)
{
    hrbf_compat::check(hrbf_copy_curvature_map(cmap_src.ptr(), cmap_dst.ptr(), cmap_dst.step(), cmap_dst.rows() / 4, cmap_dst.cols(), curvatureThreshold, nullptr),
                       "copyCurvatureMap");
}
/* copyicpWeightMap, Cuda/cudafuncs.cuh:193-194 (cudafuncs.cu:452-491) */
template <class Map1F, class MapF>
inline void copyicpWeightMap(const Map1F& icpwmap_src, MapF& icpwmap_dst)
{
    hrbf_compat::check(hrbf_copy_icpweight_map(icpwmap_src.ptr(), icpwmap_dst.ptr(), icpwmap_dst.step(), icpwmap_dst.rows(), icpwmap_dst.cols(), nullptr),
                       "copyicpWeightMap");
}

/* resize*Map, Cuda/cudafuncs.cuh:196-206 (cudafuncs.cu:526-743).  The reference (re)creates `output` at half size;
 * with its container that is output.create(rows/2, cols/2), kept here so callers need no change. */
#define HRBF_COMPAT_RESIZE(NAME, CALL, PLANES)                                                                                         \
    template <class MapF>                                                                                                              \
    inline void NAME(const MapF& input, MapF& output)                                                                                  \
    {                                                                                                                                  \
        const int in_rows = input.rows() / PLANES, in_cols = input.cols();                                                            \
        output.create((in_rows / 2) * PLANES, in_cols / 2);                                                                           \
        hrbf_compat::check(CALL(input.ptr(), input.step(), output.ptr(), output.step(), in_rows, in_cols, nullptr), #NAME);           \
    }
HRBF_COMPAT_RESIZE(resizeVMap, hrbf_resize_vmap, 4)
HRBF_COMPAT_RESIZE(resizeNMap, hrbf_resize_nmap, 4)
HRBF_COMPAT_RESIZE(resizeCMap, hrbf_resize_cmap, 4)
HRBF_COMPAT_RESIZE(resizeicpWeightMap, hrbf_resize_icpweight_map, 1)
#undef HRBF_COMPAT_RESIZE

/* ---- the GPUTest- and RGB-branch preparation functions (Cuda/cudafuncs.cuh:140-239), pitched arrays throughout ---- */
/* pyrDown, cudafuncs.cuh:177 (cudafuncs.cu:96-107): dst is (re)created at half size like the reference does */
template <class MapF>
inline void pyrDown(const MapF& src, MapF& dst)
{
    dst.create(src.rows() / 2, src.cols() / 2);
    hrbf_compat::check(hrbf_pyr_down(src.ptr(), src.step(), dst.ptr(), dst.step(), src.rows(), src.cols(), nullptr), "pyrDown");
}
/* createVMap / createNMap, cudafuncs.cuh:140-145 (cudafuncs.cu:138-211) */
template <class Cam, class MapF>
inline void createVMap(const Cam& intr, const MapF& depth, MapF& vmap, const float depthCutoff, const float mDepthMapFactor)
{
    vmap.create(depth.rows() * 4, depth.cols());
    if (depth.step() != (size_t)depth.cols() * sizeof(float)) throw std::runtime_error("createVMap: dense (unpitched) depth expected");
    const hrbf_camera cam = { intr.fx, intr.fy, intr.cx, intr.cy };
    hrbf_compat::check(hrbf_create_vmap(cam, depth.ptr(), depth.step(), vmap.ptr(), vmap.step(), depth.rows(), depth.cols(), depthCutoff, mDepthMapFactor, nullptr), "createVMap");
}
template <class MapF>
inline void createNMap(const MapF& vmap, MapF& nmap)
{
    nmap.create(vmap.rows(), vmap.cols());
    hrbf_compat::check(hrbf_create_nmap(vmap.ptr(), vmap.step(), nmap.ptr(), nmap.step(), vmap.rows() / 4, vmap.cols(), nullptr), "createNMap");
}
/* verticesToDepth, cudafuncs.cuh:218 (cudafuncs.cu:887-894): dst already created (rows x cols) */
template <class Map1F, class MapF>
inline void verticesToDepth(Map1F& vmap_src, MapF& dst, float cutOff)
{
    hrbf_compat::check(hrbf_vertices_to_depth(vmap_src.ptr(), dst.ptr(), dst.step(), dst.rows(), dst.cols(), cutOff, nullptr), "verticesToDepth");
}
/* pyrDownGaussF / pyrDownUcharGauss, cudafuncs.cuh:181-183 (cudafuncs.cu:794-871) */
template <class MapF>
inline void pyrDownGaussF(const MapF& src, MapF& dst)
{
    dst.create(src.rows() / 2, src.cols() / 2);
    hrbf_compat::check(hrbf_pyr_down_gauss_f(src.ptr(), src.step(), dst.ptr(), dst.step(), src.rows(), src.cols(), nullptr), "pyrDownGaussF");
}
template <class MapU8>
inline void pyrDownUcharGauss(const MapU8& src, MapU8& dst)
{
    dst.create(src.rows() / 2, src.cols() / 2);
    hrbf_compat::check(hrbf_pyr_down_uchar_gauss(src.ptr(), src.step(), dst.ptr(), dst.step(), src.rows(), src.cols(), nullptr), "pyrDownUcharGauss");
}
/* imageBGRToIntensity, cudafuncs.cuh:213 (cudafuncs.cu:913-928).  The reference reads a cudaArray through a texture reference; a
 * GL-free caller holds the RGBA8 image in linear device memory, which is what this overload takes. */
template <class MapU8>
inline void imageBGRToIntensity(const unsigned char* rgba8_dev, MapU8& dst)
{
    hrbf_compat::check(hrbf_image_bgr_to_intensity(rgba8_dev, dst.ptr(), dst.step(), dst.rows(), dst.cols(), nullptr), "imageBGRToIntensity");
}
/* computeDerivativeImages, cudafuncs.cuh:185 (cudafuncs.cu:956-993) */
template <class MapU8, class MapS>
inline void computeDerivativeImages(MapU8& src, MapS& dx, MapS& dy)
{
    hrbf_compat::check(hrbf_compute_derivative_images(src.ptr(), src.step(), dx.ptr(), dx.step(), dy.ptr(), dy.step(), src.rows(), src.cols(), nullptr),
                       "computeDerivativeImages");
}
/* projectToPointCloud, cudafuncs.cuh:216 (cudafuncs.cu:1015-1028) */
template <class MapF, class MapF3, class Cam>
inline void projectToPointCloud(const MapF& depth, const MapF3& cloud, Cam& intrinsics, const int& level)
{
    const hrbf_camera cam = { intrinsics.fx, intrinsics.fy, intrinsics.cx, intrinsics.cy };
    hrbf_compat::check(hrbf_project_to_point_cloud(depth.ptr(), depth.step(), (float*)cloud.ptr(), cloud.step(), cam, level, depth.rows(), depth.cols(), nullptr),
                       "projectToPointCloud");
}

#endif /* HRBF_CUDAFUNCS_COMPAT_HPP_ */
