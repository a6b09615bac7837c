/*
 * oracle/ref_shim_odometry.cpp -- TEST INFRASTRUCTURE ONLY.
 * extern "C" wrapper (ours) around the REFERENCE's own RGBDOdometry class (Core/src/Utils/RGBDOdometry.{h,cpp}, see
 * oracle/build_ref_odometry.py for what is and is not the reference in this library).  "Textures" are cudaArrays filled from host
 * memory; the calls mirror the class methods one to one, with the argument meaning of oracle/orc_py.Odometry.
 */
#include "RGBDOdometry.h"
#include <chrono>
#include <cstring>

namespace {
struct Tex {
    cudaArray_t arr = nullptr;
    RefTexResource res{};
    GPUTexture gt;
    int kind = -1, w = 0, h = 0;
};
struct RefOdom {
    RGBDOdometry* o = nullptr;
    Tex tex[16];
    int w = 0, h = 0;
};
GPUTexture* tex(RefOdom* r, int slot) { return &r->tex[slot].gt; }
Eigen::Matrix4f pose4(const float* p16)
{
    Eigen::Matrix4f m;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) m(i, j) = p16[i * 4 + j];
    return m;
}
}  // namespace

extern "C" {

void* refodom_create(int w, int h, float cx, float cy, float fx, float fy)
{
    if (cudaFree(0) != cudaSuccess) return nullptr;
    // the hot-path knobs RGBDOdometry reads from the global parameter singleton (GUI/GlobalStateParam.txt defaults)
    GlobalStateParam& g = GlobalStateParam::get();
    g.registrationICPUseSparseICP = false;
    g.registrationICPUseCoorespondenceSearch = false;
    g.registrationICPNeighborSearchRadius = 2;
    g.registrationColorUseRGBGrad = false;
    g.preprocessingCurvValidThreshold = 300;
    RefOdom* r = new RefOdom();
    r->w = w; r->h = h;
    r->o = new RGBDOdometry(w, h, cx, cy, fx, fy);
    return r;
}
void refodom_destroy(void* p)
{
    RefOdom* r = (RefOdom*)p;
    if (!r) return;
    for (auto& t : r->tex) if (t.arr) cudaFreeArray(t.arr);
    delete r->o;
    delete r;
}
void refodom_set_search(int use_search, int radius)
{
    GlobalStateParam::get().registrationICPUseCoorespondenceSearch = use_search != 0;
    GlobalStateParam::get().registrationICPNeighborSearchRadius = radius;
}
/* kind 0: RGBA32F (float[h][w][4]), 1: R32F (float[h][w]), 2: RGBA8 (uchar[h][w][4]) */
int refodom_set_texture(void* p, int slot, const void* host, int kind)
{
    RefOdom* r = (RefOdom*)p;
    Tex& t = r->tex[slot];
    if (t.arr && t.kind != kind) { cudaFreeArray(t.arr); t.arr = nullptr; }
    if (!t.arr) {
        cudaChannelFormatDesc d = kind == 0 ? cudaCreateChannelDesc(32, 32, 32, 32, cudaChannelFormatKindFloat)
                                  : kind == 1 ? cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat) : cudaCreateChannelDesc(8, 8, 8, 8, cudaChannelFormatKindUnsigned);
        if (cudaMallocArray(&t.arr, &d, r->w, r->h) != cudaSuccess) return -1;
        t.kind = kind;
        t.res.array = t.arr;
        t.gt.cudaRes = reinterpret_cast<cudaGraphicsResource*>(&t.res);
    }
    const size_t px = kind == 0 ? 16 : 4;
    return cudaMemcpy2DToArray(t.arr, 0, 0, host, r->w * px, r->w * px, r->h, cudaMemcpyHostToDevice) == cudaSuccess ? 0 : -1;
}
void refodom_initICP(void* p, int v, int n, float cutoff) { RefOdom* r = (RefOdom*)p; r->o->initICP(tex(r, v), tex(r, n), cutoff); }
void refodom_initICPModel(void* p, int v, int n, float cutoff, const float* pose16) { RefOdom* r = (RefOdom*)p; r->o->initICPModel(tex(r, v), tex(r, n), cutoff, pose4(pose16)); }
void refodom_initRGB(void* p, int rgba) { RefOdom* r = (RefOdom*)p; r->o->initRGB(tex(r, rgba)); }
void refodom_initRGBModel(void* p, int rgba) { RefOdom* r = (RefOdom*)p; r->o->initRGBModel(tex(r, rgba)); }
void refodom_initFirstRGB(void* p, int rgba) { RefOdom* r = (RefOdom*)p; r->o->initFirstRGB(tex(r, rgba)); }
void refodom_initCurvature(void* p, int k1, int k2) { RefOdom* r = (RefOdom*)p; r->o->initCurvature(tex(r, k1), tex(r, k2)); }
void refodom_initCurvatureModel(void* p, int k1, int k2, const float* pose16) { RefOdom* r = (RefOdom*)p; r->o->initCurvatureModel(tex(r, k1), tex(r, k2), pose4(pose16)); }
void refodom_initICPweight(void* p, int w) { RefOdom* r = (RefOdom*)p; r->o->initICPweight(tex(r, w)); }

/* stats8: lastICPError, lastICPCount, lastRGBError, lastRGBCount, lastSO3Error, lastSO3Count, wall microseconds of the call, 0 */
int refodom_track(void* p, float* trans3, float* rot9, int rgbOnly, float icpWeight, int pyramid, int fastOdom, int so3, int curv, float* stats8)
{
    RefOdom* r = (RefOdom*)p;
    Eigen::Vector3f t(trans3[0], trans3[1], trans3[2]);
    Eigen::Matrix<float, 3, 3, Eigen::RowMajor> R;
    std::memcpy(R.data(), rot9, 9 * sizeof(float));
    cudaDeviceSynchronize();
    const auto t0 = std::chrono::steady_clock::now();
    r->o->getIncrementalTransformation(t, R, rgbOnly != 0, icpWeight, pyramid != 0, fastOdom != 0, so3 != 0, curv != 0, 0);
    cudaDeviceSynchronize();
    const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
    for (int k = 0; k < 3; ++k) trans3[k] = t(k);
    std::memcpy(rot9, R.data(), 9 * sizeof(float));
    if (stats8) {
        stats8[0] = r->o->lastICPError; stats8[1] = r->o->lastICPCount; stats8[2] = r->o->lastRGBError; stats8[3] = r->o->lastRGBCount;
        stats8[4] = r->o->lastSO3Error; stats8[5] = r->o->lastSO3Count; stats8[6] = (float)us; stats8[7] = 0.f;
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
/* lastA (6x6 row-major) and lastb of the last Gauss-Newton iteration */
void refodom_last_system(void* p, double* A36, double* b6)
{
    RefOdom* r = (RefOdom*)p;
    std::memcpy(A36, r->o->lastA.data(), 36 * sizeof(double));
    std::memcpy(b6, r->o->lastb.data(), 6 * sizeof(double));
}
}
