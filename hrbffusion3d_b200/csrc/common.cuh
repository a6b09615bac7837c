// common.cuh -- shared helpers for the sm_100a HRBFFusion hot-path kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include "../../include/hrbf_b200.h"

namespace hrbf {

constexpr int kNumSMs = 148;              // B200: 2 dies x 74 SMs
constexpr int kReduceThreads = 256;
constexpr int kMaxReduceBlocks = 2 * kNumSMs;

void set_error(const char* fmt, ...);
extern std::atomic<unsigned long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

#define HRBF_CUDA(call)                                                                  \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            hrbf::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return HRBF_ERR_CUDA;                                                        \
        }                                                                                \
    } while (0)

#define HRBF_CHECK_ARG(cond)                                                             \
    do {                                                                                 \
        if (!(cond)) {                                                                   \
            hrbf::set_error("%s:%d invalid argument: %s", __FILE__, __LINE__, #cond);    \
            return HRBF_ERR_INVALID_ARG;                                                 \
        }                                                                                \
    } while (0)

#define HRBF_KERNEL_CHECK()                                                              \
    do {                                                                                 \
        hrbf::count_launch();                                                            \
        HRBF_CUDA(cudaGetLastError());                                                   \
    } while (0)

// Programmatic dependent launch: a kernel launched with HRBF_LAUNCH_PDL may be scheduled while the previous kernel of the
// stream is still draining (its launch latency disappears from the dependent chain of ~18 kernels per frame); pdl_wait() --
// the first statement of every such kernel -- blocks until that previous kernel has completed and its writes are visible,
// so the stream-order semantics are unchanged.  A no-op in a kernel launched the ordinary way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Packed fp32 pairs, mul.rn.f32x2 / add.rn.f32x2 spelled in PTX.  CAUTION, measured in this repository: neither the CUDA intrinsics
// __fmul2_rn / __fadd2_rn nor these PTX forms are protected against contraction the way scalar __fmul_rn / __fadd_rn are -- a packed
// product whose only use is a packed sum comes out of ptxas as ONE FFMA2 (the PCA covariance sums lost their bit-exactness that way).
// Where the scalar rounding sequence matters, keep the product scalar (or give it a second use), and check the SASS:
// tests/test_abi.py::test_exact_packed_sequences_are_not_contracted.
__device__ __forceinline__ float2 mul2_rn(float2 a, float2 b)
{
    float2 r;
    asm("{ .reg .b64 ma, mb, md; mov.b64 ma, {%2, %3}; mov.b64 mb, {%4, %5}; mul.rn.f32x2 md, ma, mb; mov.b64 {%0, %1}, md; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 add2_rn(float2 a, float2 b)
{
    float2 r;
    asm("{ .reg .b64 ma, mb, md; mov.b64 ma, {%2, %3}; mov.b64 mb, {%4, %5}; add.rn.f32x2 md, ma, mb; mov.b64 {%0, %1}, md; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
// lets the NEXT kernel of the stream (if launched with programmatic serialization) be scheduled as soon as SM resources free up,
// instead of after this grid has drained; it still blocks in its own pdl_wait() until this grid has completed
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define HRBF_LAUNCH_PDL(kernel, grid, block, smem, stream, ...)                          \
    do {                                                                                 \
        cudaError_t e__ = hrbf::launch_pdl(kernel, grid, block, smem, stream, __VA_ARGS__); \
        hrbf::count_launch();                                                            \
        if (e__ != cudaSuccess) { hrbf::set_error("%s:%d launch of %s -> %s", __FILE__, __LINE__, #kernel, cudaGetErrorString(e__)); return HRBF_ERR_CUDA; } \
    } while (0)

__device__ __forceinline__ float qnan() { return __int_as_float(0x7fffffff); }

struct Mat33 { float m[9]; };   // row-major
struct Vec3 { float x, y, z; };

__device__ __forceinline__ float3 mul(const float* M, float3 v)
{
    return make_float3(M[0] * v.x + M[1] * v.y + M[2] * v.z,
                       M[3] * v.x + M[4] * v.y + M[5] * v.z,
                       M[6] * v.x + M[7] * v.y + M[8] * v.z);
}
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator*(float s, float3 a) { return make_float3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float3 cross(float3 a, float3 b)
{
    return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float norm(float3 a) { return sqrtf(dot(a, a)); }

// pose (R[9], t[3]) -> its rigid inverse (R^T, -R^T t), the arithmetic of the host path (IndexMap.cpp:207, pose.inverse())
__device__ __forceinline__ void pose_inverse_dev(const float* pose, float* inv)
{
    float Ri[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Ri[i * 3 + j] = pose[j * 3 + i];
    for (int k = 0; k < 9; ++k) inv[k] = Ri[k];
    for (int i = 0; i < 3; ++i) inv[9 + i] = -(__fadd_rn(__fadd_rn(__fmul_rn(Ri[i * 3], pose[9]), __fmul_rn(Ri[i * 3 + 1], pose[10])), __fmul_rn(Ri[i * 3 + 2], pose[11])));
}
// HRBFFusion.cpp:1112-1123 : fusion weight from the inter-frame motion (|t| vs rotation angle)
__device__ __forceinline__ float velocity_weighting_dev(const float* curr, const float* last, float weightMultiplier)
{
    // diff = curr^-1 * last
    float R[9], t[3];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) R[i * 3 + j] = (curr[0 * 3 + i] * last[0 * 3 + j] + curr[1 * 3 + i] * last[1 * 3 + j]) + curr[2 * 3 + i] * last[2 * 3 + j];
        const float d0 = last[9] - curr[9], d1 = last[10] - curr[10], d2 = last[11] - curr[11];
        t[i] = (curr[0 * 3 + i] * d0 + curr[1 * 3 + i] * d1) + curr[2 * 3 + i] * d2;
    }
    const double rx = (double)R[7] - (double)R[5], ry = (double)R[2] - (double)R[6], rz = (double)R[3] - (double)R[1];
    const double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
    double c = ((double)((R[0] + R[4]) + R[8]) - 1.0) * 0.5;
    c = c > 1. ? 1. : c < -1. ? -1. : c;
    double theta = acos(c);
    if (s < 1e-5 && c > 0) theta = 0.0;
    const float tn = sqrtf((t[0] * t[0] + t[1] * t[1]) + t[2] * t[2]);
    float w = fmaxf(tn, (float)theta);
    const float largest = 0.01f, minWeight = 0.5f;
    if (w > largest) w = largest;
    return fmaxf(1.0f - (w / largest), minWeight) * weightMultiplier;
}

inline int div_up(int a, int b) { return (a + b - 1) / b; }

// Guard of an 8-slot staging ring (a pinned host buffer + its device twin per slot) that asynchronous copies and kernels read after
// the API call has returned: a slot is taken again only when the work enqueued by its previous user has completed.
struct SlotRing {
    cudaEvent_t ev[8] = {};
    bool used[8] = {};
    int next = 0;
    int acquire()      // returns the slot; blocks only if the call 8 uses ago is still in flight
    {
        const int k = next++ & 7;
        if (used[k]) cudaEventSynchronize(ev[k]);
        return k;
    }
    void release(int k, cudaStream_t s)      // call after the last enqueue that reads the slot
    {
        cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(s, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) { (void)cudaGetLastError(); used[k] = false; return; }
        if (!ev[k] && cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming) != cudaSuccess) { (void)cudaGetLastError(); ev[k] = nullptr; used[k] = false; return; }
        used[k] = cudaEventRecord(ev[k], s) == cudaSuccess;
    }
    void destroy() { for (int k = 0; k < 8; ++k) if (ev[k]) { cudaEventDestroy(ev[k]); ev[k] = nullptr; used[k] = false; } }
};

// SoA map view: plane k, row y, col x -> p[(k*rows + y)*pitch + x]   (pitch in elements)
struct SoA {
    float* p;
    int pitch;
};
struct CSoA {
    const float* p;
    int pitch;
};

}  // namespace hrbf
