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

inline int div_up(int a, int b) { return (a + b - 1) / b; }

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
