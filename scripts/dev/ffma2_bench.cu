// Development microbenchmark: throughput of scalar FFMA against packed FFMA2 (fma.rn.f32x2) on sm_100a, and of the mixes the
// ALU-bound kernels of the frame use (FMUL / FADD packed).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int kIters = 4096, kChains = 8;
__global__ void __launch_bounds__(256) k_scalar(float* out, float a, float b)
{
    float x[kChains];
    for (int c = 0; c < kChains; ++c) x[c] = threadIdx.x * 1e-3f + c;
    for (int i = 0; i < kIters; ++i)
#pragma unroll
        for (int c = 0; c < kChains; ++c) x[c] = fmaf(x[c], a, b);
    float s = 0;
    for (int c = 0; c < kChains; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_packed(float* out, float a, float b)
{
    float2 x[kChains];
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int c = 0; c < kChains; ++c) x[c] = make_float2(threadIdx.x * 1e-3f + c, threadIdx.x * 2e-3f + c);
    for (int i = 0; i < kIters; ++i)
#pragma unroll
        for (int c = 0; c < kChains; ++c) x[c] = __ffma2_rn(x[c], a2, b2);
    float s = 0;
    for (int c = 0; c < kChains; ++c) s += x[c].x + x[c].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_mix_scalar(float* out, float a, float b)      // FMUL + FADD + FFMA per step, scalar
{
    float x[kChains];
    for (int c = 0; c < kChains; ++c) x[c] = threadIdx.x * 1e-3f + c;
    for (int i = 0; i < kIters; ++i)
#pragma unroll
        for (int c = 0; c < kChains; ++c) { const float d = __fsub_rn(x[c], b); const float e = __fmul_rn(d, d); x[c] = fmaf(e, a, x[c]); }
    float s = 0;
    for (int c = 0; c < kChains; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_mix_packed(float* out, float a, float b)
{
    float2 x[kChains];
    const float2 a2 = make_float2(a, a), nb2 = make_float2(-b, -b);
    for (int c = 0; c < kChains; ++c) x[c] = make_float2(threadIdx.x * 1e-3f + c, threadIdx.x * 2e-3f + c);
    for (int i = 0; i < kIters; ++i)
#pragma unroll
        for (int c = 0; c < kChains; ++c) { const float2 d = __fadd2_rn(x[c], nb2); const float2 e = __fmul2_rn(d, d); x[c] = __ffma2_rn(e, a2, x[c]); }
    float s = 0;
    for (int c = 0; c < kChains; ++c) s += x[c].x + x[c].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename K> static float run(K k, float* out, int blocks)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<blocks, 256>>>(out, 0.999f, 0.001f);
    cudaEventRecord(e0);
    for (int r = 0; r < 10; ++r) k<<<blocks, 256>>>(out, 0.999f, 0.001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / 10;
}
int main()
{
    int sms = 0, khz = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0); cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const int blocks = sms * 8;
    float* out; cudaMalloc(&out, (size_t)blocks * 256 * 4);
    const double thread_steps = (double)blocks * 256 * kIters * kChains;
    auto report = [&](const char* name, float ms, double flops_per_step) {
        const double per_clk_sm = thread_steps * flops_per_step / (ms * 1e-3) / ((double)khz * 1e3) / sms;
        printf("%-28s %8.3f ms   %7.1f fp32 lane-ops / clk / SM   (%.1f Tops/s)\n", name, ms, per_clk_sm, thread_steps * flops_per_step / (ms * 1e-3) / 1e12);
    };
    report("FFMA scalar", run(k_scalar, out, blocks), 1);
    report("FFMA2 packed", run(k_packed, out, blocks), 2);
    report("FADD+FMUL+FFMA scalar", run(k_mix_scalar, out, blocks), 3);
    report("FADD2+FMUL2+FFMA2 packed", run(k_mix_packed, out, blocks), 6);
    printf("SMs %d, clock %d MHz (nominal)\n", sms, khz / 1000);
    return 0;
}
