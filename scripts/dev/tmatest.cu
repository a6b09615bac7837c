// Development probe: 2-D tensor-map TMA loads of a [rows][cols] array of 16-byte pixels, variants of data type / box / descriptor home.
// nvcc -gencode arch=compute_100a,code=sm_100a -o build/tmatest scripts/dev/tmatest.cu && gpurun -- build/tmatest
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
struct alignas(64) Maps { CUtensorMap m; };
__global__ void k(const __grid_constant__ Maps maps, const CUtensorMap* gmap, int use_global, int cx, int cy, int box_bytes, float4* out, int n)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(box_bytes) : "memory");
        const void* d = use_global ? (const void*)gmap : (const void*)&maps.m;
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem)), "l"(d),
                     "r"(smem_u32(&bar)), "r"(cx), "r"(cy) : "memory");
    }
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = reinterpret_cast<float4*>(smem)[i];
}
int main()
{
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    auto encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    printf("encode fn %p qres %d\n", fn, (int)q);
    const int rows = 72, cols = 96;
    std::vector<float4> h(rows * cols);
    for (int i = 0; i < rows * cols; ++i) h[i] = make_float4((float)i, 1.f, 2.f, 3.f);
    float4 *d, *out; cudaMalloc(&d, h.size() * 16); cudaMalloc(&out, 1 << 20);
    cudaMemcpy(d, h.data(), h.size() * 16, cudaMemcpyHostToDevice);
    struct V { const char* name; CUtensorMapDataType dt; int epp; int bx, by; } vs[] = {
        { "u64 x2, box 12x2", CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, 12, 2 }, { "f32 x4, box 12x2", CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, 12, 2 },
        { "u64 x2, box 80x13", CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, 80, 13 }, { "f32 x4, box 64x13", CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, 64, 13 },
        { "u64 x2, box 24x10", CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, 24, 10 } };
    for (auto& v : vs)
        for (int use_global = 0; use_global < 2; ++use_global) {
            Maps m;
            const cuuint64_t dims[2] = { (cuuint64_t)cols * v.epp, rows }; const cuuint64_t strides[1] = { (cuuint64_t)cols * 16 };
            const cuuint32_t box[2] = { (cuuint32_t)(v.bx * v.epp), (cuuint32_t)v.by }; const cuuint32_t es[2] = { 1, 1 };
            CUresult r = encode(&m.m, v.dt, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            CUtensorMap* gm; cudaMalloc(&gm, sizeof(CUtensorMap)); cudaMemcpy(gm, &m.m, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
            const int n = v.bx * v.by, bytes = n * 16;
            cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
            const int ox = use_global ? -4 : 4, oy = use_global ? -3 : 3;      // the global-descriptor runs also probe NEGATIVE coordinates
            k<<<1, 128, bytes + 128>>>(m, gm, use_global, ox * v.epp, oy, bytes, out, n);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<float4> o(n);
            if (e == cudaSuccess) cudaMemcpy(o.data(), out, n * 16, cudaMemcpyDeviceToHost);
            int bad = 0;
            if (e == cudaSuccess) for (int y = 0; y < v.by; ++y) for (int x = 0; x < v.bx; ++x) { int gx = ox + x, gy = oy + y; float want = (gx >= 0 && gy >= 0 && gx < cols && gy < rows) ? (float)(gy * cols + gx) : 0.f; if (o[y * v.bx + x].x != want) ++bad; }
            printf("%-20s desc in %s: encode %d, run: %s, mismatches %d\n", v.name, use_global ? "global" : "param ", (int)r, cudaGetErrorString(e), bad);
            if (e != cudaSuccess) { printf("context lost, stop\n"); return 0; }
        }
    return 0;
}
