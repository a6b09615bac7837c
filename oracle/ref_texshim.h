/*
 * oracle/ref_texshim.h -- TEST INFRASTRUCTURE ONLY.  Force-included (nvcc -include) in front of the REFERENCE's unmodified
 * Core/src/Cuda/cudafuncs.cu so that it compiles with CUDA 12: texture REFERENCES (`texture<uchar4, 2, ...> inTex`,
 * cudafuncs.cu:896; cudaBindTextureToArray / tex2D(texref) / cudaUnbindTexture, :906-924) were removed from the toolkit in 12.0.
 * This header re-creates exactly that much of the old API on top of texture OBJECTS: the reference variable becomes a __managed__
 * struct holding a cudaTextureObject_t (point sampling, clamp, unnormalised coordinates, element type = the old defaults),
 * bind creates the object, tex2D(ref, x, y) fetches through it, unbind destroys it.  No reference source is touched or copied.
 */
#pragma once
#include <cuda_runtime.h>

template <class T, int dim, cudaTextureReadMode mode>
struct hrbf_ref_texture { cudaTextureObject_t obj; };

template <class T, int dim, cudaTextureReadMode mode>
static inline cudaError_t cudaBindTextureToArray(hrbf_ref_texture<T, dim, mode>& t, cudaArray* arr)
{
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = arr;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModePoint;
    td.readMode = mode;
    td.normalizedCoords = 0;
    cudaTextureObject_t o = 0;
    const cudaError_t e = cudaCreateTextureObject(&o, &rd, &td, nullptr);
    if (e != cudaSuccess) return e;
    cudaDeviceSynchronize();      // the managed variable is written from the host
    t.obj = o;
    return cudaSuccess;
}
template <class T, int dim, cudaTextureReadMode mode>
static inline cudaError_t cudaUnbindTexture(hrbf_ref_texture<T, dim, mode>& t)
{
    cudaDeviceSynchronize();
    const cudaError_t e = cudaDestroyTextureObject(t.obj);
    t.obj = 0;
    return e;
}
template <class T, int dim, cudaTextureReadMode mode>
static __device__ __forceinline__ T tex2D(const hrbf_ref_texture<T, dim, mode>& t, float x, float y)
{
    return tex2D<T>(t.obj, x, y);
}
/* `texture<uchar4, 2, cudaReadModeElementType> inTex;`  ->  `__managed__ hrbf_ref_texture<uchar4, 2, cudaReadModeElementType> inTex;` */
#define texture __managed__ hrbf_ref_texture
