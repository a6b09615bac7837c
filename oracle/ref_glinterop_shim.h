/* oracle/ref_glinterop_shim.h -- TEST INFRASTRUCTURE ONLY.  Force-included in front of the reference's RGBDOdometry.cpp
 * (oracle/build_ref_odometry.py): the three CUDA-GL interop calls the init* functions make (Core/src/Utils/RGBDOdometry.cpp:165-168 and
 * alike) are served from a plain cudaArray, so that the file runs on a headless box.  Nothing else is touched. */
#pragma once
#include <cuda_runtime_api.h>
struct RefTexResource { cudaArray_t array; };
static inline cudaError_t refshim_map_resources(int, cudaGraphicsResource_t*, cudaStream_t = 0) { return cudaSuccess; }
static inline cudaError_t refshim_unmap_resources(int, cudaGraphicsResource_t*, cudaStream_t = 0) { return cudaSuccess; }
static inline cudaError_t refshim_mapped_array(cudaArray_t* a, cudaGraphicsResource_t r, unsigned int, unsigned int)
{
    *a = reinterpret_cast<RefTexResource*>(r)->array;
    return cudaSuccess;
}
#define cudaGraphicsMapResources refshim_map_resources
#define cudaGraphicsUnmapResources refshim_unmap_resources
#define cudaGraphicsSubResourceGetMappedArray refshim_mapped_array
