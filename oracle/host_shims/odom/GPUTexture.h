/* oracle/host_shims/odom/GPUTexture.h -- TEST INFRASTRUCTURE ONLY.
 * GL-free stand-in for Core/src/GPUTexture.h, put in front of the reference's RGBDOdometry.{h,cpp} by oracle/build_ref_odometry.py:
 * the odometry only ever touches `cudaRes` (mapped to a cudaArray by the interop calls that oracle/ref_glinterop_shim.h serves). */
#ifndef GPUTEXTURE_H_
#define GPUTEXTURE_H_
#include <driver_types.h>
#include <cuda_runtime_api.h>
#include <string>
class GPUTexture
{
    public:
        GPUTexture() : texture(0), cudaRes(0), draw(false) {}
        void * texture;
        cudaGraphicsResource * cudaRes;
        const bool draw;
};
#endif
