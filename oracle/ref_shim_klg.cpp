/*
 * oracle/ref_shim_klg.cpp -- TEST INFRASTRUCTURE ONLY.
 * extern "C" wrapper (ours) around the REFERENCE's own .klg reader, GUI/src/Tools/RawLogReader.cpp (+ Core/src/Utils/Resolution.cpp),
 * compiled unmodified where they lie by oracle/build_ref_host.sh into oracle/_ref/libref_klg.so.  Pangolin's FileExists and libjpeg's
 * declarations come from oracle/host_shims (neither is installed; the JPEG branch aborts and is never taken by the tests), Eigen/Core
 * (only included by Utils/Img.h) from oracle/eigen_mini, zlib is the system's.
 */
#include "RawLogReader.h"
#include <cstring>

extern "C" {
/* reads every frame of `path` with the reference reader; ts[n], depth[n][h][w] u16, rgb[n][h][w][3].  Returns the frame count (<= max_frames). */
int refk_read_klg(const char* path, int width, int height, int flip_colors, int max_frames, long long* ts, unsigned short* depth, unsigned char* rgb)
{
    Resolution::getInstance(width, height);
    RawLogReader r(path, flip_colors != 0);
    int n = 0;
    const size_t P = (size_t)width * height;
    while (r.hasMore() && n < max_frames) {
        r.getNext();
        ts[n] = r.timestamp;
        std::memcpy(depth + (size_t)n * P, r.depth, P * 2);
        std::memcpy(rgb + (size_t)n * P * 3, r.rgb, P * 3);
        ++n;
    }
    return n;
}
int refk_num_frames(const char* path, int width, int height)
{
    Resolution::getInstance(width, height);
    RawLogReader r(path, false);
    return r.getNumFrames();
}
}
