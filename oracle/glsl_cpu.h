/*
 * oracle/glsl_cpu.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Just enough of the GLSL 3.30 language and built-ins, as C++ types and functions, to compile the REFERENCE's own shader
 * sources (Core/src/Shaders/*.frag / *.glsl, read where they lie under /root/reference by oracle/build_ref_glsl.py) for the
 * CPU and run their main() per fragment.  It lets the tests pin the oracle's restatement of the GLSL passes (SURVEY 8a rows
 * 6 and 10, FillIn) to the shader text itself.  What it is NOT: a GL implementation.  Arithmetic is IEEE fp32 evaluated by
 * the host compiler (-fsingle-precision-constant, no FMA contraction); a GPU's GLSL compiler may fuse and reorder, so the
 * comparison is to tolerance, never bit-for-bit.  Textures are sampled GL_NEAREST with clamp-to-edge (the reference creates
 * them with draw = false -> GL_NEAREST, GPUTexture.cpp:47 / pangolin GlTexture); see texel() for the boundary rule.
 */
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace glsl {

typedef unsigned int uint;

struct vec2;
struct vec3;
struct vec2_self { float x, y; inline operator vec2() const; };            // `v.xy` of a vec2
struct vec3_self { float x, y, z; inline operator vec3() const; };         // `v.xyz` of a vec3
struct vec2 {
    union {
        struct { float x, y; };
        struct { float r, g; };
        vec2_self xy;
    };
    vec2() = default;
    vec2(float a, float b) : x(a), y(b) {}
    explicit vec2(float a) : x(a), y(a) {}
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
};
struct vec3 {
    union {
        struct { float x, y, z; };
        struct { float r, g, b; };
        vec2 xy;
        vec3_self xyz;
    };
    vec3() = default;
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    explicit vec3(float a) : x(a), y(a), z(a) {}
    vec3(vec2 a, float c) : x(a.x), y(a.y), z(c) {}
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
};
struct vec4 {
    union {
        struct { float x, y, z, w; };
        struct { float r, g, b, a; };
        vec3 xyz;
        vec2 xy;
    };
    vec4() = default;
    vec4(float a_, float b_, float c_, float d_) : x(a_), y(b_), z(c_), w(d_) {}
    explicit vec4(float a_) : x(a_), y(a_), z(a_), w(a_) {}
    vec4(vec3 v, float d_) : x(v.x), y(v.y), z(v.z), w(d_) {}
    vec4(vec2 v, float c_, float d_) : x(v.x), y(v.y), z(c_), w(d_) {}
    explicit operator float() const { return x; }        // float(texture(...)): the first component
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
};
inline vec2_self::operator vec2() const { return vec2(x, y); }
inline vec3_self::operator vec3() const { return vec3(x, y, z); }
struct uvec4 {
    uint x, y, z, w;
    explicit operator uint() const { return x; }
    explicit operator float() const { return (float)x; }
    explicit operator int() const { return (int)x; }
};

#define GLSL_VEC_OPS(V, N)                                                                                         \
    inline V operator+(V a, V b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] + b[i]; return r; }                \
    inline V operator-(V a, V b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] - b[i]; return r; }                \
    inline V operator*(V a, V b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] * b[i]; return r; }                \
    inline V operator/(V a, V b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] / b[i]; return r; }                \
    inline V operator*(V a, float s) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] * s; return r; }               \
    inline V operator*(float s, V a) { V r; for (int i = 0; i < N; ++i) r[i] = s * a[i]; return r; }               \
    inline V operator/(V a, float s) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] / s; return r; }               \
    inline V operator+(V a, float s) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] + s; return r; }               \
    inline V operator-(V a, float s) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] - s; return r; }               \
    inline V operator-(V a) { V r; for (int i = 0; i < N; ++i) r[i] = -a[i]; return r; }                           \
    inline V& operator+=(V& a, V b) { for (int i = 0; i < N; ++i) a[i] += b[i]; return a; }                        \
    inline V& operator-=(V& a, V b) { for (int i = 0; i < N; ++i) a[i] -= b[i]; return a; }                        \
    inline V& operator*=(V& a, float s) { for (int i = 0; i < N; ++i) a[i] *= s; return a; }                       \
    inline V& operator/=(V& a, float s) { for (int i = 0; i < N; ++i) a[i] /= s; return a; }                       \
    inline bool operator==(V a, V b) { for (int i = 0; i < N; ++i) if (!(a[i] == b[i])) return false; return true; } \
    inline bool operator!=(V a, V b) { return !(a == b); }                                                         \
    inline float dot(V a, V b) { float s = a[0] * b[0]; for (int i = 1; i < N; ++i) s = s + a[i] * b[i]; return s; } \
    inline float length(V a) { return ::sqrtf(dot(a, a)); }                                                        \
    inline float distance(V a, V b) { return length(a - b); }                                                      \
    inline V normalize(V a) { return a / length(a); }                                                              \
    inline V abs(V a) { V r; for (int i = 0; i < N; ++i) r[i] = ::fabsf(a[i]); return r; }
GLSL_VEC_OPS(vec2, 2)
GLSL_VEC_OPS(vec3, 3)
GLSL_VEC_OPS(vec4, 4)
#undef GLSL_VEC_OPS

inline vec3 cross(vec3 a, vec3 b) { return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }

// column-major, m[col][row] like GLSL
struct mat4 {
    vec4 c[4];
    mat4() = default;
    vec4& operator[](int i) { return c[i]; }
    const vec4& operator[](int i) const { return c[i]; }
};
struct mat3 {
    vec3 c[3];
    mat3() = default;
    explicit mat3(const mat4& m) { c[0] = m[0].xyz; c[1] = m[1].xyz; c[2] = m[2].xyz; }      // upper-left 3x3
    explicit mat3(float d) { c[0] = vec3(d, 0, 0); c[1] = vec3(0, d, 0); c[2] = vec3(0, 0, d); }
    mat3(vec3 a, vec3 b, vec3 d) { c[0] = a; c[1] = b; c[2] = d; }
    mat3(float a0, float a1, float a2, float b0, float b1, float b2, float c0, float c1, float c2) { c[0] = vec3(a0, a1, a2); c[1] = vec3(b0, b1, b2); c[2] = vec3(c0, c1, c2); }
    vec3& operator[](int i) { return c[i]; }
    const vec3& operator[](int i) const { return c[i]; }
};
inline mat3 operator/(mat3 m, float s) { return mat3(m[0] / s, m[1] / s, m[2] / s); }
inline mat3 operator*(mat3 m, float s) { return mat3(m[0] * s, m[1] * s, m[2] * s); }
inline mat3 operator*(float s, mat3 m) { return m * s; }
inline mat3 operator+(mat3 a, mat3 b) { return mat3(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
inline mat3 operator-(mat3 a, mat3 b) { return mat3(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
inline vec3 operator*(mat3 m, vec3 v) { return m[0] * v.x + m[1] * v.y + m[2] * v.z; }
inline vec3 operator*(vec3 v, mat3 m) { return vec3(dot(v, m[0]), dot(v, m[1]), dot(v, m[2])); }
inline mat3 operator*(mat3 a, mat3 b) { return mat3(a * b[0], a * b[1], a * b[2]); }
inline mat3 transpose(mat3 m) { return mat3(vec3(m[0].x, m[1].x, m[2].x), vec3(m[0].y, m[1].y, m[2].y), vec3(m[0].z, m[1].z, m[2].z)); }
inline vec4 operator*(mat4 m, vec4 v) { return m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3] * v.w; }
// a row-major host 4x4 (as the callers keep their poses) -> GLSL's column-major mat4
inline mat4 mat4_from_row_major(const float* p)
{
    mat4 m;
    for (int c = 0; c < 4; ++c) m[c] = vec4(p[0 * 4 + c], p[1 * 4 + c], p[2 * 4 + c], p[3 * 4 + c]);
    return m;
}

// scalar built-ins, fp32 throughout
inline float sqrt(float x) { return ::sqrtf(x); }
inline float inversesqrt(float x) { return 1.0f / ::sqrtf(x); }
inline float abs(float x) { return ::fabsf(x); }
inline int abs(int x) { return x < 0 ? -x : x; }
inline float exp(float x) { return ::expf(x); }
inline float log(float x) { return ::logf(x); }
inline float pow(float x, float y) { return ::powf(x, y); }
inline float sin(float x) { return ::sinf(x); }
inline float cos(float x) { return ::cosf(x); }
inline float acos(float x) { return ::acosf(x); }
inline float asin(float x) { return ::asinf(x); }
inline float atan(float y, float x) { return ::atan2f(y, x); }
inline float atan(float x) { return ::atanf(x); }
inline float floor(float x) { return ::floorf(x); }
inline float ceil(float x) { return ::ceilf(x); }
inline float round(float x) { return ::roundf(x); }
inline float fract(float x) { return x - ::floorf(x); }
inline float sign(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }
inline bool isnan(float x) { return x != x; }
inline bool isinf(float x) { return std::isinf(x); }
inline float min(float a, float b) { return b < a ? b : a; }
inline float max(float a, float b) { return a < b ? b : a; }
inline int min(int a, int b) { return b < a ? b : a; }
inline int max(int a, int b) { return a < b ? b : a; }
inline float min(float a, int b) { return min(a, (float)b); }
inline float max(float a, int b) { return max(a, (float)b); }
inline float min(int a, float b) { return min((float)a, b); }
inline float max(int a, float b) { return max((float)a, b); }
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline vec3 mix(vec3 a, vec3 b, float t) { return a * (1.0f - t) + b * t; }

// ---- textures: dense row-major host images, GL_NEAREST, clamp-to-edge ----
struct sampler2D {
    const float* data = nullptr;   // [h][w][ch]
    int w = 0, h = 0, ch = 4;
};
struct usampler2D {
    const uint* data = nullptr;    // [h][w]
    int w = 0, h = 0;
};
// Texel selection.  Several shaders sample EXACTLY on a texel boundary (u = float(cx) / cols, depth_bilateral.frag, geometry.glsl):
// in fp32 (cx / cols) * cols lands a few ulps above or below cx, and a plain floor() would pick texel cx - 1 for about half of
// the columns.  Texture units do not work that way: the scaled coordinate is converted to fixed point with 8 fractional bits
// (round to nearest) before the integer part is taken, so a coordinate within 1/512 texel of a boundary selects the texel that
// starts there.  That model is used here; it is also what the oracle's integer-offset restatement assumes (SURVEY 8a hazard iii).
inline int texel(float u, int n)
{
    const float fixed = ::floorf(u * (float)n * 256.0f + 0.5f);      // 8 fractional bits, round to nearest
    int i = (int)::floorf(fixed / 256.0f);
    return i < 0 ? 0 : (i >= n ? n - 1 : i);
}
inline vec4 textureLod(const sampler2D& s, vec2 uv, float)
{
    const float* p = s.data + ((size_t)texel(uv.y, s.h) * s.w + texel(uv.x, s.w)) * s.ch;
    // missing components read (0, 0, 1) like a GL_RED / GL_LUMINANCE texture
    return vec4(p[0], s.ch > 1 ? p[1] : 0.f, s.ch > 2 ? p[2] : 0.f, s.ch > 3 ? p[3] : 1.f);
}
inline vec4 texture(const sampler2D& s, vec2 uv) { return textureLod(s, uv, 0.f); }
inline vec4 texture(const sampler2D& s, vec2 uv, float /* LOD bias: single-level textures */) { return textureLod(s, uv, 0.f); }
inline vec4 texture2D(const sampler2D& s, vec2 uv) { return textureLod(s, uv, 0.f); }
inline uvec4 textureLod(const usampler2D& s, vec2 uv, float)
{
    uvec4 r;
    r.x = s.data[(size_t)texel(uv.y, s.h) * s.w + texel(uv.x, s.w)];
    r.y = r.z = 0u; r.w = 1u;
    return r;
}
inline uvec4 texture(const usampler2D& s, vec2 uv) { return textureLod(s, uv, 0.f); }

}  // namespace glsl
