// TEST INFRASTRUCTURE -- CPU oracle for the Raygun per-frame ray-tracing path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// leg may use anything under oracle/.  The product (raygun_b200/) never links it.
//
// orc_math.h: GLSL built-in semantics used by the reference shaders, written out as
// strict IEEE-754 binary32 operations (build with -ffp-contract=off).  Where GLSL
// leaves the operation order open we follow GLM 0.9.9.5 (the reference's host math
// library, vendor/glm) so that oracle/_ref -- the reference's own shader sources
// compiled for the CPU against GLM -- can be compared bit for bit:
//   normalize  v * (1/sqrt(dot(v,v)))        glm/detail/func_geometric.inl:88
//   reflect    I - N * dot(N,I) * 2          glm/detail/func_geometric.inl (compute_reflect)
//   refract    k<0 ? 0 : eta*I-(eta*d+sqrt(k))*N   (compute_refract)
//   mix        x*(1-a) + y*a                 glm/detail/func_common.inl:87,110
//   clamp      min(max(x,lo),hi), min(x,y) = (y<x)?y:x, max(x,y) = (x<y)?y:x
//   mod        x - y*floor(x/y)
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace orc {

struct vec2 { float x, y; };
struct vec3 {
    float x, y, z;
    vec3() : x(0), y(0), z(0) {}
    explicit vec3(float s) : x(s), y(s), z(s) {}
    vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
struct vec4 {
    float x, y, z, w;
    vec4() : x(0), y(0), z(0), w(0) {}
    explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
    vec4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
    vec4(const vec3& v, float w_) : x(v.x), y(v.y), z(v.z), w(w_) {}
    vec3 xyz() const { return vec3(x, y, z); }
};

inline vec3 operator+(vec3 a, vec3 b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(vec3 a, vec3 b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(vec3 a, vec3 b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator*(vec3 a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, vec3 a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator/(vec3 a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 operator+(vec3 a, float s) { return vec3(a.x + s, a.y + s, a.z + s); }
inline vec3 operator-(float s, vec3 a) { return vec3(s - a.x, s - a.y, s - a.z); }
inline vec3 operator-(vec3 a) { return vec3(-a.x, -a.y, -a.z); }
inline vec3& operator+=(vec3& a, vec3 b) { a = a + b; return a; }
inline vec3& operator*=(vec3& a, vec3 b) { a = a * b; return a; }
inline vec3& operator*=(vec3& a, float s) { a = a * s; return a; }

inline vec4 operator+(vec4 a, vec4 b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline vec4 operator-(vec4 a, vec4 b) { return vec4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
inline vec4 operator*(vec4 a, float s) { return vec4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline vec4 operator/(vec4 a, float s) { return vec4(a.x / s, a.y / s, a.z / s, a.w / s); }
inline vec4& operator+=(vec4& a, vec4 b) { a = a + b; return a; }

inline float fmin_glsl(float x, float y) { return (y < x) ? y : x; }
inline float fmax_glsl(float x, float y) { return (x < y) ? y : x; }
inline float clampf(float x, float lo, float hi) { return fmin_glsl(fmax_glsl(x, lo), hi); }
inline vec3 clamp3(vec3 v, float lo, float hi) { return vec3(clampf(v.x, lo, hi), clampf(v.y, lo, hi), clampf(v.z, lo, hi)); }
inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float dot(vec4 a, vec4 b) { return (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w); }
inline float inversesqrt(float x) { return 1.0f / std::sqrt(x); }
inline vec3 normalize(vec3 v) { return v * inversesqrt(dot(v, v)); }
inline float length(vec3 v) { return std::sqrt(dot(v, v)); }
inline float length(vec4 v) { return std::sqrt(dot(v, v)); }
inline float distance(vec3 a, vec3 b) { return length(b - a); }
inline float distance(vec4 a, vec4 b) { return length(b - a); }
inline vec3 reflect(vec3 I, vec3 N) { return I - N * dot(N, I) * 2.0f; }
inline vec3 refract(vec3 I, vec3 N, float eta) {
    const float d = dot(N, I);
    const float k = 1.0f - eta * eta * (1.0f - d * d);
    return (k >= 0.0f) ? (eta * I - (eta * d + std::sqrt(k)) * N) : vec3(0.0f);
}
inline float mixf(float x, float y, float a) { return x * (1.0f - a) + y * a; }
inline vec3 mix(vec3 x, vec3 y, float a) { return x * (1.0f - a) + y * a; }
inline float modf_glsl(float x, float y) { return x - y * std::floor(x / y); }

// ---------------------------------------------------------------------------------
// IEEE binary16 <-> binary32 (round-to-nearest-even, NaN/Inf preserved): the storage
// semantics of the six rgba16f images (raygun/gpu/image.hpp:31).
inline uint16_t f32_to_f16(float f) {
    uint32_t x; std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    x &= 0x7fffffffu;
    if(x >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | ((x > 0x7f800000u) ? (0x0200u | ((x >> 13) & 0x3ffu)) : 0u));
    if(x >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);  // rounds to >= 65520 -> Inf
    if(x < 0x33000001u) return (uint16_t)sign;               // < 2^-25 (or == 2^-25: ties to even 0)
    const int e = (int)(x >> 23) - 127;
    uint32_t m = (x & 0x7fffffu) | 0x800000u;
    int shift; uint32_t he;
    if(e < -14) { shift = 13 + (-14 - e); he = 0; } else { shift = 13; he = (uint32_t)(e + 15); }
    const uint32_t halfway = 1u << (shift - 1), mask = (1u << shift) - 1u;
    uint32_t q = m >> shift;
    const uint32_t rem = m & mask;
    if(rem > halfway || (rem == halfway && (q & 1u))) q++;
    // q carries the implicit bit for normals; adding handles mantissa overflow into the exponent
    const uint32_t h = (he == 0) ? q : (((he - 1) << 10) + q);
    return (uint16_t)(sign | h);
}
inline float f16_to_f32(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    const uint32_t e = (h >> 10) & 0x1fu, m = h & 0x3ffu;
    uint32_t x;
    if(e == 0) {
        if(m == 0) x = sign;
        else { const float v = (float)m * 5.9604644775390625e-8f; std::memcpy(&x, &v, 4); x |= sign; }
    } else if(e == 31) x = sign | 0x7f800000u | (m << 13);
    else x = sign | ((e + 112u) << 23) | (m << 13);
    float f; std::memcpy(&f, &x, 4); return f;
}

struct half4 { uint16_t x, y, z, w; };
inline half4 pack_half4(vec4 v) { return half4{f32_to_f16(v.x), f32_to_f16(v.y), f32_to_f16(v.z), f32_to_f16(v.w)}; }
inline vec4 unpack_half4(half4 h) { return vec4(f16_to_f32(h.x), f16_to_f32(h.y), f16_to_f32(h.z), f16_to_f32(h.w)); }

}  // namespace orc
