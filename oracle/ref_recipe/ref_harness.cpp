// TEST INFRASTRUCTURE (oracle/_ref recipe) -- not product code.
//
// Runs the reference's OWN shader sources on the CPU.  gen.py rewrites the GLSL files lexically into C++ (no arithmetic is
// touched); this file supplies what the Vulkan driver supplies on a GPU: the GLSL environment (types and built-ins from the
// reference's vendored GLM), traceRayEXT (closest hit = oracle/orc_scene.cpp, the same driver stand-in the oracle uses),
// image load / store (binary16 rounding, SNORM8, out-of-bounds rules) and the linear / clamp sampler.  Everything the
// reference wrote itself -- raygen.h, closesthit.rchit, miss.rmiss, shadowMiss.rmiss, rough_prepare.comp, rough_blur.h,
// postprocess.comp, fxaa.h -- executes unmodified, so comparing its images with the oracle's pins the restatement.
// Single-threaded (shader globals).  Built only where /root/reference exists; the .so lands in oracle/_ref/.
#define GLM_FORCE_SWIZZLE
#define GLM_ENABLE_EXPERIMENTAL
#include <glm/glm.hpp>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../orc_render.h"
#include "../orc_scene.h"

namespace glsl {
using namespace glm;

// GLSL converts int / double arguments implicitly; GLM's templates do not.  Non-template overloads on float.
inline float clamp(float x, float a, float b) { return glm::clamp(x, a, b); }
inline vec3 clamp(vec3 x, float a, float b) { return glm::clamp(x, a, b); }
inline float mix(float x, float y, float a) { return glm::mix(x, y, a); }
inline float pow(float x, float y) { return glm::pow(x, y); }
inline float max(float x, float y) { return glm::max(x, y); }
inline float min(float x, float y) { return glm::min(x, y); }
inline int max(int x, int y) { return x < y ? y : x; }
inline int min(int x, int y) { return y < x ? y : x; }
inline float mod(float x, float y) { return glm::mod(x, y); }

struct accelerationStructureEXT {};
struct image2D { orc::half4* data = nullptr; int8_t* r8 = nullptr; int W = 0, H = 0; };
struct sampler2D { const image2D* img = nullptr; };

inline ivec2 imageSize(const image2D& im) { return ivec2(im.W, im.H); }
inline vec4 imageLoad(const image2D& im, ivec2 p) {
    if(p.x < 0 || p.y < 0 || p.x >= im.W || p.y >= im.H) return vec4(0.0f);   // OOB load = 0
    const size_t i = (size_t)p.y * im.W + p.x;
    if(im.r8) { const float f = (float)im.r8[i] / 127.0f; return vec4(f < -1.0f ? -1.0f : f, 0.0f, 0.0f, 1.0f); }   // R8_SNORM (raytracer.cpp:187)
    const orc::vec4 v = orc::unpack_half4(im.data[i]);
    return vec4(v.x, v.y, v.z, v.w);
}
inline void imageStore(image2D& im, ivec2 p, vec4 v) {
    if(p.x < 0 || p.y < 0 || p.x >= im.W || p.y >= im.H) return;                // OOB store dropped
    const size_t i = (size_t)p.y * im.W + p.x;
    if(im.r8) { float x = v.x; if(!(x == x)) x = 0.0f; x = x < -1.0f ? -1.0f : (x > 1.0f ? 1.0f : x); im.r8[i] = (int8_t)std::lrintf(x * 127.0f); return; }
    im.data[i] = orc::pack_half4(orc::vec4(v.x, v.y, v.z, v.w));
}
inline vec4 texelClamped(const image2D& im, int x, int y) {
    x = x < 0 ? 0 : (x >= im.W ? im.W - 1 : x); y = y < 0 ? 0 : (y >= im.H ? im.H - 1 : y);
    const orc::vec4 v = orc::unpack_half4(im.data[(size_t)y * im.W + x]);
    return vec4(v.x, v.y, v.z, v.w);
}
// linear filter, clamp-to-edge, one mip (compute_system.cpp:76-84)
inline vec4 textureLodOffset(const sampler2D& s, vec2 p, float, ivec2 o) {
    const image2D& im = *s.img;
    const float u = p.x * (float)im.W - 0.5f, v = p.y * (float)im.H - 0.5f;
    const float fu = std::floor(u), fv = std::floor(v);
    const float ax = u - fu, ay = v - fv;
    const int x0 = (int)fu + o.x, y0 = (int)fv + o.y;
    const vec4 c00 = texelClamped(im, x0, y0), c10 = texelClamped(im, x0 + 1, y0), c01 = texelClamped(im, x0, y0 + 1), c11 = texelClamped(im, x0 + 1, y0 + 1);
    const vec4 top = c00 * (1.0f - ax) + c10 * ax, bot = c01 * (1.0f - ax) + c11 * ax;
    return top * (1.0f - ay) + bot * ay;
}
inline vec4 textureLod(const sampler2D& s, vec2 p, float lod) { return textureLodOffset(s, p, lod, ivec2(0, 0)); }

// built-ins
struct GidXY {
    uint x, y;
    operator ivec2() const { return ivec2((int)x, (int)y); }
    operator vec2() const { return vec2((float)x, (float)y); }
};
inline vec2 operator+(GidXY a, vec2 b) { return vec2((float)a.x, (float)a.y) + b; }
struct Gid { uint x = 0, y = 0, z = 0; GidXY xy() const { return GidXY{x, y}; } };
Gid gl_GlobalInvocationID;
uvec3 gl_LaunchIDEXT, gl_LaunchSizeEXT;
vec3 gl_WorldRayOriginEXT, gl_WorldRayDirectionEXT;
float gl_HitTEXT;
uint gl_InstanceCustomIndexEXT, gl_PrimitiveID;
mat4x3 gl_ObjectToWorldEXT;
const uint gl_RayFlagsOpaqueEXT = 1u;

void traceRayEXT(accelerationStructureEXT, uint rayFlags, uint cullMask, uint sbtOffset, uint sbtStride, uint missIndex, vec3 origin, float tmin,
                 vec3 direction, float tmax, int payloadLocation);

#define hitAttributeEXT
#define rayPayloadEXT
#define rayPayloadInEXT
#define main shader_main

namespace raygen {
#include "raygen.rgen"
}
namespace closesthit {
#include "closesthit.rchit"
}
namespace miss0 {
#include "miss.rmiss"
}
namespace miss1 {
#include "shadowMiss.rmiss"
}
namespace rough_prepare {
#include "rough_prepare.comp"
}
namespace rough_blur_h {
#include "rough_blur_h.comp"
}
#undef IN_IMAGE
#undef OUT_IMAGE
#undef OFFSET_1
#undef OFFSET_2
namespace rough_blur_v {
#include "rough_blur_v.comp"
}
namespace postprocess {
#include "postprocess.comp"
}
namespace fxaa {
#include "fxaa.comp"
}
#undef main

// ------------------------------------------------------------------------------------------------ driver stand-in
const orc::Scene* g_scene = nullptr;
uint32_t g_flags = 0;
int g_depth = 0;
uint64_t g_rays = 0;

void traceRayEXT(accelerationStructureEXT, uint, uint cullMask, uint, uint, uint missIndex, vec3 origin, float tmin, vec3 direction, float tmax, int) {
    // per-invocation built-ins of the caller
    const vec3 so = gl_WorldRayOriginEXT, sd = gl_WorldRayDirectionEXT;
    const float st = gl_HitTEXT;
    const uint si = gl_InstanceCustomIndexEXT, sp = gl_PrimitiveID;
    const mat4x3 sm = gl_ObjectToWorldEXT;
    const vec3 sa = closesthit::attribs;
    if(g_depth == 0) std::memcpy(&closesthit::payload, &raygen::payload, sizeof(closesthit::payload));   // rayPayloadEXT -> rayPayloadInEXT
    static_assert(sizeof(closesthit::payload) == sizeof(raygen::payload) && sizeof(miss0::payload) == sizeof(raygen::payload) &&
                      sizeof(miss1::payload) == sizeof(raygen::payload), "payload.h must give one layout");
    ++g_depth;
    orc::Hit hit;
    bool found = false;
    if(cullMask != 0u) {
        ++g_rays;
        found = g_scene->closestHit(orc::vec3(origin.x, origin.y, origin.z), orc::vec3(direction.x, direction.y, direction.z), tmin, tmax, hit,
                                    (g_flags & orc::ORC_BRUTE_FORCE) != 0);
    }
    gl_WorldRayOriginEXT = origin; gl_WorldRayDirectionEXT = direction;
    if(found) {
        gl_HitTEXT = hit.t; gl_InstanceCustomIndexEXT = hit.inst; gl_PrimitiveID = hit.prim;
        const float* m = g_scene->instances[hit.inst].m;   // 3x4 row-major -> mat4x3 (4 columns of vec3)
        gl_ObjectToWorldEXT = mat4x3(m[0], m[4], m[8], m[1], m[5], m[9], m[2], m[6], m[10], m[3], m[7], m[11]);
        closesthit::attribs = vec3(hit.u, hit.v, 0.0f);
        closesthit::shader_main();
    } else if(missIndex == 0u) {
        std::memcpy(&miss0::payload, &closesthit::payload, sizeof(miss0::payload));
        miss0::shader_main();
        std::memcpy(&closesthit::payload, &miss0::payload, sizeof(miss0::payload));
    } else {
        std::memcpy(&miss1::payload, &closesthit::payload, sizeof(miss1::payload));
        miss1::shader_main();
        std::memcpy(&closesthit::payload, &miss1::payload, sizeof(miss1::payload));
    }
    --g_depth;
    if(g_depth == 0) std::memcpy(&raygen::payload, &closesthit::payload, sizeof(closesthit::payload));
    gl_WorldRayOriginEXT = so; gl_WorldRayDirectionEXT = sd; gl_HitTEXT = st; gl_InstanceCustomIndexEXT = si; gl_PrimitiveID = sp; gl_ObjectToWorldEXT = sm;
    closesthit::attribs = sa;
}

template <class U>
void fillUbo(U& u, const orc::Ubo& s) {
    std::memcpy(&u.viewInverse, s.viewInverse, 64);
    std::memcpy(&u.projInverse, s.projInverse, 64);
    u.clearColor = vec3(s.clearColor[0], s.clearColor[1], s.clearColor[2]);
    u.numSamples = s.numSamples;
    u.lightDir = vec3(s.lightDir[0], s.lightDir[1], s.lightDir[2]);
    u.maxRecursions = s.maxRecursions;
    u.time = s.time;
    u.showAlpha = (s.showAlpha & 0xffu) != 0;
    u.fadeColor = vec4(s.fadeColor[0], s.fadeColor[1], s.fadeColor[2], s.fadeColor[3]);
}

}  // namespace glsl

using namespace glsl;

extern "C" {

// raygen + recursion with the reference's shaders.  Scene arrays as for orc_scene_create.  Returns the number of rays traced.
uint64_t ref_trace(const void* vertices, uint32_t nVtx, const uint32_t* indices, uint32_t nIdx, const uint32_t* meshes, uint32_t nMeshes,
                   const void* materials, uint32_t nMat, const float* instXform, const uint32_t* instMeta, uint32_t nInst, const void* ubo192,
                   uint32_t W, uint32_t H, uint32_t flags, void* base, void* normal, void* rough) {
    orc::Scene s;
    s.vertices.assign((const orc::Vertex*)vertices, (const orc::Vertex*)vertices + nVtx);
    s.indices.assign(indices, indices + nIdx);
    s.meshes.resize(nMeshes);
    for(uint32_t m = 0; m < nMeshes; ++m) s.meshes[m] = orc::MeshRange{meshes[4 * m], meshes[4 * m + 1], meshes[4 * m + 2], meshes[4 * m + 3]};
    s.materials.assign((const orc::Material*)materials, (const orc::Material*)materials + nMat);
    s.instances.resize(nInst);
    for(uint32_t i = 0; i < nInst; ++i) {
        std::memcpy(s.instances[i].m, instXform + 12 * i, 48);
        s.instances[i].mesh = instMeta[4 * i]; s.instances[i].vtxOff = instMeta[4 * i + 1]; s.instances[i].idxOff = instMeta[4 * i + 2];
        s.instances[i].matOff = instMeta[4 * i + 3];
    }
    s.build();
    g_scene = &s; g_flags = flags; g_depth = 0; g_rays = 0;
    orc::Ubo u; std::memcpy(&u, ubo192, 192);
    fillUbo(raygen::ubo, u); fillUbo(closesthit::ubo, u); fillUbo(miss0::ubo, u);

    // buffers as the descriptor set binds them (raytracer.cpp:149-171)
    std::vector<closesthit::Vertex> V(nVtx);
    for(uint32_t i = 0; i < nVtx; ++i) {
        const orc::Vertex& a = s.vertices[i];
        V[i].position = vec3(a.px, a.py, a.pz); V[i].matIndex = a.matIndex; V[i].normal = vec3(a.nx, a.ny, a.nz); V[i].pad1 = 0;
    }
    std::vector<closesthit::Material> M(nMat);
    for(uint32_t i = 0; i < nMat; ++i) {
        const orc::Material& a = s.materials[i];
        M[i].diffuse = vec3(a.diffuse[0], a.diffuse[1], a.diffuse[2]); M[i].transparency = a.transparency;
        M[i].specular = vec3(a.specular[0], a.specular[1], a.specular[2]); M[i].reflectivity = a.reflectivity;
        M[i].roughness = a.roughness; M[i].ior = a.ior; M[i].effectId = a.effectId; M[i].rayConsumption = a.rayConsumption; M[i].emission = a.emission;
    }
    std::vector<closesthit::InstanceOffsetTableEntry> T(nInst);
    for(uint32_t i = 0; i < nInst; ++i) {
        T[i].vertexBufferOffset = s.instances[i].vtxOff; T[i].indexBufferOffset = s.instances[i].idxOff; T[i].materialBufferOffset = s.instances[i].matOff;
    }
    closesthit::vertices.v = V.data();
    closesthit::indices.i = const_cast<uint32_t*>(s.indices.data());
    closesthit::materials.m = M.data();
    closesthit::instanceOffsetTable.e = T.data();

    raygen::image = image2D{(orc::half4*)base, nullptr, (int)W, (int)H};
    raygen::normalImage = image2D{(orc::half4*)normal, nullptr, (int)W, (int)H};
    raygen::roughImage = image2D{(orc::half4*)rough, nullptr, (int)W, (int)H};
    gl_LaunchSizeEXT = uvec3(W, H, 1);
    for(uint32_t y = 0; y < H; ++y)
        for(uint32_t x = 0; x < W; ++x) {
            gl_LaunchIDEXT = uvec3(x, y, 0);
            raygen::shader_main();
        }
    g_scene = nullptr;
    return g_rays;
}

// The post chain with the reference's compute shaders, in the order and with the swap of raytracer.cpp:106-144.
void ref_post(const void* ubo192, uint32_t W, uint32_t H, int useFXAA, void* base, void* normal, void* rough, void* final_, void* roughA, void* roughB,
              int8_t* transitions) {
    orc::Ubo u; std::memcpy(&u, ubo192, 192);
    image2D iBase{(orc::half4*)base, nullptr, (int)W, (int)H}, iNormal{(orc::half4*)normal, nullptr, (int)W, (int)H},
        iRough{(orc::half4*)rough, nullptr, (int)W, (int)H}, iFinal{(orc::half4*)final_, nullptr, (int)W, (int)H},
        iA{(orc::half4*)roughA, nullptr, (int)W, (int)H}, iB{(orc::half4*)roughB, nullptr, (int)W, (int)H}, iT{nullptr, transitions, (int)W, (int)H};
#define RG_BIND(ns)                                                                                                                       \
    fillUbo(ns::ubo, u);                                                                                                                   \
    ns::finalImage = iFinal; ns::baseImage = iBase; ns::normalImage = iNormal; ns::roughImage = iRough; ns::roughTransitions = iT;        \
    ns::roughColorsA = iA; ns::roughColorsB = iB;                                                                                          \
    ns::finalSampler.img = &ns::finalImage; ns::baseSampler.img = &ns::baseImage; ns::normalSampler.img = &ns::normalImage;               \
    ns::roughSampler.img = &ns::roughImage; ns::roughTransitionsSampler.img = &ns::roughTransitions;                                      \
    ns::roughColorsASampler.img = &ns::roughColorsA; ns::roughColorsBSampler.img = &ns::roughColorsB;
    const uint32_t gw = (W + 15) / 16 * 16, gh = (H + 15) / 16 * 16;   // dispatch rounded up to 16x16 work groups (raytracer.cpp:108-109)
#define RG_DISPATCH(ns)                                                                  \
    for(uint32_t y = 0; y < gh; ++y)                                                     \
        for(uint32_t x = 0; x < gw; ++x) { gl_GlobalInvocationID = Gid{x, y, 0}; ns::shader_main(); }
    RG_BIND(rough_prepare) RG_DISPATCH(rough_prepare)
    RG_BIND(rough_blur_h) RG_BIND(rough_blur_v)
    for(int i = 0; i < 10; ++i) { RG_DISPATCH(rough_blur_h) RG_DISPATCH(rough_blur_v) }
    RG_BIND(postprocess) RG_DISPATCH(postprocess)
    if(useFXAA) {
        RG_BIND(fxaa) RG_DISPATCH(fxaa)
        // std::swap(m_baseImage, m_finalImage) (raytracer.cpp:138-140): the caller's arrays swap contents
        const size_t n = (size_t)W * H;
        std::vector<orc::half4> tmp((orc::half4*)base, (orc::half4*)base + n);
        std::memcpy(base, final_, n * sizeof(orc::half4));
        std::memcpy(final_, tmp.data(), n * sizeof(orc::half4));
    }
#undef RG_BIND
#undef RG_DISPATCH
}

}  // extern "C"
