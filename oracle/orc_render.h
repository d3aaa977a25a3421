// TEST INFRASTRUCTURE -- CPU oracle (see orc_math.h header).
#pragma once
#include "orc_scene.h"

namespace orc {

enum { ORC_FXAA = 1u, ORC_SRGB8 = 2u, ORC_STRICT_IEEE = 4u, ORC_BRUTE_FORCE = 8u, ORC_TRANSITIONS_UNORM = 16u };

// a "ray" = one traceRayEXT with non-zero cull mask; sky look-ups
// (cullMask 0, closesthit.rchit:144) are counted separately.
enum { RAY_PRIMARY = 0, RAY_SHADOW = 1, RAY_REFLECT = 2, RAY_REFRACT = 3, RAY_SKYLOOKUP = 4, RAY_KINDS = 5 };
struct Counters { uint64_t rays[RAY_KINDS] = {0, 0, 0, 0, 0}; uint64_t zeroDirRays = 0; };

struct PixelOut { vec4 base, normal, rough; uint32_t inst, prim; float t; };

void tracePixel(const Scene& scene, const Ubo& ubo, uint32_t W, uint32_t H, uint32_t px, uint32_t py, uint32_t flags, PixelOut& out,
                Counters& counters);

// The seven images of raygun/render/raytracer.cpp:173-195 (+ the 8-bit blit target).
struct Frame {
    uint32_t W, H;
    half4 *base, *normal, *rough, *final_, *roughA, *roughB;  // rgba16f
    int8_t* transitions;                                      // R8_SNORM (raytracer.cpp:187)
    uint8_t* rgba8;                                           // blit target (render_system.cpp:130-144)
};

// raytracer.cpp:106-144: rough_prepare, 10 x (blur_h, blur_v), postprocess, optional fxaa + swap.
// After the call `final_` holds what Raytracer::doRaytracing returns (m_finalImage after the swap)
// and `base` what the swap left there; rgba8 is the nearest blit of final_.
void postChain(const Ubo& ubo, Frame& f, uint32_t flags, int threads);

}  // namespace orc
