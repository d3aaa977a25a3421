// TEST INFRASTRUCTURE -- CPU oracle (see orc_math.h header).
//
// orc_shade.cpp: literal restatement of the reference's ray-tracing shaders with ONE
// mutable payload per pixel and real recursion, so that every stale-state effect listed
// in SURVEY.md 8a "hazards" is reproduced by construction:
//   raygen      resources/shaders/raygen.rgen:30-39, raygen.h:37-115
//   closesthit  resources/shaders/closesthit.rchit:74-268
//   miss 0      resources/shaders/miss.rmiss:38-83
//   miss 1      resources/shaders/shadowMiss.rmiss:30-34
//   payload     resources/shaders/payload.h:29-43
// Float literals are binary32 as in GLSL; operation order follows GLM where GLSL leaves
// it open (orc_math.h).  Hazard 8 (0/0 in the reflect/refract weight, closesthit.rchit:257):
// default is NaN-free (weight := 0 when transparency + reflectivity == 0); strict IEEE on request.
#include "orc_render.h"

#include <cmath>

namespace orc {

namespace {

enum { RT_GENERIC = 0, RT_SHADOW_TRACE = 1, RT_SHADOW_INTERNAL = 2 };

struct Payload {  // payload.h:29-39
    vec3 hitValue; float reflectContribution; vec3 normal; vec4 roughValue;
    float depth, curIOR, refDepth; int rayType, recDepth;
};

struct Ctx {
    const Scene* scene; const Ubo* ubo; uint32_t flags;
    vec3 lightDir; int maxRecursions;
    Counters cnt;
    // primary-hit capture (sample 0 only)
    bool capture; uint32_t capInst, capPrim; float capT;
};

void traceRay(Ctx& c, Payload& payload, uint32_t cullMask, int missIndex, vec3 origin, float tmin, vec3 dir, float tmax, int kind);

// raygen.h:37-67
const float kAAOffsets[9][8][2] = {
    {{0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}},
    {{0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}},
    {{0.25f, 0.25f}, {-0.25f, -0.25f}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}},
    {{-0.125f, -0.375f}, {0.375f, -0.125f}, {-0.375f, 0.125f}, {0.125f, 0.375f}, {0, 0}, {0, 0}, {0, 0}, {0, 0}},
    {{-0.125f, -0.375f}, {0.375f, -0.125f}, {-0.375f, 0.125f}, {0.125f, 0.375f}, {0, 0}, {0, 0}, {0, 0}, {0, 0}},
    {{0.0625f, -0.1875f}, {-0.0625f, 0.1875f}, {0.3125f, 0.0625f}, {-0.1875f, -0.3125f}, {-0.3125f, 0.3125f}, {-0.4375f, -0.0625f}, {0.1875f, 0.4375f}, {0.4375f, -0.4375f}},
    {{0.0625f, -0.1875f}, {-0.0625f, 0.1875f}, {0.3125f, 0.0625f}, {-0.1875f, -0.3125f}, {-0.3125f, 0.3125f}, {-0.4375f, -0.0625f}, {0.1875f, 0.4375f}, {0.4375f, -0.4375f}},
    {{0.0625f, -0.1875f}, {-0.0625f, 0.1875f}, {0.3125f, 0.0625f}, {-0.1875f, -0.3125f}, {-0.3125f, 0.3125f}, {-0.4375f, -0.0625f}, {0.1875f, 0.4375f}, {0.4375f, -0.4375f}},
    {{0.0625f, -0.1875f}, {-0.0625f, 0.1875f}, {0.3125f, 0.0625f}, {-0.1875f, -0.3125f}, {-0.3125f, 0.3125f}, {-0.4375f, -0.0625f}, {0.1875f, 0.4375f}, {0.4375f, -0.4375f}},
};

// miss.rmiss:38-74.  normalize(0) would be NaN (hazard 7): a zero direction is kept as 0 (ORC_STRICT_IEEE: literal NaN).
vec3 skyMix(const Ctx& c, vec3 worldRayDir, vec3 sunTone, vec3 skyTone, vec3 scatterTone, float scatterFactor, float powFactor) {
    const bool zero = worldRayDir.x == 0.0f && worldRayDir.y == 0.0f && worldRayDir.z == 0.0f;
    const vec3 rayDir = (zero && !(c.flags & ORC_STRICT_IEEE)) ? vec3(0.0f) : normalize(worldRayDir);
    const float y = std::fabs(worldRayDir.y + 1.5f) / 3.0f;

    float sun = 1.0f - distance(rayDir, normalize(-c.lightDir));
    sun = clampf(sun, 0.0f, 2.0f);

    float glow = sun;
    glow = clampf(glow, 0.0f, 1.0f);

    sun = std::pow(sun, powFactor);
    sun *= 1000.0f;
    sun = clampf(sun, 0.0f, 16.0f);

    glow = std::pow(glow, 6.0f) * 1.0f;
    glow = std::pow(glow, y);
    glow = clampf(glow, 0.0f, 1.0f);

    sun *= std::pow(y * y, 1.0f / 1.65f);

    glow *= std::pow(y * y, 1.0f / 2.0f);

    sun += glow;

    const vec3 sunColor = sunTone * sun;

    const float atmosphere = std::sqrt(1.0f - y);

    float scatter = std::pow(4.0f - c.lightDir.y, 1.0f / 15.0f);
    scatter = 1.0f - clampf(scatter, 0.8f, 1.0f);

    const vec3 scatterColor = mix(vec3(1.0f), scatterTone * 1.5f, scatter);
    const vec3 skyScatter = mix(skyTone, scatterColor, atmosphere / scatterFactor);

    return sunColor + skyScatter;
}

// miss.rmiss:76-83
void miss0(const Ctx& c, Payload& payload, vec3 worldRayDir) {
    const vec3 res = skyMix(c, worldRayDir, vec3(1.0f, 0.6f, 0.05f), vec3(0.2f, 0.4f, 0.8f), vec3(1.0f, 0.3f, 0.0f), 1.3f, 80.0f);
    payload.hitValue = res;
    payload.roughValue = vec4(res, 0.0f);
    payload.depth = 10000.f;
}
// shadowMiss.rmiss:30-34
void miss1(Payload& payload) { payload.hitValue = vec3(1.0f, 1.0f, 1.0f); }

struct Mat {  // local copy of gpu::Material (closesthit.rchit:109) as GLSL types
    vec3 diffuse; float transparency; vec3 specular; float reflectivity;
    float roughness, ior; uint32_t effectId, rayConsumption; float emission;
};

// closesthit.rchit:74-91
void gridEffect(const Payload& payload, float hitT, Mat& mat, vec3 pos) {
    const float aa = (payload.refDepth + hitT + 8.0f) / 30.0f;
    const float aa2 = aa / 2.0f;

    float minmod = fmin_glsl(std::fabs(modf_glsl((pos.x + 1000.0f) * 10.0f + aa2, 20.0f) - aa2),
                             std::fabs(modf_glsl((pos.z + 1000.0f) * 10.0f + aa2, 20.0f) - aa2));
    if(minmod < aa2) {
        minmod -= aa2 - std::pow(aa, 2.0f) / 3.0f;
        minmod *= 3.0f / std::pow(aa, 2.0f);
        mat.diffuse *= mixf(aa / 10.0f, 1.0f, minmod);
        mat.specular *= mixf(aa / 10.0f, 1.0f, minmod);
        mat.reflectivity *= mixf(aa / 10.0f, 1.0f, minmod);
    }

    if(modf_glsl((pos.x + 1000.0f) * 5.0f, 20.0f) < 10.0f && modf_glsl((pos.z + 1000.0f) * 5.0f, 20.0f) < 10.0f) {
        mat.reflectivity *= 1.5f;
    }
}

// closesthit.rchit:93-268.  worldRayOrigin/Direction/hitT/attribs/instance/primitive are the
// gl_* built-ins of the invocation.
void closestHit(Ctx& c, Payload& payload, vec3 worldRayOrigin, vec3 worldRayDirection, const Hit& hit) {
    const Scene& s = *c.scene;
    const float hitT = hit.t;
    const int maxRecursions = c.maxRecursions;
    // :96-109
    const vec3 barycentrics(1.0f - hit.u - hit.v, hit.u, hit.v);
    const Instance& inst = s.instances[hit.inst];
    const uint32_t i0 = s.indices[inst.idxOff + 3 * hit.prim + 0];
    const uint32_t i1 = s.indices[inst.idxOff + 3 * hit.prim + 1];
    const uint32_t i2 = s.indices[inst.idxOff + 3 * hit.prim + 2];
    const Vertex& v0 = s.vertices[inst.vtxOff + i0];
    const Vertex& v1 = s.vertices[inst.vtxOff + i1];
    const Vertex& v2 = s.vertices[inst.vtxOff + i2];
    const Material& gm = s.materials[inst.matOff + v0.matIndex];
    Mat mat{vec3(gm.diffuse[0], gm.diffuse[1], gm.diffuse[2]), gm.transparency, vec3(gm.specular[0], gm.specular[1], gm.specular[2]),
            gm.reflectivity, gm.roughness, gm.ior, gm.effectId, gm.rayConsumption, gm.emission};
    // gpu_material.def documents rayConsumption as 1..5; the kernels clamp it to 1..8 so that every reflection advances recDepth (same here)
    mat.rayConsumption = mat.rayConsumption < 1u ? 1u : (mat.rayConsumption > 8u ? 8u : mat.rayConsumption);

    // :111-114
    const float tmin = 0.01f;
    const float tmax = 1000.0f;
    const vec3 origin = worldRayOrigin + worldRayDirection * hitT;

    // :117-119  mat3(gl_ObjectToWorldEXT) * vn, then normalize
    const vec3 vn = vec3(v0.nx, v0.ny, v0.nz) * barycentrics.x + vec3(v1.nx, v1.ny, v1.nz) * barycentrics.y + vec3(v2.nx, v2.ny, v2.nz) * barycentrics.z;
    const float* m = inst.m;
    vec3 vnInWorldSpace = normalize(vec3(m[0] * vn.x + m[1] * vn.y + m[2] * vn.z, m[4] * vn.x + m[5] * vn.y + m[6] * vn.z,
                                         m[8] * vn.x + m[9] * vn.y + m[10] * vn.z));

    // :122
    if(mat.effectId == 1) gridEffect(payload, hitT, mat, origin);

    // :125-152
    if(payload.rayType == RT_SHADOW_INTERNAL) {
        const float thicknessModulation = clampf(hitT * (1.0f - mat.transparency) * 10.0f, 0.0f, 1.0f);
        vec3 shadowCol = payload.hitValue - mix(vec3(0.0f), normalize(1.1f - mat.diffuse) + 0.1f, thicknessModulation);

        if(payload.recDepth < maxRecursions) {
            payload.rayType = RT_SHADOW_TRACE;
            payload.recDepth++;
            traceRay(c, payload, 0xFF, 0, origin, tmin, worldRayDirection, tmax, RAY_SHADOW);
            payload.recDepth--;

            if(payload.depth < 1000.0f) {
                payload.hitValue *= shadowCol;
            } else {
                const float eta = mat.ior / 1.0f;
                const vec3 dir = refract(worldRayDirection, vnInWorldSpace, eta);
                const float dot_product = std::pow(dot(c.lightDir, dir), 5.0f) + 0.75f;
                shadowCol *= dot_product;
                traceRay(c, payload, 0x0, 0, origin, tmin, -dir, tmax, RAY_SKYLOOKUP);
                payload.hitValue = shadowCol + 0.1f * payload.hitValue;
            }
        } else {
            payload.hitValue = shadowCol * vec3(0.4f);
        }
        return;
    }
    // :153-166
    if(payload.rayType == RT_SHADOW_TRACE) {
        if(mat.transparency > 0.0f) {
            if(payload.recDepth < maxRecursions) {
                payload.rayType = RT_SHADOW_INTERNAL;
                payload.recDepth++;
                traceRay(c, payload, 0xFF, 1, origin, tmin, worldRayDirection, tmax, RAY_SHADOW);
                payload.recDepth--;
            }
        } else {
            payload.hitValue *= mix(vec3(0.4f), vec3(0.8f), clampf(std::log(hitT) / 8.0f, 0.0f, 1.0f));
        }
        return;
    }

    // :170-171
    const bool frontFacing = dot(-worldRayDirection, vnInWorldSpace) > 0.0f;
    if(!frontFacing) vnInWorldSpace = normalize(-vnInWorldSpace);

    // :174-175
    const float dot_product = fmax_glsl(dot(-c.lightDir, vnInWorldSpace), 0.2f);
    vec3 baseColor = dot_product * mat.diffuse;

    // :187-204
    vec3 shadowColor(1.0f, 1.0f, 1.0f);
    if(dot(-c.lightDir, vnInWorldSpace) > 0.07f) {
        if(payload.recDepth < maxRecursions) {
            payload.hitValue = vec3(1.0f);
            payload.rayType = RT_SHADOW_TRACE;
            payload.recDepth++;
            traceRay(c, payload, 0xFF, 1, origin, tmin * 10.0f, -c.lightDir, tmax, RAY_SHADOW);
            payload.recDepth--;
            shadowColor = payload.hitValue;
            payload.rayType = RT_GENERIC;
        }
    } else {
        const float shadowModulation = std::pow(mat.transparency, 2.0f);
        shadowColor = mat.transparency < 1.f ? mix(vec3(1.0f, 1.0f, 1.0f), mat.diffuse * shadowModulation, mat.transparency) : vec3(0.4f, 0.4f, 0.4f);
    }

    // :207-221
    vec3 reflectColor(1.0f, 1.0f, 1.0f);
    float reflectDepth = 0.f;
    if(payload.recDepth < maxRecursions && mat.reflectivity > 0.f) {
        const vec3 dir = reflect(worldRayDirection, vnInWorldSpace);

        payload.recDepth += int(mat.rayConsumption);
        payload.refDepth += hitT;
        traceRay(c, payload, 0xff, 0, origin, tmin, dir, tmax, RAY_REFLECT);
        payload.recDepth -= int(mat.rayConsumption);

        reflectColor = payload.hitValue * mat.specular;
        reflectDepth = payload.depth;
    }

    // :224-251
    vec3 refractColor(1.0f, 1.0f, 1.0f);
    if(payload.recDepth < maxRecursions && mat.transparency > 0.f) {
        if(frontFacing) {
            const float eta = payload.curIOR / mat.ior;
            const vec3 dir = refract(worldRayDirection, vnInWorldSpace, eta);

            payload.recDepth++;
            payload.curIOR = mat.ior;
            traceRay(c, payload, 0xff, 0, origin, tmin, dir, tmax, RAY_REFRACT);
            payload.recDepth--;

            refractColor = payload.hitValue;
        } else {
            const float eta = mat.ior / 1.0f;
            const vec3 dir = refract(worldRayDirection, vnInWorldSpace, eta);

            payload.recDepth++;
            payload.curIOR = 1.0f;
            traceRay(c, payload, 0xff, 0, origin, tmin, dir, tmax, RAY_REFRACT);
            payload.recDepth--;

            const vec3 transmittanceModulation = mix(vec3(1.0f, 1.0f, 1.0f), mat.diffuse, std::log(1.0f + hitT));
            refractColor = transmittanceModulation * payload.hitValue;
        }
    }

    // :254-258
    baseColor *= shadowColor;
    baseColor += mat.emission * mat.diffuse;
    const float totalContrib = fmax_glsl(mat.transparency, mat.reflectivity);
    float weight = mat.reflectivity / (mat.transparency + mat.reflectivity);
    if(!(c.flags & ORC_STRICT_IEEE) && (mat.transparency + mat.reflectivity) == 0.0f) weight = 0.0f;  // hazard 8
    const vec3 roughCol = mix(refractColor, reflectColor, weight);
    payload.hitValue = mix(baseColor, roughCol, totalContrib);

    // :260-265
    if(payload.recDepth == 0) {
        payload.hitValue = baseColor;
        payload.normal = vnInWorldSpace;
        payload.roughValue = vec4(roughCol, fmin_glsl((reflectDepth / 50.f) * mat.roughness, mat.roughness / 2.1f));
        payload.reflectContribution = totalContrib;
    }

    // :267
    payload.depth = hitT;
}

// traceRayEXT: all call sites use gl_RayFlagsOpaqueEXT, SBT offset/stride 0.
void traceRay(Ctx& c, Payload& payload, uint32_t cullMask, int missIndex, vec3 origin, float tmin, vec3 dir, float tmax, int kind) {
    c.cnt.rays[kind]++;
    Hit hit;
    bool found = false;
    if(cullMask != 0) {
        found = c.scene->closestHit(origin, dir, tmin, tmax, hit, (c.flags & ORC_BRUTE_FORCE) != 0);
        if(dir.x == 0.0f && dir.y == 0.0f && dir.z == 0.0f) c.cnt.zeroDirRays++;
    }
    if(kind == RAY_PRIMARY && c.capture) {
        c.capInst = found ? hit.inst : 0xffffffffu; c.capPrim = found ? hit.prim : 0xffffffffu; c.capT = found ? hit.t : 0.0f;
    }
    if(found) closestHit(c, payload, origin, dir, hit);
    else if(missIndex == 0) miss0(c, payload, dir);
    else miss1(payload);
}

inline vec4 mulMat4(const float* m, vec4 v) {  // GLM operator*(mat4, vec4): (c0*x + c1*y) + (c2*z + c3*w)
    vec4 r;
    r.x = (m[0] * v.x + m[4] * v.y) + (m[8] * v.z + m[12] * v.w);
    r.y = (m[1] * v.x + m[5] * v.y) + (m[9] * v.z + m[13] * v.w);
    r.z = (m[2] * v.x + m[6] * v.y) + (m[10] * v.z + m[14] * v.w);
    r.w = (m[3] * v.x + m[7] * v.y) + (m[11] * v.z + m[15] * v.w);
    return r;
}

}  // namespace

// raygen.rgen:30-39 + raygen.h:69-115 for one pixel
void tracePixel(const Scene& scene, const Ubo& ubo, uint32_t W, uint32_t H, uint32_t px, uint32_t py, uint32_t flags, PixelOut& out,
                Counters& counters) {
    Ctx c{};
    c.scene = &scene; c.ubo = &ubo; c.flags = flags;
    c.lightDir = vec3(ubo.lightDir[0], ubo.lightDir[1], ubo.lightDir[2]);
    c.maxRecursions = ubo.maxRecursions;
    const int numSamples = ubo.numSamples;

    vec3 color(0.0f), normal(0.0f);
    vec4 roughValue(0.0f);
    float reflectContrib = 0, depth = 0;

    const vec4 origin = mulMat4(ubo.viewInverse, vec4(0, 0, 0, 1));
    const float tmin = 0.001f, tmax = 10000.0f;
    Payload payload{};

    for(int i = 0; i < numSamples; ++i) {
        const float* off = kAAOffsets[numSamples < 8 ? numSamples : 8][i % 8];
        const float pcx = (float)px + 0.5f + off[0], pcy = (float)py + 0.5f + off[1];
        const float uvx = pcx / (float)W, uvy = pcy / (float)H;
        const float dx = uvx * 2.0f - 1.0f, dy = uvy * 2.0f - 1.0f;

        const vec4 target = mulMat4(ubo.projInverse, vec4(dx, dy, 1, 1));
        const vec4 direction = mulMat4(ubo.viewInverse, vec4(normalize(target.xyz()), 0));

        payload.hitValue = vec3(0.0f);
        payload.normal = vec3(0.0f);
        payload.roughValue = vec4(0.0f);
        payload.depth = 0;
        payload.refDepth = 0;
        payload.curIOR = 1.0f;
        payload.rayType = RT_GENERIC;
        payload.recDepth = 0;
        payload.reflectContribution = 0;

        c.capture = (i == 0);
        traceRay(c, payload, 0xff, 0, origin.xyz(), tmin, direction.xyz(), tmax, RAY_PRIMARY);

        color += payload.hitValue;
        normal += payload.normal;
        roughValue += payload.roughValue;
        reflectContrib += payload.reflectContribution;
        depth += payload.depth;
    }
    const float n = (float)numSamples;
    out.base = vec4(color, reflectContrib) / n;
    out.normal = vec4(normal, std::log(depth) * 0.25f) / n;
    out.rough = roughValue / n;
    out.inst = c.capInst; out.prim = c.capPrim; out.t = c.capT;
    for(int k = 0; k < RAY_KINDS; ++k) counters.rays[k] += c.cnt.rays[k];
    counters.zeroDirRays += c.cnt.zeroDirRays;
}

}  // namespace orc
