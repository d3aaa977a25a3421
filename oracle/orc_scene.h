// TEST INFRASTRUCTURE -- CPU oracle (see orc_math.h header).
//
// orc_scene.h: the scene exactly as the reference hands it to the GPU (SURVEY.md 8a
// rows a1-a4) plus the oracle's own closest-hit search.  The reference delegates the
// search to the Vulkan driver (raygun/render/acceleration_structure.cpp:134,193;
// traceRayEXT in resources/shaders/*.rchit/*.rgen), so the oracle states its SEMANTICS:
//   * two-level: every instance's ray is the world ray mapped by worldToObject; t is
//     preserved (Vulkan instance semantics, acceleration_structure.cpp:34-52);
//   * no culling (eTriangleCullDisable :41), opaque, closest hit, tmin < t < tmax;
//   * watertight ray/triangle test (Woop, Benthin, Wald 2013) so shared edges never leak;
//   * equal-t ties resolve to the smallest (instance, primitive) -- the Vulkan spec
//     leaves ties open; this makes the answer independent of traversal order.
// The binary BVH here is only an accelerator for that definition; `brute` walks every
// triangle and must give identical answers (tests/test_oracle.py checks it).
#pragma once
#include <cstdint>
#include <vector>

#include "orc_math.h"

namespace orc {

// resources/shaders/vertex.def:3-7 (32 bytes)
struct Vertex { float px, py, pz; uint32_t matIndex; float nx, ny, nz; float pad1; };
static_assert(sizeof(Vertex) == 32, "Vertex layout");
// resources/shaders/gpu_material.def:11-26 (64 bytes)
struct Material {
    float diffuse[3]; float transparency; float specular[3]; float reflectivity;
    float roughness; float ior; uint32_t effectId; uint32_t rayConsumption;
    float emission; float pad0, pad1, pad2;
};
static_assert(sizeof(Material) == 64, "Material layout");
// resources/shaders/uniform_buffer_object.def:3-17 (192 bytes, GLSL std140 view)
struct Ubo {
    float viewInverse[16]; float projInverse[16];  // column-major
    float clearColor[3]; int32_t numSamples;
    float lightDir[3]; int32_t maxRecursions;
    float time; uint32_t showAlpha; float pad0, pad1;
    float fadeColor[4];
};
static_assert(sizeof(Ubo) == 192, "UBO layout");

struct MeshRange { uint32_t vtxOff, vtxCnt, idxOff, idxCnt; };
struct Instance {
    float m[12];    // 3x4 row-major object->world (acceleration_structure.cpp:44-45)
    float inv[12];  // world->object
    uint32_t mesh, vtxOff, idxOff, matOff;  // offset table entry (instance_offset_table.def)
};

struct Hit { float t, u, v; uint32_t inst, prim; };

struct Bvh2 {
    struct Node { float lo[3], hi[3]; uint32_t left, count; };  // count>0: leaf, left = first item
    std::vector<Node> nodes;
    std::vector<uint32_t> items;
};

struct Scene {
    std::vector<Vertex> vertices;
    std::vector<uint32_t> indices;
    std::vector<MeshRange> meshes;
    std::vector<Material> materials;
    std::vector<Instance> instances;
    std::vector<Bvh2> blas;  // one per mesh, object space
    Bvh2 tlas;               // over instance world boxes
    void build();
    // closest hit; returns false on miss.  brute=true ignores both BVHs.
    bool closestHit(vec3 org, vec3 dir, float tmin, float tmax, Hit& hit, bool brute) const;
};

}  // namespace orc
