// TEST INFRASTRUCTURE -- CPU oracle (see orc_math.h header).
//
// orc_api.cpp: C entry points (ctypes) + the oracle's definition of the LBVH sort keys.
//
// PARITY STATUS: the reference holds no tests, golden images or known-answer vectors for
// this path and its ray traversal lives in the Vulkan driver, so
// the ray-tracing part of this oracle is "parity unpinned" by the reference's own tests.
// What IS pinned: the shading / post-pass arithmetic against the reference's own shader
// sources compiled for the CPU (oracle/_ref, see oracle/ref_recipe/), camera / transform
// math against the reference's vendored GLM (tests/golden/glm_golden.json), asset ingestion
// against SURVEY.md Appendix C known answers, and binary16 conversion against numpy.
#include <omp.h>

#include <algorithm>
#include <chrono>
#include <cstring>
#include <numeric>
#include <vector>

#include "orc_render.h"

using namespace orc;

namespace {

inline uint32_t expandBits10(uint32_t v) {  // 10 bits -> every third bit
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

// The sort key both the oracle and the CUDA builder must produce bit for bit
// (north_star: "Morton/sort output must be bit-exact").  All operations are single IEEE
// binary32 operations in the order written:
//   c    = (boxLo + boxHi) * 0.5                      per axis, box = AABB of the primitive
//   s    = ext > 0 ? 1024 / ext : 0                   ext = sceneHi - sceneLo
//   q    = uint(min(max((c - sceneLo) * s, 0), 1023)) truncation
//   code = expand(qx) << 2 | expand(qy) << 1 | expand(qz)   (30 bits)
inline uint32_t mortonOf(const float* blo, const float* bhi, const float* slo, const float* shi) {
    uint32_t q[3];
    for(int a = 0; a < 3; ++a) {
        const float c = (blo[a] + bhi[a]) * 0.5f;
        const float ext = shi[a] - slo[a];
        const float s = ext > 0.0f ? 1024.0f / ext : 0.0f;
        float v = (c - slo[a]) * s;
        v = v > 0.0f ? v : 0.0f;
        v = v < 1023.0f ? v : 1023.0f;
        q[a] = (uint32_t)v;
    }
    return (expandBits10(q[0]) << 2) | (expandBits10(q[1]) << 1) | expandBits10(q[2]);
}

}  // namespace

extern "C" {

struct orc_scene { Scene s; };

orc_scene* orc_scene_create(const void* vertices, uint32_t nVtx, const uint32_t* indices, uint32_t nIdx, const uint32_t* meshes,
                            uint32_t nMeshes, const void* materials, uint32_t nMat, const float* instXform, const uint32_t* instMeta,
                            uint32_t nInst) {
    auto* h = new orc_scene();
    Scene& s = h->s;
    s.vertices.assign((const Vertex*)vertices, (const Vertex*)vertices + nVtx);
    s.indices.assign(indices, indices + nIdx);
    s.meshes.resize(nMeshes);
    for(uint32_t m = 0; m < nMeshes; ++m) s.meshes[m] = MeshRange{meshes[4 * m], meshes[4 * m + 1], meshes[4 * m + 2], meshes[4 * m + 3]};
    s.materials.assign((const Material*)materials, (const Material*)materials + nMat);
    s.instances.resize(nInst);
    for(uint32_t i = 0; i < nInst; ++i) {
        std::memcpy(s.instances[i].m, instXform + 12 * i, 48);
        s.instances[i].mesh = instMeta[4 * i]; s.instances[i].vtxOff = instMeta[4 * i + 1];
        s.instances[i].idxOff = instMeta[4 * i + 2]; s.instances[i].matOff = instMeta[4 * i + 3];
    }
    s.build();
    return h;
}

void orc_scene_destroy(orc_scene* h) { delete h; }

// Replace the instance list (per-frame TLAS input, raytracer.cpp:76-85).
void orc_scene_set_instances(orc_scene* h, const float* instXform, const uint32_t* instMeta, uint32_t nInst) {
    Scene& s = h->s;
    s.instances.resize(nInst);
    for(uint32_t i = 0; i < nInst; ++i) {
        std::memcpy(s.instances[i].m, instXform + 12 * i, 48);
        s.instances[i].mesh = instMeta[4 * i]; s.instances[i].vtxOff = instMeta[4 * i + 1];
        s.instances[i].idxOff = instMeta[4 * i + 2]; s.instances[i].matOff = instMeta[4 * i + 3];
    }
    s.build();
}

// Trace only (raygen + hit/miss recursion): fills base/normal/rough + primary ids.
// counters: 5 ray kinds + zero-direction count (6 x u64).  Returns wall seconds.
double orc_trace(orc_scene* h, const void* ubo192, uint32_t W, uint32_t H, uint32_t flags, int threads, uint32_t y0, uint32_t y1,
                 void* base, void* normal, void* rough, uint32_t* instId, uint32_t* primId, float* hitT, uint64_t* counters) {
    Ubo ubo; std::memcpy(&ubo, ubo192, 192);
    if(threads <= 0) threads = omp_get_max_threads();
    if(y1 > H) y1 = H;
    std::vector<Counters> cnt((size_t)threads);
    const auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
    for(int y = (int)y0; y < (int)y1; ++y) {
        Counters& c = cnt[(size_t)omp_get_thread_num()];
        for(uint32_t x = 0; x < W; ++x) {
            PixelOut o;
            tracePixel(h->s, ubo, W, H, x, (uint32_t)y, flags, o, c);
            const size_t i = (size_t)y * W + x;
            ((half4*)base)[i] = pack_half4(o.base);
            ((half4*)normal)[i] = pack_half4(o.normal);
            ((half4*)rough)[i] = pack_half4(o.rough);
            if(instId) instId[i] = o.inst;
            if(primId) primId[i] = o.prim;
            if(hitT) hitT[i] = o.t;
        }
    }
    const auto t1 = std::chrono::steady_clock::now();
    if(counters) {
        for(int k = 0; k < 6; ++k) counters[k] = 0;
        for(auto& c: cnt) { for(int k = 0; k < RAY_KINDS; ++k) counters[k] += c.rays[k]; counters[5] += c.zeroDirRays; }
    }
    return std::chrono::duration<double>(t1 - t0).count();
}

// Post chain on caller-owned images (all W*H): base/normal/rough are inputs (and are modified
// exactly as the reference modifies them: the FXAA swap, showAlpha).  Returns wall seconds.
double orc_post(const void* ubo192, uint32_t W, uint32_t H, uint32_t flags, int threads, void* base, void* normal, void* rough,
                void* final_, void* roughA, void* roughB, int8_t* transitions, uint8_t* rgba8) {
    Ubo ubo; std::memcpy(&ubo, ubo192, 192);
    Frame f{W, H, (half4*)base, (half4*)normal, (half4*)rough, (half4*)final_, (half4*)roughA, (half4*)roughB, transitions, rgba8};
    const auto t0 = std::chrono::steady_clock::now();
    postChain(ubo, f, flags, threads);
    const auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

// Single closest-hit query (for direct traversal tests).  Returns 1 on hit.
int orc_closest_hit(orc_scene* h, const float* org, const float* dir, float tmin, float tmax, int brute, float* tuv, uint32_t* instPrim) {
    Hit hit;
    const bool f = h->s.closestHit(vec3(org[0], org[1], org[2]), vec3(dir[0], dir[1], dir[2]), tmin, tmax, hit, brute != 0);
    tuv[0] = hit.t; tuv[1] = hit.u; tuv[2] = hit.v; instPrim[0] = hit.inst; instPrim[1] = hit.prim;
    return f ? 1 : 0;
}

// Morton keys of the triangles of one mesh + the stable order (std::stable_sort by key).
// vertices: 32-byte records; idx: mesh-local indices (3 per triangle) relative to vtxOff.
void orc_morton_triangles(const void* vertices, uint32_t vtxOff, const uint32_t* idx, uint32_t nTri, uint32_t* codes, uint32_t* order,
                          float* sceneBox /*6*/) {
    const Vertex* V = (const Vertex*)vertices + vtxOff;
    float slo[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f}, shi[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
    std::vector<float> lo((size_t)nTri * 3), hi((size_t)nTri * 3);
    for(uint32_t p = 0; p < nTri; ++p) {
        for(int a = 0; a < 3; ++a) { lo[3 * p + a] = 3.402823466e+38f; hi[3 * p + a] = -3.402823466e+38f; }
        for(int k = 0; k < 3; ++k) {
            const float* pos = &V[idx[3 * p + k]].px;
            for(int a = 0; a < 3; ++a) { lo[3 * p + a] = std::min(lo[3 * p + a], pos[a]); hi[3 * p + a] = std::max(hi[3 * p + a], pos[a]); }
        }
        for(int a = 0; a < 3; ++a) { slo[a] = std::min(slo[a], lo[3 * p + a]); shi[a] = std::max(shi[a], hi[3 * p + a]); }
    }
    for(uint32_t p = 0; p < nTri; ++p) codes[p] = mortonOf(&lo[3 * p], &hi[3 * p], slo, shi);
    std::iota(order, order + nTri, 0u);
    std::stable_sort(order, order + nTri, [&](uint32_t a, uint32_t b) { return codes[a] < codes[b]; });
    if(sceneBox) { for(int a = 0; a < 3; ++a) { sceneBox[a] = slo[a]; sceneBox[3 + a] = shi[a]; } }
}

// Morton keys + stable order for arbitrary boxes (TLAS: instance world boxes). boxes: nBox x 6 (lo, hi).
void orc_morton_boxes(const float* boxes, uint32_t nBox, uint32_t* codes, uint32_t* order) {
    float slo[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f}, shi[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
    for(uint32_t p = 0; p < nBox; ++p)
        for(int a = 0; a < 3; ++a) { slo[a] = std::min(slo[a], boxes[6 * p + a]); shi[a] = std::max(shi[a], boxes[6 * p + 3 + a]); }
    for(uint32_t p = 0; p < nBox; ++p) codes[p] = mortonOf(&boxes[6 * p], &boxes[6 * p + 3], slo, shi);
    std::iota(order, order + nBox, 0u);
    std::stable_sort(order, order + nBox, [&](uint32_t a, uint32_t b) { return codes[a] < codes[b]; });
}

// World AABB of an instance as the builder must compute it: the 8 corners of the mesh's
// object-space AABB mapped by the 3x4 (each coordinate ((m0*x + m1*y) + m2*z) + m3, single
// IEEE operations), then min/max.  out: 6 floats.
void orc_instance_world_box(const float* m, const float* meshBox /*lo,hi*/, float* out) {
    for(int a = 0; a < 3; ++a) { out[a] = 3.402823466e+38f; out[3 + a] = -3.402823466e+38f; }
    for(int c = 0; c < 8; ++c) {
        const float x = (c & 1) ? meshBox[3] : meshBox[0], y = (c & 2) ? meshBox[4] : meshBox[1], z = (c & 4) ? meshBox[5] : meshBox[2];
        for(int r = 0; r < 3; ++r) {
            const float w = ((m[4 * r] * x + m[4 * r + 1] * y) + m[4 * r + 2] * z) + m[4 * r + 3];
            out[r] = std::min(out[r], w); out[3 + r] = std::max(out[3 + r], w);
        }
    }
}

void orc_f32_to_f16(const float* in, uint16_t* out, uint64_t n) { for(uint64_t i = 0; i < n; ++i) out[i] = f32_to_f16(in[i]); }
void orc_f16_to_f32(const uint16_t* in, float* out, uint64_t n) { for(uint64_t i = 0; i < n; ++i) out[i] = f16_to_f32(in[i]); }
int orc_max_threads() { return omp_get_max_threads(); }

}  // extern "C"
