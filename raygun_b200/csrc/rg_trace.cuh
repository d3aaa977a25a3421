// rg_trace.cuh -- parameters of the persistent traversal + shading kernel.
#pragma once
#include "rg_types.cuh"

namespace rg {

constexpr int kMaxRecursions = 8;   // BASELINE config 3 uses 8; the reference UI allows 0..7 (render_system.cpp:264)
constexpr int kMaxFrames = kMaxRecursions + 1;  // a generic hit entered at recDepth >= max still gets a frame
constexpr int kStackSize = 40;      // traversal stack entries (uint2) per ray: TLAS + BLAS

struct TraceParams {
    const Node8* tlasNodes;
    const InstTrav* tlasLeaves;
    const Node8* blasNodes;
    const Tri* tris;
    const InstShade* instShade;
    const float4* vertices;   // 2 x float4 per Vertex
    const uint32_t* indices;
    const float4* materials;  // 4 x float4 per gpu::Material
    const float* ubo;         // 48 words (device)
    uint32_t nInst;
    uint32_t W, H;            // full frame (launch size for ray generation)
    uint32_t rx0, ry0, rw, rh;  // rectangle rendered by this context (region + halo, clipped), pitch = rw
    uint2* base; uint2* normal; uint2* rough;  // rgba16f images (4 halves = uint2)
    uint32_t* idInst; uint32_t* idPrim;        // optional primary ids (same pitch)
    uint32_t* workCounter;
    unsigned long long* counters;  // 5 ray kinds + nodes, tris, instances
    uint32_t flags;
};

void launchTrace(const TraceParams& p, int numSms, cudaStream_t stream);
// debug: n rays (8 floats each) -> closest hits
void launchTraceRays(const TraceParams& p, const float* rays8, uint32_t n, float* tuv, uint32_t* instPrim, cudaStream_t stream);

}  // namespace rg
