// rg_trace.cuh -- parameters of the persistent traversal + shading kernel.
#pragma once
#include "rg_types.cuh"

namespace rg {

constexpr int kMaxRecursions = 8;   // BASELINE config 3 uses 8; the reference UI allows 0..7 (render_system.cpp:264)
constexpr int kMaxFrames = kMaxRecursions + 1;  // a generic hit entered at recDepth >= max still gets a frame
constexpr int kStackSize = 48;     // traversal stack entries (uint2) per ray: TLAS + BLAS + postponed primitive groups
#ifndef RG_POOL_CTX
#define RG_POOL_CTX 64
#endif
constexpr int kPoolCtx = RG_POOL_CTX;   // ray-tree contexts (pixel samples in progress) per warp
constexpr int kCtxQuads = kMaxFrames * 8 + 2;   // float4 per context in global memory: 8 per frame + the payload members only observable at recDepth 0
constexpr int kMaxPeers = 8;        // GPUs of one node
#ifndef RG_CHUNK_TILES
#define RG_CHUNK_TILES 16
#endif
constexpr uint32_t kChunkTiles = RG_CHUNK_TILES; // tiles per round-robin chunk in partitioned mode

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
    // trace domain: the pixels this launch is responsible for.  Single GPU / overdraw mode: the context's own rectangle.
    // Partitioned mode (world > 1): the full frame, of which this rank takes every world-th chunk of kChunkTiles 8x4 tiles.
    uint32_t dx0, dy0, dw, dh;
    uint32_t magicS, magicTilesX;   // floor(2^32 / numSamples), floor(2^32 / tiles per row of the domain) (0xffffffff for a divisor of 1): rg_trace.cu divMagic
    uint32_t rank, world;
    // where a finished pixel is stored: every target whose rectangle (region + halo) contains it.  Target `self` is this
    // context's own images; the others are peer GPUs' images, written over NVLink (peer pointers, same process or CUDA IPC).
    struct Target { uint2* base; uint2* normal; uint2* rough; int32_t x0, y0, w, h; };
    Target targets[kMaxPeers];
    uint32_t nTargets, self;
    int32_t sx0, sy0, sw, sh;   // own rectangle (for the id images)
    uint32_t* idInst; uint32_t* idPrim;        // optional primary ids, own rectangle only
    // One work item = one SAMPLE of one pixel (the heaviest pixels are glass with ~100 sequential rays per sample; splitting the
    // numSamples samples over lanes cuts that critical path).  Samples park their result in sampleScratch (3 x float4) and the lane
    // that finishes a pixel's last sample adds them up in the shader's order i = 0..S-1 (bit-identical to the sequential loop).
    float4* sampleScratch; uint32_t* sampleDone;
    // Heavy-first scheduling: tiles are handed out in the order of tileOrder (built from last frame's per-tile ray counts,
    // most expensive first) so the long ray trees start early and overlap with the bulk; tileCost collects this frame's counts.
    const uint32_t* tileOrder; uint32_t* tileCost;
    float4* ctxPool;   // tracePoolBytes(): frames of the per-warp context pools
    uint32_t* workCounter;
    unsigned long long* counters;  // 5 ray kinds + nodes, tris, instances
    uint32_t flags;
    float scatterColor[3];   // miss.rmiss:63-66 scatterColor: a function of lightDir alone, evaluated by the host once per frame
};

// pool: per-warp context pools with ray / hit queues (incoherent bounces); otherwise one context per lane (coherent scenes)
// seq (lanes kernel only): the numSamples samples of a pixel follow each other in one lane and are summed there (no sampleScratch /
// sampleDone traffic); otherwise one work item per sample, parked in sampleScratch and summed by the lane that finishes the last one
void launchTrace(const TraceParams& p, int numSms, bool pool, bool seq, cudaStream_t stream);
int traceLanesWarps(int numSms);   // resident warps of the lanes kernel's persistent grid
size_t tracePoolBytes(int numSms);
// number of tile slots of this rank's share of the trace domain (sizes tileOrder / tileCost / sampleDone / sampleScratch)
uint32_t traceShareTiles(uint32_t dw, uint32_t dh, uint32_t rank, uint32_t world);
// order[] = tile slots sorted by descending cost class (8 classes relative to the mean = rays counted by the trace kernel / slots); cost[]
// is cleared for the next frame.  work: 17 zeroed words of scratch (left zeroed).  Two launches.
void launchOrderTiles(uint32_t* cost, uint32_t nSlots, uint32_t* order, const unsigned long long* counters, uint32_t* work, int numSms, cudaStream_t stream);
// debug: n rays (8 floats each) -> closest hits
void launchTraceRays(const TraceParams& p, const float* rays8, uint32_t n, float* tuv, uint32_t* instPrim, cudaStream_t stream);

}  // namespace rg
