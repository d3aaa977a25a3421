// rg_build.cuh -- LBVH builder interface (device-side acceleration structures).
// Replaces the driver-side vkCmdBuildAccelerationStructuresKHR calls recorded by
// raygun/render/acceleration_structure.cpp:134 (TLAS, every frame) and :193 (BLAS, once per mesh).
#pragma once
#include "rg_types.cuh"

struct rg_instance;

namespace rg {

// Scratch + result of one LBVH build over n primitives (triangles of a mesh, or instances).
struct LbvhScratch {
    uint32_t capacity = 0;      // primitives
    Aabb* primBox = nullptr;    // [n]  per-primitive box (original order)
    uint32_t* keys[2] = {nullptr, nullptr};   // [n] Morton keys (ping-pong)
    uint32_t* vals[2] = {nullptr, nullptr};   // [n] primitive ids (ping-pong)
    uint32_t* hist = nullptr;   // [256 * sortBlocks]
    BNode* bnodes = nullptr;    // [n-1] binary internal nodes
    uint2* range = nullptr;     // [n-1] first,last sorted position covered by each internal node
    uint32_t* parent = nullptr; // [2n-1] parent of internal i at [i], of leaf s at [n-1+s]
    uint32_t* flags = nullptr;  // [n-1] refit arrival counters
    uint2* queue[2] = {nullptr, nullptr};     // [n] collapse work items {binary ref, wide index}
    uint32_t* counters = nullptr;             // [8]: 0,1 queue counts, 2 node counter, 3 prim counter
    int32_t* sceneBox = nullptr;              // [6] order-preserving int encoding of the scene box
    uint32_t* wideRef = nullptr;              // [8 * maxWideNodes] binary ref per wide slot (refit)
    uint32_t* gridSync = nullptr;             // [128] k_collapse_grid: barrier, error flag, items per level (from word 16)
    uint32_t sortedBuf = 0;                   // which of keys[]/vals[] holds the sorted result
    uint64_t launches = 0;
    void reserve(uint32_t n);
    void release();
};

constexpr uint32_t kTlasFusedMax = 1024;             // instances: the whole TLAS build in one block / one launch

struct TriSource { const void* vertices; const uint32_t* indices; uint32_t vtxOff, idxOff, nTri; };

// Builds the BLAS of one mesh.  nodesOut / trisOut point at the first free slot of the shared arrays;
// returns the number of wide nodes / Tri records written through nNodes / nTris (host values; synchronous).
// rootBoxOut (device, 6 floats) receives the mesh's object-space box, sphereOut (device) its bounding sphere (xyz = centre, w = r^2, or
// FLT_MAX where the sphere cannot reject anything the box lets through).
void buildBlas(LbvhScratch& s, const TriSource& src, Node8* nodesBase, uint32_t nodeOffset, Tri* trisBase, uint32_t triOffset,
               uint32_t* nNodes, uint32_t* nTris, float* rootBoxOut, float4* sphereOut, cudaStream_t stream);

// Refit: vertices changed, topology kept (needs the scratch of the original build, kept per mesh).
void refitBlas(LbvhScratch& s, const TriSource& src, Node8* nodesBase, uint32_t nodeOffset, uint32_t nNodes, Tri* trisBase, uint32_t triOffset,
               float* rootBoxOut, float4* sphereOut, cudaStream_t stream);

// Per-frame TLAS from the raw instance records: instance preparation (world->object, offset table), world boxes from
// meshBoxes (device, 6 floats per mesh), LBVH, collapse.  instTrav / instShade are indexed by instance id; tlasLeavesOut
// receives the InstTrav records in leaf order.  One launch for n <= kTlasFusedMax; fully asynchronous (no host round trip) at every size.
void buildTlas(LbvhScratch& s, const rg_instance* raw, uint32_t nInst, const uint32_t* meshRoots, uint32_t nMeshes, InstTrav* instTrav,
               InstShade* instShade, const float* meshBoxes, const float4* meshSpheres, Node8* tlasNodes, InstTrav* tlasLeavesOut, cudaStream_t stream);

}  // namespace rg
