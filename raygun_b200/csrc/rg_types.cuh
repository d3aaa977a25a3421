// rg_types.cuh -- device data layout shared by the builder, the traversal kernel and the C ABI.
//
// Layout in HBM (all arrays are plain cudaMalloc allocations owned by rg_ctx):
//   vertices   n_vtx  x 32 B   the reference's Vertex records, unchanged          (vertex.def:3-7)
//   indices    n_idx  x  4 B   u32, mesh-local                                     (render_system.cpp:270-305)
//   materials  n_mat  x 64 B   gpu::Material records, unchanged                    (gpu_material.def:11-26)
//   nodes      Node8  x 80 B   compressed 8-wide BVH nodes, 5 x 128-bit loads each; all BLASes back to back,
//                              the per-frame TLAS in its own array
//   tris       Tri    x 48 B   3 x float4: vertex positions re-laid out in leaf order, w0 = primitive id
//   tlasLeaves InstTrav x 80 B world->object 3x4 + BLAS root + the mesh's bounding sphere, gathered in TLAS leaf order each frame
//   instShade  InstShade x 64 B object->world 3x4 + the offset-table entry, indexed by instance id
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace rg {

// Compressed 8-wide node (after Ylitie, Karras, Laine 2017).  Child boxes are quantised to 8 bits
// relative to the node origin p with per-axis power-of-two scale 2^e.  The n children of a node own
// NIBBLES 0..n-1 of codes / vm (leaves first, then internal children); the plane bytes of nibble m sit at byte
// POSITION posOfNibble(m) of the six plane arrays (empty positions hold an inverted box), so that the traversal
// tests the children two at a time (positions 2i, 2i + 1 = nibbles i, i + 4).  Every internal child also has
// an octant CODE c (bit 2/1/0 = +x/+y/+z side of the node centre where possible): XOR-ing the code with the
// ray's direction octant yields a front-to-back order without sorting.
//   codes  nibble m = code of the internal child at position j (8 for a leaf / empty position), m = nibbleOfPos(j)
//   vm     nibble m = bit 3: internal child; bits 0..2: valid primitives of the leaf at position j
//          (primitive k of that leaf is element (primBase & 0x7fffffff) + 3 m + k of the primitive array: stride 3,
//          unused elements are never read)
//   imask  bit c = an internal child with code c exists; child nodes are stored contiguously from childBase
//          in code order
//   primBase carries bit 31 (marks primitive groups on the traversal stack)
struct alignas(16) Node8 {
    float px, py, pz;
    uint8_t ex, ey, ez, imask;
    uint32_t childBase;
    uint32_t primBase;
    uint32_t codes;
    uint32_t vm;
    uint8_t qlox[8], qloy[8];
    uint8_t qloz[8], qhix[8];
    uint8_t qhiy[8], qhiz[8];
};
static_assert(sizeof(Node8) == 80, "Node8 must be 5 x 16 bytes");

struct alignas(16) Tri {  // 48 B
    float v0x, v0y, v0z; uint32_t prim;
    float v1x, v1y, v1z; uint32_t pad1;
    float v2x, v2y, v2z; uint32_t pad2;
};
static_assert(sizeof(Tri) == 48, "Tri must be 3 x 16 bytes");

struct alignas(16) InstTrav {  // 80 B = 5 x 128-bit loads, all issued at once when an instance is reached
    float w2o[12];       // world -> object, 3x4 row-major
    uint32_t blasRoot;   // absolute index of the BLAS root node; 0xffffffff = empty mesh
    uint32_t instId;     // gl_InstanceCustomIndexEXT
    uint32_t pad0;       // 1: pure translation (the ray direction is kept)
    uint32_t pad1;       // mesh index
    float sphere[4];     // the mesh's bounding sphere in object space: centre, r^2 (FLT_MAX: no test) -- see k_mesh_sphere_store
};
static_assert(sizeof(InstTrav) == 80, "InstTrav");

struct alignas(16) InstShade {  // 64 B
    float o2w[12];  // object -> world, 3x4 row-major
    uint32_t vtxOff, idxOff, matOff, mesh;
};
static_assert(sizeof(InstShade) == 64, "InstShade");

struct Aabb { float lo[3], hi[3]; };

// binary LBVH scratch node (builder only)
struct BNode {
    float lo[3]; uint32_t left;   // child index; bit 31 set = leaf (sorted position)
    float hi[3]; uint32_t right;
};
static_assert(sizeof(BNode) == 32, "BNode");

constexpr uint32_t kLeafBit = 0x80000000u;
constexpr uint32_t kInvalid = 0xffffffffu;
// Primitives per leaf child (<= 3: a leaf child owns three bits of its vm nibble).  Measured with the binary16 node test (C2 / C3 / C4
// trace ms): 3 triangles and 3 instances 3.55 / 41.1 / 23.1; 2 and 2: 3.51 / 40.0 / 21.3; 1 and 1: 3.59 / 38.4 / 20.4 -- entering an
// instance (ray transform, set-up, BLAS root) costs far more than one more box in a TLAS node, a triangle test about as much as a
// third of a node step.
#ifndef RG_MAX_LEAF_TRIS
#define RG_MAX_LEAF_TRIS 2
#endif
#ifndef RG_MAX_LEAF_INSTS
#define RG_MAX_LEAF_INSTS 1
#endif
constexpr int kMaxLeafTris = RG_MAX_LEAF_TRIS, kMaxLeafInsts = RG_MAX_LEAF_INSTS;
constexpr uint32_t kPrimGroupBit = 0x80000000u;   // Node8::primBase / traversal stack entries
// RG_HALF_SLAB (default): the traversal tests the children of a node TWO AT A TIME in packed binary16 arithmetic (rg_trace.cu pairTest):
// children at positions 2i and 2i + 1 share every instruction, and their results land in the low / high half of one mask, so the nibbles of
// codes / vm / the primitive group are dealt out as nibbleOfPos below.  0: one child at a time in binary32 (round 1 / 2 formulation).
#ifndef RG_HALF_SLAB
#define RG_HALF_SLAB 1
#endif
// RG_PLANE_DIFF (with RG_HALF_SLAB): the "high" plane words of a node hold hi - lo as 32-bit integers (mod 2^32), so that the traversal
// picks the near / far planes of an axis by the ray's sign s in {0, 1} as lo + s * diff and lo + (1 - s) * diff: integer multiply-adds
// on the FMA pipe instead of bit selects on the ALU pipe, which is the one the node step saturates.
#ifndef RG_PLANE_DIFF
#define RG_PLANE_DIFF RG_HALF_SLAB
#endif
constexpr uint32_t kExpBias = RG_HALF_SLAB ? 0 : 15;   // Node8::ex/ey/ez hold e + 127 + kExpBias: the plane of byte q is p + q * 2^e
// nibble of codes / vm (and the triple of primitive-array elements) that belongs to the child whose plane bytes sit at position j
__host__ __device__ constexpr int nibbleOfPos(int j) { return RG_HALF_SLAB ? (j >> 1) + 4 * (j & 1) : j; }
__host__ __device__ constexpr int posOfNibble(int m) { return RG_HALF_SLAB ? ((m & 3) << 1) | (m >> 2) : m; }   // the inverse
static_assert(posOfNibble(nibbleOfPos(0)) == 0 && posOfNibble(nibbleOfPos(1)) == 1 && posOfNibble(nibbleOfPos(2)) == 2 && posOfNibble(nibbleOfPos(3)) == 3 &&
              posOfNibble(nibbleOfPos(4)) == 4 && posOfNibble(nibbleOfPos(5)) == 5 && posOfNibble(nibbleOfPos(6)) == 6 && posOfNibble(nibbleOfPos(7)) == 7,
              "posOfNibble must invert nibbleOfPos");
static_assert(!RG_HALF_SLAB || (nibbleOfPos(0) == 0 && nibbleOfPos(1) == 4 && nibbleOfPos(2) == 1 && nibbleOfPos(3) == 5),
              "pair i = positions 2i, 2i + 1 = nibbles i, i + 4 (the low / high half of pairTest's mask)");
constexpr uint32_t kLeafStride = 3;               // primitive array elements reserved per leaf child

struct Hit { float t, u, v; uint32_t inst, prim; };

}  // namespace rg
