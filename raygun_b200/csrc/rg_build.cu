// rg_build.cu -- LBVH builder for sm_100a: primitive boxes -> 30-bit Morton keys -> LSD radix sort
// (8-bit digits, stable) -> Karras 2012 hierarchy -> bottom-up refit -> collapse to compressed 8-wide nodes.
//
// Replaces the driver-side acceleration-structure builds of the reference
// (raygun/render/acceleration_structure.cpp:134 TLAS per frame, :193 BLAS per mesh).
// The sort key is contractually bit-exact with oracle/orc_api.cpp (mortonOf): every operation below is an
// explicitly rounded single binary32 operation (__fadd_rn & co are never contracted into FMAs).
#include "rg_build.cuh"

#include "../../include/rgb200.h"

#include <cfloat>
#include <cstdio>

namespace rg {

#define RG_CUDA_OK(x) do { cudaError_t e_ = (x); if(e_ != cudaSuccess) { fprintf(stderr, "rgb200: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); } } while(0)

namespace {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 8;
constexpr int kSortTile = kSortThreads * kSortItems;

__device__ __forceinline__ int encodeFloat(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float decodeFloat(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__device__ __forceinline__ uint32_t expandBits10(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__global__ void k_init_build(int32_t* sceneBox, uint32_t* counters) {
    if(threadIdx.x < 3) sceneBox[threadIdx.x] = 0x7fffffff;
    else if(threadIdx.x < 6) sceneBox[threadIdx.x] = (int)0x80000000;
    if(threadIdx.x < 8) counters[threadIdx.x] = 0;
}

__device__ __forceinline__ void reduceSceneBox(const float lo[3], const float hi[3], bool valid, int32_t* sceneBox) {
    // warp reduce, then one atomic per warp and axis
#pragma unroll
    for(int a = 0; a < 3; ++a) {
        int l = valid ? encodeFloat(lo[a]) : 0x7fffffff, h = valid ? encodeFloat(hi[a]) : (int)0x80000000;
#pragma unroll
        for(int o = 16; o; o >>= 1) { l = min(l, __shfl_xor_sync(0xffffffffu, l, o)); h = max(h, __shfl_xor_sync(0xffffffffu, h, o)); }
        if((threadIdx.x & 31) == 0) { atomicMin(&sceneBox[a], l); atomicMax(&sceneBox[3 + a], h); }
    }
}

// Per-triangle box (min / max of the three vertex positions; exact) + scene box.
__global__ void k_tri_boxes(const float4* __restrict__ vertices /*2 float4 per vertex*/, const uint32_t* __restrict__ indices, uint32_t vtxOff,
                            uint32_t idxOff, uint32_t nTri, Aabb* __restrict__ primBox, int32_t* sceneBox) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    const bool valid = p < nTri;
    if(valid) {
#pragma unroll
        for(int k = 0; k < 3; ++k) {
            const uint32_t vi = vtxOff + indices[idxOff + 3 * p + k];
            const float4 v = vertices[2 * (size_t)vi];
            lo[0] = fminf(lo[0], v.x); lo[1] = fminf(lo[1], v.y); lo[2] = fminf(lo[2], v.z);
            hi[0] = fmaxf(hi[0], v.x); hi[1] = fmaxf(hi[1], v.y); hi[2] = fmaxf(hi[2], v.z);
        }
        Aabb b; for(int a = 0; a < 3; ++a) { b.lo[a] = lo[a]; b.hi[a] = hi[a]; }
        primBox[p] = b;
    }
    reduceSceneBox(lo, hi, valid, sceneBox);
}

// Per-instance world box: the 8 corners of the mesh box mapped by the 3x4, each coordinate
// ((m0*x + m1*y) + m2*z) + m3 in single rounded operations (oracle: orc_instance_world_box).
__device__ __forceinline__ void d_inst_boxes(uint32_t i, const InstShade* __restrict__ inst, const float* __restrict__ meshBoxes, uint32_t nInst,
                                             Aabb* __restrict__ primBox, int32_t* sceneBox) {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    const bool valid = i < nInst;
    if(valid) {
        const InstShade in = inst[i];
        const float* mb = meshBoxes + 6 * in.mesh;
        const bool empty = mb[0] > mb[3];
        if(empty) {  // empty mesh: degenerate box at the instance origin, never entered (blasRoot invalid)
            for(int a = 0; a < 3; ++a) lo[a] = hi[a] = in.o2w[4 * a + 3];
        } else {
            for(int c = 0; c < 8; ++c) {
                const float x = (c & 1) ? mb[3] : mb[0], y = (c & 2) ? mb[4] : mb[1], z = (c & 4) ? mb[5] : mb[2];
#pragma unroll
                for(int r = 0; r < 3; ++r) {
                    const float w = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(in.o2w[4 * r], x), __fmul_rn(in.o2w[4 * r + 1], y)), __fmul_rn(in.o2w[4 * r + 2], z)),
                                              in.o2w[4 * r + 3]);
                    lo[r] = fminf(lo[r], w); hi[r] = fmaxf(hi[r], w);
                }
            }
        }
        Aabb b; for(int a = 0; a < 3; ++a) { b.lo[a] = lo[a]; b.hi[a] = hi[a]; }
        primBox[i] = b;
    }
    reduceSceneBox(lo, hi, valid, sceneBox);
}
__global__ void k_inst_boxes(const InstShade* __restrict__ inst, const float* __restrict__ meshBoxes, uint32_t nInst, Aabb* __restrict__ primBox,
                             int32_t* sceneBox) {
    d_inst_boxes(blockIdx.x * blockDim.x + threadIdx.x, inst, meshBoxes, nInst, primBox, sceneBox);
}

// The world box of an instance, cut down to the box of its mesh's bounding sphere mapped by the 3x4 (an ellipsoid: half extent r |row| per
// axis around the mapped centre).  The corner box of a rotated mesh box is up to sqrt(2) (sqrt(3)) times too wide per axis; for round
// meshes the sphere box is exact under any rotation.  Runs AFTER the Morton keys were taken from the corner boxes (their order stays
// bit-identical to the oracle's) and before the refit, so only the boxes the TLAS nodes are quantised from shrink.  Conservative: radius
// and centre carry their rounding errors outwards, and the result is the intersection of two boxes that both contain the mesh.
__device__ __forceinline__ void d_tighten_inst_box(uint32_t i, const InstShade* __restrict__ inst, const float4* __restrict__ meshSpheres, uint32_t nMeshes,
                                                   Aabb* __restrict__ primBox) {
    const InstShade in = inst[i];
    if(!meshSpheres || in.mesh >= nMeshes) return;
    const float4 sph = meshSpheres[in.mesh];
    if(!(sph.w < 3.0e38f)) return;   // no sphere for this mesh (k_mesh_sphere_store)
    const float r = sqrtf(sph.w) * 1.000002f;
    Aabb b = primBox[i];
#pragma unroll
    for(int a = 0; a < 3; ++a) {
        const float m0 = in.o2w[4 * a], m1 = in.o2w[4 * a + 1], m2 = in.o2w[4 * a + 2], m3 = in.o2w[4 * a + 3];
        const float wc = fmaf(m0, sph.x, fmaf(m1, sph.y, fmaf(m2, sph.z, m3)));
        const float len = sqrtf(fmaf(m0, m0, fmaf(m1, m1, m2 * m2)));
        const float e = r * len * 1.000002f + 4.0e-7f * (fabsf(m0 * sph.x) + fabsf(m1 * sph.y) + fabsf(m2 * sph.z) + fabsf(m3));
        b.lo[a] = fmaxf(b.lo[a], wc - e); b.hi[a] = fminf(b.hi[a], wc + e);
    }
    primBox[i] = b;
}
__global__ void k_tighten_inst_boxes(const InstShade* __restrict__ inst, const float4* __restrict__ meshSpheres, uint32_t nMeshes, uint32_t nInst,
                                     Aabb* __restrict__ primBox) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < nInst) d_tighten_inst_box(i, inst, meshSpheres, nMeshes, primBox);
}

__device__ __forceinline__ void d_morton(uint32_t p, const Aabb* __restrict__ primBox, uint32_t n, const int32_t* __restrict__ sceneBox,
                                         uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    if(p >= n) return;
    const Aabb b = primBox[p];
    uint32_t q[3];
#pragma unroll
    for(int a = 0; a < 3; ++a) {
        const float slo = decodeFloat(sceneBox[a]), shi = decodeFloat(sceneBox[3 + a]);
        const float c = __fmul_rn(__fadd_rn(b.lo[a], b.hi[a]), 0.5f);
        const float ext = __fsub_rn(shi, slo);
        const float s = ext > 0.0f ? __fdiv_rn(1024.0f, ext) : 0.0f;
        float v = __fmul_rn(__fsub_rn(c, slo), s);
        v = v > 0.0f ? v : 0.0f;
        v = v < 1023.0f ? v : 1023.0f;
        q[a] = __float2uint_rz(v);
    }
    keys[p] = (expandBits10(q[0]) << 2) | (expandBits10(q[1]) << 1) | expandBits10(q[2]);
    vals[p] = p;
}
__global__ void k_morton(const Aabb* __restrict__ primBox, uint32_t n, const int32_t* __restrict__ sceneBox, uint32_t* __restrict__ keys,
                         uint32_t* __restrict__ vals) {
    d_morton(blockIdx.x * blockDim.x + threadIdx.x, primBox, n, sceneBox, keys, vals);
}

// ---------------------------------------------------------------------------------------------
// Stable LSD radix sort, 8-bit digits.  Tile = 256 threads x 8 keys; warp w owns 256 consecutive keys.
__global__ void __launch_bounds__(kSortThreads) k_sort_hist(const uint32_t* __restrict__ keys, uint32_t n, int shift, uint32_t numTiles,
                                                            uint32_t* __restrict__ histG) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * kSortTile;
#pragma unroll
    for(int j = 0; j < kSortItems; ++j) {
        const uint32_t i = base + j * kSortThreads + threadIdx.x;
        if(i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    histG[threadIdx.x * numTiles + blockIdx.x] = h[threadIdx.x];
}

// Exclusive scan of histG (digit-major), single block.
__global__ void __launch_bounds__(1024) k_sort_scan(uint32_t* __restrict__ histG, uint32_t count) {
    __shared__ uint32_t warpSums[32];
    __shared__ uint32_t carry;
    if(threadIdx.x == 0) carry = 0;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for(uint32_t base = 0; base < count; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < count ? histG[i] : 0u;
        uint32_t x = v;
#pragma unroll
        for(int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if(lane >= (uint32_t)o) x += y; }
        if(lane == 31) warpSums[warp] = x;
        __syncthreads();
        if(warp == 0) {
            uint32_t w = warpSums[lane];
#pragma unroll
            for(int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, w, o); if(lane >= (uint32_t)o) w += y; }
            warpSums[lane] = w;
        }
        __syncthreads();
        const uint32_t warpOff = warp ? warpSums[warp - 1] : 0u;
        const uint32_t c = carry;
        if(i < count) histG[i] = c + warpOff + x - v;
        __syncthreads();
        if(threadIdx.x == 1023) carry = c + warpOff + x;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kSortThreads) k_sort_scatter(const uint32_t* __restrict__ keysIn, const uint32_t* __restrict__ valsIn, uint32_t n,
                                                               int shift, uint32_t numTiles, const uint32_t* __restrict__ histG,
                                                               uint32_t* __restrict__ keysOut, uint32_t* __restrict__ valsOut) {
    __shared__ uint32_t whist[kSortThreads / 32][256];
    __shared__ uint32_t gbase[256];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for(int w = 0; w < kSortThreads / 32; ++w) whist[w][threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * kSortTile + warp * (32 * kSortItems);
    uint32_t key[kSortItems], val[kSortItems], local[kSortItems];
#pragma unroll
    for(int j = 0; j < kSortItems; ++j) {
        const uint32_t i = base + j * 32 + lane;
        const bool valid = i < n;
        key[j] = valid ? keysIn[i] : 0u;
        val[j] = valid ? valsIn[i] : 0u;
        const uint32_t digit = valid ? ((key[j] >> shift) & 255u) : 0xffffu;
        const uint32_t peers = __match_any_sync(0xffffffffu, digit);
        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        uint32_t b = 0;
        if(valid) b = whist[warp][digit];
        __syncwarp();
        if(valid && rank == 0) whist[warp][digit] = b + __popc(peers);
        __syncwarp();
        local[j] = b + rank;
    }
    __syncthreads();
    {
        uint32_t running = 0;
#pragma unroll
        for(int w = 0; w < kSortThreads / 32; ++w) { const uint32_t c = whist[w][threadIdx.x]; whist[w][threadIdx.x] = running; running += c; }
        gbase[threadIdx.x] = histG[threadIdx.x * numTiles + blockIdx.x];
    }
    __syncthreads();
#pragma unroll
    for(int j = 0; j < kSortItems; ++j) {
        const uint32_t i = base + j * 32 + lane;
        if(i < n) {
            const uint32_t digit = (key[j] >> shift) & 255u;
            const uint32_t pos = gbase[digit] + whist[warp][digit] + local[j];
            keysOut[pos] = key[j];
            valsOut[pos] = val[j];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Karras 2012: one thread per internal node.
__device__ __forceinline__ int deltaKey(const uint32_t* __restrict__ keys, int n, int i, int j) {
    if(j < 0 || j >= n) return -1;
    const uint32_t a = keys[i], b = keys[j];
    return a == b ? 32 + __clz((uint32_t)i ^ (uint32_t)j) : __clz(a ^ b);
}

__device__ __forceinline__ void d_hierarchy(int i, const uint32_t* __restrict__ keys, uint32_t n, BNode* __restrict__ bnodes, uint2* __restrict__ range,
                                            uint32_t* __restrict__ parent, uint32_t* __restrict__ flags) {
    const int N = (int)n;
    if(i >= N - 1) return;
    const int d = (deltaKey(keys, N, i, i + 1) - deltaKey(keys, N, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = deltaKey(keys, N, i, i - d);
    int lmax = 2;
    while(deltaKey(keys, N, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for(int t = lmax >> 1; t >= 1; t >>= 1)
        if(deltaKey(keys, N, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = deltaKey(keys, N, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if(deltaKey(keys, N, i, i + (s + t) * d) > dnode) s += t;
    } while(t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int first = min(i, j), last = max(i, j);
    const uint32_t left = (first == gamma) ? (kLeafBit | (uint32_t)gamma) : (uint32_t)gamma;
    const uint32_t right = (last == gamma + 1) ? (kLeafBit | (uint32_t)(gamma + 1)) : (uint32_t)(gamma + 1);
    bnodes[i].left = left; bnodes[i].right = right;
    range[i] = make_uint2((uint32_t)first, (uint32_t)last);
    flags[i] = 0;
    if(left & kLeafBit) parent[(N - 1) + gamma] = (uint32_t)i; else parent[gamma] = (uint32_t)i;
    if(right & kLeafBit) parent[(N - 1) + gamma + 1] = (uint32_t)i; else parent[gamma + 1] = (uint32_t)i;
    if(i == 0) parent[0] = kInvalid;
}
__global__ void k_hierarchy(const uint32_t* __restrict__ keys, uint32_t n, BNode* __restrict__ bnodes, uint2* __restrict__ range,
                            uint32_t* __restrict__ parent, uint32_t* __restrict__ flags) {
    d_hierarchy((int)(blockIdx.x * blockDim.x + threadIdx.x), keys, n, bnodes, range, parent, flags);
}

__device__ __forceinline__ void loadRefBox(uint32_t ref, const BNode* bnodes, const Aabb* primBox, const uint32_t* vals, float lo[3], float hi[3]) {
    if(ref & kLeafBit) {
        const Aabb b = primBox[vals[ref & ~kLeafBit]];
        for(int a = 0; a < 3; ++a) { lo[a] = b.lo[a]; hi[a] = b.hi[a]; }
    } else {
        const BNode b = bnodes[ref];
        for(int a = 0; a < 3; ++a) { lo[a] = b.lo[a]; hi[a] = b.hi[a]; }
    }
}

// Bottom-up refit: one thread per leaf; the second thread to arrive at a node computes its box.
__device__ __forceinline__ void d_refit_binary(uint32_t s, uint32_t n, BNode* bnodes, const Aabb* __restrict__ primBox, const uint32_t* __restrict__ vals,
                                               const uint32_t* __restrict__ parent, uint32_t* flags) {
    if(s >= n) return;
    uint32_t node = parent[(n - 1) + s];
    while(node != kInvalid) {
        __threadfence();
        if(atomicAdd(&flags[node], 1u) == 0u) return;  // first arrival: sibling not done yet
        __threadfence();
        volatile BNode* vb = bnodes;
        const uint32_t l = vb[node].left, r = vb[node].right;
        float lo[3], hi[3], lo2[3], hi2[3];
        if(l & kLeafBit) { const Aabb b = primBox[vals[l & ~kLeafBit]]; for(int a = 0; a < 3; ++a) { lo[a] = b.lo[a]; hi[a] = b.hi[a]; } }
        else { for(int a = 0; a < 3; ++a) { lo[a] = vb[l].lo[a]; hi[a] = vb[l].hi[a]; } }
        if(r & kLeafBit) { const Aabb b = primBox[vals[r & ~kLeafBit]]; for(int a = 0; a < 3; ++a) { lo2[a] = b.lo[a]; hi2[a] = b.hi[a]; } }
        else { for(int a = 0; a < 3; ++a) { lo2[a] = vb[r].lo[a]; hi2[a] = vb[r].hi[a]; } }
        for(int a = 0; a < 3; ++a) { vb[node].lo[a] = fminf(lo[a], lo2[a]); vb[node].hi[a] = fmaxf(hi[a], hi2[a]); }
        node = parent[node];
    }
}
__global__ void k_refit_binary(uint32_t n, BNode* bnodes, const Aabb* __restrict__ primBox, const uint32_t* __restrict__ vals,
                               const uint32_t* __restrict__ parent, uint32_t* flags) {
    d_refit_binary(blockIdx.x * blockDim.x + threadIdx.x, n, bnodes, primBox, vals, parent, flags);
}

__global__ void k_reset_flags(uint32_t* flags, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) flags[i] = 0;
}

// ---------------------------------------------------------------------------------------------
// Collapse the binary tree into 8-wide compressed nodes; one thread per wide node, level by level.
struct LeafSourceTri { static constexpr uint32_t kMaxPrims = kMaxLeafTris; const float4* vertices; const uint32_t* indices; uint32_t vtxOff, idxOff; Tri* out; };
struct LeafSourceInst { static constexpr uint32_t kMaxPrims = kMaxLeafInsts; const InstTrav* inst; InstTrav* out; };

__device__ __forceinline__ uint32_t refCount(uint32_t ref, const uint2* range) {
    if(ref & kLeafBit) return 1u;
    const uint2 r = range[ref];
    return r.y - r.x + 1u;
}
__device__ __forceinline__ uint32_t refFirst(uint32_t ref, const uint2* range) { return (ref & kLeafBit) ? (ref & ~kLeafBit) : range[ref].x; }

__device__ __forceinline__ void writeLeafPrim(const LeafSourceTri& src, uint32_t dst, uint32_t primId) {
    const uint32_t i0 = src.indices[src.idxOff + 3 * primId], i1 = src.indices[src.idxOff + 3 * primId + 1], i2 = src.indices[src.idxOff + 3 * primId + 2];
    const float4 a = src.vertices[2 * (size_t)(src.vtxOff + i0)], b = src.vertices[2 * (size_t)(src.vtxOff + i1)], c = src.vertices[2 * (size_t)(src.vtxOff + i2)];
    float4* o = reinterpret_cast<float4*>(src.out + dst);
    o[0] = make_float4(a.x, a.y, a.z, __uint_as_float(primId));
    o[1] = make_float4(b.x, b.y, b.z, 0.0f);
    o[2] = make_float4(c.x, c.y, c.z, 0.0f);
}
__device__ __forceinline__ void writeLeafPrim(const LeafSourceInst& src, uint32_t dst, uint32_t primId) { src.out[dst] = src.inst[primId]; }

// Quantise one wide node from its children's boxes.  Conservative in exact arithmetic: the decoded
// planes p + q * 2^e (exact in binary64) never cut into a child box.
__device__ void quantizeNode(Node8& nd, const float (*clo)[3], const float (*chi)[3], const int* posOfChild, int nChildren) {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for(int c = 0; c < nChildren; ++c)
        for(int a = 0; a < 3; ++a) { lo[a] = fminf(lo[a], clo[c][a]); hi[a] = fmaxf(hi[a], chi[c][a]); }
    nd.px = lo[0]; nd.py = lo[1]; nd.pz = lo[2];
    uint8_t eb[3];
    float scale[3];
    for(int a = 0; a < 3; ++a) {
        const float ext = hi[a] - lo[a];
        int k = 0;
        if(ext > 0.0f) (void)frexpf(ext / 255.0f, &k); else k = -120;
        k = max(-120, min(100, k));
        // fl(hi - lo) may round down: make sure 255 steps really reach hi
        while((double)lo[a] + 255.0 * (double)__uint_as_float((uint32_t)(k + 127) << 23) < (double)hi[a]) ++k;
        eb[a] = (uint8_t)(k + 127);
        scale[a] = __uint_as_float((uint32_t)eb[a] << 23);
    }
    // stored with the bias of the traversal's byte decode (binary32 formulation: 1 + q / 32768, rg_trace.cu byteF: 2^(e + 15); none for RG_HALF_SLAB)
    nd.ex = (uint8_t)(eb[0] + kExpBias); nd.ey = (uint8_t)(eb[1] + kExpBias); nd.ez = (uint8_t)(eb[2] + kExpBias);
    uint8_t* qlo[3] = {nd.qlox, nd.qloy, nd.qloz};
    uint8_t* qhi[3] = {nd.qhix, nd.qhiy, nd.qhiz};
    for(int s = 0; s < 8; ++s)
        for(int a = 0; a < 3; ++a) { qlo[a][s] = 255; qhi[a][s] = 0; }  // empty: inverted box
    for(int c = 0; c < nChildren; ++c) {
        const int s = posOfChild[c];
        for(int a = 0; a < 3; ++a) {
            const float inv = 1.0f / scale[a];
            int ql = (int)floorf((clo[c][a] - lo[a]) * inv), qh = (int)ceilf((chi[c][a] - lo[a]) * inv);
            ql = max(0, min(255, ql)); qh = max(0, min(255, qh));
            while(ql > 0 && (double)lo[a] + (double)ql * (double)scale[a] > (double)clo[c][a]) --ql;
            while(qh < 255 && (double)lo[a] + (double)qh * (double)scale[a] < (double)chi[c][a]) ++qh;
            qlo[a][s] = (uint8_t)ql; qhi[a][s] = (uint8_t)qh;
        }
    }
#if RG_PLANE_DIFF
    for(int a = 0; a < 3; ++a)
        for(int w = 0; w < 2; ++w) {   // high words -> hi - lo (mod 2^32), see rg_types.cuh
            uint32_t* h = reinterpret_cast<uint32_t*>(qhi[a]) + w;
            *h -= reinterpret_cast<const uint32_t*>(qlo[a])[w];
        }
#endif
}

template <class LeafSource>
__device__ void collapseItem(const uint2 it, uint2* __restrict__ queueOut, uint32_t* nOutPtr, uint32_t* nodeCounter, uint32_t* primCounter,
                             const BNode* __restrict__ bnodes, const uint2* __restrict__ range, const Aabb* __restrict__ primBox,
                             const uint32_t* __restrict__ vals, Node8* __restrict__ nodes, uint32_t nodeOffset, uint32_t primOffset, LeafSource leafSrc,
                             uint32_t* __restrict__ wideRef) {
    const uint32_t bref = it.x, wide = it.y;

    // Box, surface, primitive count and (for internal refs) the two children of every current child are loaded ONCE, when the child
    // appears: the opening loop below is then one round of (independent) loads per opened subtree -- the latency of an item is what
    // bounds a level, and the levels of a small tree (the per-frame TLAS) run one after the other.
    uint32_t c[8], cLeft[8], cRight[8], cCnt[8];
    float clo[8][3], chi[8][3], cArea[8];
    auto fetch = [&](int k, uint32_t ref) {
        c[k] = ref;
        if(ref & kLeafBit) {
            const Aabb bx = primBox[vals[ref & ~kLeafBit]];
            for(int a = 0; a < 3; ++a) { clo[k][a] = bx.lo[a]; chi[k][a] = bx.hi[a]; }
            cLeft[k] = cRight[k] = kInvalid; cCnt[k] = 1u;
        } else {
            const BNode bn = bnodes[ref];
            const uint2 rg = range[ref];
            for(int a = 0; a < 3; ++a) { clo[k][a] = bn.lo[a]; chi[k][a] = bn.hi[a]; }
            cLeft[k] = bn.left; cRight[k] = bn.right; cCnt[k] = rg.y - rg.x + 1u;
        }
        const float dx = chi[k][0] - clo[k][0], dy = chi[k][1] - clo[k][1], dz = chi[k][2] - clo[k][2];
        cArea[k] = dx * dy + dy * dz + dz * dx;
    };
    int n;
    if(bref & kLeafBit) { fetch(0, bref); n = 1; }
    else { const BNode root = bnodes[bref]; fetch(0, root.left); fetch(1, root.right); n = 2; }

    // phase 0: open subtrees with more than LeafSource::kMaxPrims primitives (largest surface first);
    // phase 1: with free slots left, open small subtrees too (tighter boxes, one primitive per slot).
#ifndef RG_COLLAPSE_PHASES
#define RG_COLLAPSE_PHASES 2
#endif
    for(int phase = 0; phase < RG_COLLAPSE_PHASES; ++phase) {
        while(n < 8) {
            int best = -1; float bestArea = -1.0f;
            for(int k = 0; k < n; ++k) {
                const bool open = phase == 0 ? (cCnt[k] > LeafSource::kMaxPrims) : (cCnt[k] > 1u);
                if(open && cArea[k] > bestArea) { bestArea = cArea[k]; best = k; }
            }
            if(best < 0) break;
            const uint32_t l = cLeft[best], r = cRight[best];
            fetch(best, l);
            fetch(n++, r);
        }
    }

    float ctr[3] = {0, 0, 0}, nlo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, nhi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for(int k = 0; k < n; ++k)
        for(int a = 0; a < 3; ++a) { nlo[a] = fminf(nlo[a], clo[k][a]); nhi[a] = fmaxf(nhi[a], chi[k][a]); }
    for(int a = 0; a < 3; ++a) ctr[a] = 0.5f * (nlo[a] + nhi[a]);

    // greedy octant assignment: slot bit 2/1/0 = +x/+y/+z side of the node centre.  n rounds, each takes the (child, slot) pair of
    // largest cost among the free ones (first in child-major order on ties).  The 64 candidates of a round are unrolled over
    // registers (done flags as bit masks): this is the longest serial stretch of an item.
    int slotOfChild[8];
    {
        float d[8][3];
#pragma unroll
        for(int k = 0; k < 8; ++k)
#pragma unroll
            for(int a = 0; a < 3; ++a) d[k][a] = k < n ? 0.5f * (clo[k][a] + chi[k][a]) - ctr[a] : 0.0f;
        uint32_t childDone = ~((1u << n) - 1u), slotDone = 0u, packed = 0u;   // children >= n never compete
#pragma unroll 1
        for(int round = 0; round < n; ++round) {
            float bestCost = -FLT_MAX; int bk = 0, bs = 0;
#pragma unroll
            for(int k = 0; k < 8; ++k) {
#pragma unroll
                for(int sl = 0; sl < 8; ++sl) {
                    const float cost = ((sl & 4) ? d[k][0] : -d[k][0]) + ((sl & 2) ? d[k][1] : -d[k][1]) + ((sl & 1) ? d[k][2] : -d[k][2]);
                    const bool freePair = !((childDone >> k) & 1u) && !((slotDone >> sl) & 1u);
                    if(freePair && cost > bestCost) { bestCost = cost; bk = k; bs = sl; }
                }
            }
            childDone |= 1u << bk; slotDone |= 1u << bs; packed |= (uint32_t)bs << (4 * bk);
        }
        for(int k = 0; k < n; ++k) slotOfChild[k] = (int)((packed >> (4 * k)) & 7u);
    }

    // nibbles: leaves first, then internal children; the plane bytes of nibble m go to position posOfNibble(m) (rg_types.cuh)
    bool isInner[8];
    int posOfChild[8], childOfSlot[8];
    uint32_t nLeaf = 0, nInternal = 0;
    for(int s = 0; s < 8; ++s) childOfSlot[s] = -1;
    for(int k = 0; k < n; ++k) {
        isInner[k] = cCnt[k] > LeafSource::kMaxPrims;
        if(!isInner[k]) posOfChild[k] = posOfNibble((int)nLeaf++);
        childOfSlot[slotOfChild[k]] = k;
    }
    for(int k = 0; k < n; ++k)
        if(isInner[k]) posOfChild[k] = posOfNibble((int)(nLeaf + nInternal++));

    Node8 nd;
    quantizeNode(nd, clo, chi, posOfChild, n);

    const uint32_t childBase = nInternal ? atomicAdd(nodeCounter, nInternal) : 0u;
    const uint32_t primBase = nLeaf ? atomicAdd(primCounter, kLeafStride * nLeaf) : 0u;
    const uint32_t qBase = nInternal ? atomicAdd(nOutPtr, nInternal) : 0u;
    uint32_t refOfPos[8];
    for(int j = 0; j < 8; ++j) refOfPos[j] = kInvalid;
    uint32_t imask = 0, codes = 0x88888888u, vm = 0, ci = 0;
    for(int s = 0; s < 8; ++s) {   // code order: the order of the child nodes in memory
        const int k = childOfSlot[s];
        if(k < 0) continue;
        const int j = posOfChild[k], m = nibbleOfPos(j);
        refOfPos[j] = c[k];
        if(isInner[k]) {
            imask |= 1u << s;
            codes = (codes & ~(0xFu << (4 * m))) | ((uint32_t)s << (4 * m));
            vm |= 8u << (4 * m);
            queueOut[qBase + ci] = make_uint2(c[k], childBase + ci);
            ++ci;
        } else {
            const uint32_t cnt = cCnt[k], first = refFirst(c[k], range);
            vm |= ((1u << cnt) - 1u) << (4 * m);
            for(uint32_t kk = 0; kk < cnt; ++kk) writeLeafPrim(leafSrc, primOffset + primBase + kLeafStride * (uint32_t)m + kk, vals[first + kk]);
        }
    }
    if(wideRef)
        for(int j = 0; j < 8; ++j) wideRef[(size_t)wide * 8 + j] = refOfPos[j];
    nd.imask = (uint8_t)imask;
    nd.childBase = nodeOffset + childBase;
    nd.primBase = (primOffset + primBase) | kPrimGroupBit;
    nd.codes = codes;
    nd.vm = vm;
    nodes[nodeOffset + wide] = nd;
}

// Barrier over the whole (co-resident: cooperative launch) grid.  bar counts arrivals of all epochs; epoch e is complete at e * gridDim.x.
// A two-second escape turns a lost CTA into an error flag instead of a hung GPU.
__device__ __forceinline__ void gridBarrier(uint32_t* bar, uint32_t epoch, uint32_t* err) {
    __syncthreads();
    if(threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        const uint32_t target = epoch * gridDim.x;
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while(*reinterpret_cast<volatile uint32_t*>(bar) < target) {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if(t1 - t0 > 2000000000ull) { *err = 1u; break; }
        }
        __threadfence();
    }
    __syncthreads();
}

// Every collapse level inside ONE persistent grid (no host round trip, all SMs): level L takes its items from queue[L & 1], appends the
// internal children to queue[(L + 1) & 1] and counts them in levelCount[L + 1]; a grid barrier separates the levels.  sync: [0] barrier,
// [1] error flag, [16 + L] items of level L (all zero on entry).  counters: [2] wide nodes, [3] primitive slots.
template <class LeafSource>
__global__ void __launch_bounds__(256) k_collapse_grid(uint2* queue0, uint2* queue1, uint32_t* counters, uint32_t* sync, uint32_t n, const BNode* __restrict__ bnodes,
                                                       const uint2* __restrict__ range, const Aabb* __restrict__ primBox,
                                                       const uint32_t* __restrict__ vals, Node8* __restrict__ nodes, uint32_t nodeOffset,
                                                       uint32_t primOffset, LeafSource leafSrc, uint32_t* __restrict__ wideRef) {
    uint32_t* levelCount = sync + 16;
    if(blockIdx.x == 0 && threadIdx.x == 0) {
        queue0[0] = make_uint2(n >= 2 ? 0u : kLeafBit, 0u);   // root item: binary node 0 (or the single leaf) -> wide node 0
        levelCount[0] = 1u; counters[2] = 1u; counters[3] = 0u;
    }
    gridBarrier(sync, 1u, sync + 1);
    for(uint32_t level = 0; level < 110u; ++level) {
        const uint32_t count = __ldcg(levelCount + level);
        if(count == 0u) break;
        const uint2* qIn = (level & 1u) ? queue1 : queue0;
        uint2* qOut = (level & 1u) ? queue0 : queue1;
        for(uint32_t item = blockIdx.x * blockDim.x + threadIdx.x; item < count; item += gridDim.x * blockDim.x)   // queue slots are reused: read past L1
            collapseItem<LeafSource>(__ldcg(qIn + item), qOut, levelCount + level + 1, &counters[2], &counters[3], bnodes, range, primBox, vals, nodes, nodeOffset,
                                     primOffset, leafSrc, wideRef);
        gridBarrier(sync, level + 2u, sync + 1);
    }
}

// world->object from the 3x4 (binary64, explicitly rounded operations; same formula as oracle/orc_scene.cpp invert3x4)
__device__ __forceinline__ void d_prepare_instance(uint32_t i, const rg_instance* __restrict__ raw, uint32_t n, const uint32_t* __restrict__ meshRoots,
                                                   uint32_t nMeshes, InstTrav* __restrict__ trav, InstShade* __restrict__ shade, const float4* __restrict__ meshSpheres) {
    if(i >= n) return;
    const rg_instance in = raw[i];
    const float* m = in.xform;
    const double a = m[0], b = m[1], c = m[2], d = m[4], e = m[5], f = m[6], g = m[8], h = m[9], k = m[10];
#define DM(x, y) __dmul_rn((x), (y))
#define DS(x, y) __dsub_rn((x), (y))
#define DA(x, y) __dadd_rn((x), (y))
    const double A = DS(DM(e, k), DM(f, h)), B = -DS(DM(d, k), DM(f, g)), C = DS(DM(d, h), DM(e, g));
    const double det = DA(DA(DM(a, A), DM(b, B)), DM(c, C));
    const double r = __ddiv_rn(1.0, det);
    double R[9];
    R[0] = DM(A, r); R[1] = DM(-DS(DM(b, k), DM(c, h)), r); R[2] = DM(DS(DM(b, f), DM(c, e)), r);
    R[3] = DM(B, r); R[4] = DM(DS(DM(a, k), DM(c, g)), r); R[5] = DM(-DS(DM(a, f), DM(c, d)), r);
    R[6] = DM(C, r); R[7] = DM(-DS(DM(a, h), DM(b, g)), r); R[8] = DM(DS(DM(a, e), DM(b, d)), r);
    const double tx = m[3], ty = m[7], tz = m[11];
    InstTrav t;
    for(int rr = 0; rr < 3; ++rr) {
        t.w2o[rr * 4 + 0] = (float)R[rr * 3 + 0]; t.w2o[rr * 4 + 1] = (float)R[rr * 3 + 1]; t.w2o[rr * 4 + 2] = (float)R[rr * 3 + 2];
        t.w2o[rr * 4 + 3] = (float)(-DA(DA(DM(R[rr * 3 + 0], tx), DM(R[rr * 3 + 1], ty)), DM(R[rr * 3 + 2], tz)));
    }
#undef DM
#undef DS
#undef DA
    t.blasRoot = in.mesh < nMeshes ? meshRoots[in.mesh] : kInvalid;
    t.instId = i;
    // pure translation: the traversal keeps the ray direction and everything derived from it (bit-identical to the general path)
    t.pad0 = (m[0] == 1.0f && m[1] == 0.0f && m[2] == 0.0f && m[4] == 0.0f && m[5] == 1.0f && m[6] == 0.0f && m[8] == 0.0f && m[9] == 0.0f && m[10] == 1.0f) ? 1u : 0u;
    t.pad1 = in.mesh < nMeshes ? in.mesh : 0u;
    const float4 sph = (meshSpheres && in.mesh < nMeshes) ? meshSpheres[in.mesh] : make_float4(0.0f, 0.0f, 0.0f, FLT_MAX);
    t.sphere[0] = sph.x; t.sphere[1] = sph.y; t.sphere[2] = sph.z; t.sphere[3] = sph.w;
    trav[i] = t;
    InstShade s;
    for(int j = 0; j < 12; ++j) s.o2w[j] = m[j];
    s.vtxOff = in.vtx_off; s.idxOff = in.idx_off; s.matOff = in.mat_off; s.mesh = in.mesh < nMeshes ? in.mesh : 0;
    shade[i] = s;
}
__global__ void k_prepare_instances(const rg_instance* __restrict__ raw, uint32_t n, const uint32_t* __restrict__ meshRoots, uint32_t nMeshes,
                                    InstTrav* __restrict__ trav, InstShade* __restrict__ shade, const float4* __restrict__ meshSpheres) {
    d_prepare_instance(blockIdx.x * blockDim.x + threadIdx.x, raw, n, meshRoots, nMeshes, trav, shade, meshSpheres);
}

struct TlasFusedArgs {
    const rg_instance* raw; uint32_t n; const uint32_t* meshRoots; uint32_t nMeshes; InstTrav* trav; InstShade* shade; const float* meshBoxes; const float4* meshSpheres;
    Aabb* primBox; int32_t* sceneBox; uint32_t* keys0; uint32_t* vals0; uint32_t* keys1; uint32_t* vals1;
    BNode* bnodes; uint2* range; uint32_t* parent; uint32_t* flags; uint2* queue0; uint2* queue1; uint32_t* counters; uint32_t* wideRef;
    Node8* tlasNodes; InstTrav* tlasLeaves;
};

// The whole per-frame TLAS build in ONE block for small scenes (n <= kTlasFusedMax): instance preparation, world boxes, Morton keys,
// a stable rank sort (identical order to the radix sort), Karras hierarchy, bottom-up refit and every collapse level.
// One launch instead of ~20: the per-frame acceleration-structure cost drops to the latency of one small kernel.
__global__ void __launch_bounds__(1024) k_tlas_fused(const TlasFusedArgs A) {
    __shared__ uint32_t sCount;
    const uint32_t n = A.n, tid = threadIdx.x, nt = blockDim.x;
    const uint32_t nUp = (n + 31u) & ~31u;
    if(tid < 3) A.sceneBox[tid] = 0x7fffffff; else if(tid < 6) A.sceneBox[tid] = (int)0x80000000;
    for(uint32_t i = tid; i < n; i += nt) d_prepare_instance(i, A.raw, n, A.meshRoots, A.nMeshes, A.trav, A.shade, A.meshSpheres);
    __syncthreads();
    for(uint32_t i = tid; i < nUp; i += nt) d_inst_boxes(i, A.shade, A.meshBoxes, n, A.primBox, A.sceneBox);   // whole warps: shuffles inside
    __syncthreads();
    for(uint32_t i = tid; i < n; i += nt) d_morton(i, A.primBox, n, A.sceneBox, A.keys0, A.vals0);
    __syncthreads();
    for(uint32_t i = tid; i < n; i += nt) {   // rank sort by (key, index): the order of a stable sort
        d_tighten_inst_box(i, A.shade, A.meshSpheres, A.nMeshes, A.primBox);   // the keys are taken; the boxes are next read by the refit
        const uint32_t ki = A.keys0[i];
        uint32_t rank = 0;
        for(uint32_t j = 0; j < n; ++j) { const uint32_t kj = A.keys0[j]; rank += (kj < ki || (kj == ki && j < i)) ? 1u : 0u; }
        A.keys1[rank] = ki; A.vals1[rank] = i;
    }
    __syncthreads();
    if(n >= 2) {
        for(uint32_t i = tid; i + 1 < n; i += nt) d_hierarchy((int)i, A.keys1, n, A.bnodes, A.range, A.parent, A.flags);
        __syncthreads();
        for(uint32_t i = tid; i < n; i += nt) d_refit_binary(i, n, A.bnodes, A.primBox, A.vals1, A.parent, A.flags);
        __threadfence();
        __syncthreads();
    }
    // collapse, level by level (as k_collapse_all)
    if(tid == 0) {
        A.queue0[0] = make_uint2(n >= 2 ? 0u : kLeafBit, 0u);
        A.counters[0] = 1; A.counters[1] = 0; A.counters[2] = 1; A.counters[3] = 0;
    }
    __syncthreads();
    const LeafSourceInst ls{A.trav, A.tlasLeaves};
    int in = 0;
    while(true) {
        if(tid == 0) sCount = A.counters[in];
        __syncthreads();
        const uint32_t count = sCount;
        if(count == 0) break;
        uint2* qIn = in ? A.queue1 : A.queue0;
        uint2* qOut = in ? A.queue0 : A.queue1;
        for(uint32_t item = tid; item < count; item += nt)
            collapseItem<LeafSourceInst>(qIn[item], qOut, &A.counters[in ^ 1], &A.counters[2], &A.counters[3], A.bnodes, A.range, A.primBox, A.vals1, A.tlasNodes,
                                         0, 0, ls, A.wideRef);
        __threadfence_block();
        __syncthreads();
        if(tid == 0) A.counters[in] = 0;
        in ^= 1;
        __syncthreads();
    }
}

// Refit of the wide nodes after the binary boxes changed: re-quantise every node from wideRef.
template <class LeafSource>
__global__ void k_requantize(uint32_t nWide, Node8* __restrict__ nodes, uint32_t nodeOffset, uint32_t primOffset, const uint32_t* __restrict__ wideRef,
                             const BNode* __restrict__ bnodes, const uint2* __restrict__ range, const Aabb* __restrict__ primBox,
                             const uint32_t* __restrict__ vals, LeafSource leafSrc) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if(w >= nWide) return;
    Node8 nd = nodes[nodeOffset + w];
    float clo[8][3], chi[8][3];
    int posOfChild[8];
    int n = 0;
    const uint32_t pb = nd.primBase & ~kPrimGroupBit;
    for(int j = 0; j < 8; ++j) {
        const uint32_t r = wideRef[(size_t)w * 8 + j];
        if(r == kInvalid) continue;
        loadRefBox(r, bnodes, primBox, vals, clo[n], chi[n]);
        posOfChild[n++] = j;
        const uint32_t cnt = refCount(r, range);
        if(cnt <= LeafSource::kMaxPrims) {
            const uint32_t first = refFirst(r, range);
            for(uint32_t k = 0; k < cnt; ++k) writeLeafPrim(leafSrc, pb + kLeafStride * (uint32_t)nibbleOfPos(j) + k, vals[first + k]);
        }
    }
    const uint8_t imask = nd.imask;
    const uint32_t cb = nd.childBase, pbFlag = nd.primBase, codes = nd.codes, vm = nd.vm;
    quantizeNode(nd, clo, chi, posOfChild, n);
    nd.imask = imask; nd.childBase = cb; nd.primBase = pbFlag; nd.codes = codes; nd.vm = vm;
    nodes[nodeOffset + w] = nd;
    (void)primOffset;
}

// Bounding sphere of a mesh around the centre of its box: r^2 = max |v - c|^2 over the vertices its triangles use.  The traversal tests a
// ray against it before it enters an instance of the mesh (rg_trace.cu travPrim): for round meshes the sphere is half the volume of
// the box the TLAS knows, and an avoided entry saves the ray transform, the set-up and the BLAS root visit.
__global__ void k_mesh_sphere_reduce(const float4* __restrict__ vertices, const uint32_t* __restrict__ indices, uint32_t vtxOff, uint32_t idxOff, uint32_t nTri,
                                     const int32_t* __restrict__ sceneBox, uint32_t* maxBits) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    const float cx = 0.5f * (decodeFloat(sceneBox[0]) + decodeFloat(sceneBox[3])), cy = 0.5f * (decodeFloat(sceneBox[1]) + decodeFloat(sceneBox[4])),
                cz = 0.5f * (decodeFloat(sceneBox[2]) + decodeFloat(sceneBox[5]));
    float d2 = 0.0f;
    if(p < nTri) {
#pragma unroll
        for(int k = 0; k < 3; ++k) {
            const float4 v = vertices[2 * (size_t)(vtxOff + indices[idxOff + 3 * p + k])];
            const float dx = v.x - cx, dy = v.y - cy, dz = v.z - cz;
            d2 = fmaxf(d2, dx * dx + dy * dy + dz * dz);
        }
    }
#pragma unroll
    for(int o = 16; o; o >>= 1) d2 = fmaxf(d2, __shfl_xor_sync(0xffffffffu, d2, o));
    if((threadIdx.x & 31) == 0) atomicMax(maxBits, __float_as_uint(d2));   // non-negative floats order like their bit patterns
}
__global__ void k_mesh_sphere_store(const int32_t* sceneBox, const uint32_t* maxBits, uint32_t n, float4* out) {
    if(threadIdx.x != 0) return;
    if(n == 0) { *out = make_float4(0.0f, 0.0f, 0.0f, FLT_MAX); return; }
    const float lx = decodeFloat(sceneBox[0]), ly = decodeFloat(sceneBox[1]), lz = decodeFloat(sceneBox[2]);
    const float hx = decodeFloat(sceneBox[3]), hy = decodeFloat(sceneBox[4]), hz = decodeFloat(sceneBox[5]);
    const float r2 = __uint_as_float(*maxBits) * 1.0001f + FLT_MIN;   // rounding of the distances above
    const float r = sqrtf(r2);
    const float vSphere = 4.18879f * r * r2, vBox = (hx - lx) * (hy - ly) * (hz - lz);
    // only worth a test where it can reject what the box lets through: w = FLT_MAX switches it off (flat and boxy meshes)
    *out = make_float4(0.5f * (lx + hx), 0.5f * (ly + hy), 0.5f * (lz + hz), vSphere < 0.8f * vBox ? r2 : FLT_MAX);
}

__global__ void k_store_root_box(const int32_t* sceneBox, float* out, uint32_t n) {
    if(threadIdx.x < 6) {
        if(n == 0) out[threadIdx.x] = threadIdx.x < 3 ? 1.0f : -1.0f;  // empty: inverted
        else out[threadIdx.x] = decodeFloat(sceneBox[threadIdx.x]);
    }
}

inline uint32_t cdiv(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

void radixSort(LbvhScratch& s, uint32_t n, cudaStream_t st) {
    const uint32_t numTiles = cdiv(n, kSortTile);
    int cur = 0;
    for(int pass = 0; pass < 4; ++pass) {
        const int shift = 8 * pass;
        k_sort_hist<<<numTiles, kSortThreads, 0, st>>>(s.keys[cur], n, shift, numTiles, s.hist);
        k_sort_scan<<<1, 1024, 0, st>>>(s.hist, 256u * numTiles);
        k_sort_scatter<<<numTiles, kSortThreads, 0, st>>>(s.keys[cur], s.vals[cur], n, shift, numTiles, s.hist, s.keys[cur ^ 1], s.vals[cur ^ 1]);
        s.launches += 3;
        cur ^= 1;
    }
    s.sortedBuf = (uint32_t)cur;  // 4 passes -> back in buffer 0
}

// Launches k_collapse_grid cooperatively: as many CTAs as the widest level can use, never more than fit on the device at once.
template <class LeafSource>
void collapseGrid(LbvhScratch& s, uint32_t n, Node8* nodes, uint32_t nodeOffset, uint32_t primOffset, LeafSource leafSrc, cudaStream_t st) {
    static int perSm = [] { int v = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k_collapse_grid<LeafSource>, 256, 0); return v < 1 ? 1 : v; }();
    int dev = 0, numSms = 1;
    cudaGetDevice(&dev); cudaDeviceGetAttribute(&numSms, cudaDevAttrMultiProcessorCount, dev);
    uint32_t grid = cdiv(n, 1024u);
    if(grid > (uint32_t)(numSms * perSm)) grid = (uint32_t)(numSms * perSm);
    if(grid < 1u) grid = 1u;
    RG_CUDA_OK(cudaMemsetAsync(s.gridSync, 0, 4 * 128, st));
    uint2 *q0 = s.queue[0], *q1 = s.queue[1];
    uint32_t *counters = s.counters, *sync = s.gridSync, *wideRef = s.wideRef;
    const BNode* bnodes = s.bnodes; const uint2* range = s.range; const Aabb* primBox = s.primBox; const uint32_t* vals = s.vals[s.sortedBuf];
    void* args[] = {&q0, &q1, &counters, &sync, &n, &bnodes, &range, &primBox, &vals, &nodes, &nodeOffset, &primOffset, &leafSrc, &wideRef};
    RG_CUDA_OK(cudaLaunchCooperativeKernel((const void*)k_collapse_grid<LeafSource>, dim3(grid), dim3(256), args, 0, st));
    s.launches += 2;
}

}  // namespace

void LbvhScratch::reserve(uint32_t n) {
    if(n <= capacity) return;
    release();
    capacity = n;
    const uint32_t numTiles = cdiv(n, kSortTile);
    RG_CUDA_OK(cudaMalloc(&primBox, sizeof(Aabb) * (size_t)n));
    for(int i = 0; i < 2; ++i) {
        RG_CUDA_OK(cudaMalloc(&keys[i], 4 * (size_t)n));
        RG_CUDA_OK(cudaMalloc(&vals[i], 4 * (size_t)n));
        RG_CUDA_OK(cudaMalloc(&queue[i], sizeof(uint2) * (size_t)n));
    }
    RG_CUDA_OK(cudaMalloc(&hist, 4 * 256 * (size_t)numTiles));
    RG_CUDA_OK(cudaMalloc(&bnodes, sizeof(BNode) * (size_t)n));
    RG_CUDA_OK(cudaMalloc(&range, sizeof(uint2) * (size_t)n));
    RG_CUDA_OK(cudaMalloc(&parent, 4 * 2 * (size_t)n));
    RG_CUDA_OK(cudaMalloc(&flags, 4 * (size_t)n));
    RG_CUDA_OK(cudaMalloc(&counters, 4 * 16));
    RG_CUDA_OK(cudaMalloc(&gridSync, 4 * 128));
    RG_CUDA_OK(cudaMalloc(&sceneBox, 4 * 6));
    RG_CUDA_OK(cudaMalloc(&wideRef, 4 * 8 * (size_t)n));
}

void LbvhScratch::release() {
    cudaFree(primBox); cudaFree(hist); cudaFree(bnodes); cudaFree(range); cudaFree(parent); cudaFree(flags); cudaFree(counters); cudaFree(sceneBox);
    cudaFree(wideRef); cudaFree(gridSync); gridSync = nullptr;
    for(int i = 0; i < 2; ++i) { cudaFree(keys[i]); cudaFree(vals[i]); cudaFree(queue[i]); keys[i] = vals[i] = nullptr; queue[i] = nullptr; }
    primBox = nullptr; hist = nullptr; bnodes = nullptr; range = nullptr; parent = nullptr; flags = nullptr; counters = nullptr; sceneBox = nullptr;
    wideRef = nullptr;
    capacity = 0;
}

namespace {
struct TightenArgs { const InstShade* inst; const float4* meshSpheres; uint32_t nMeshes; };
void lbvhCommon(LbvhScratch& s, uint32_t n, cudaStream_t st, const TightenArgs* tighten = nullptr) {
    k_morton<<<cdiv(n, 256), 256, 0, st>>>(s.primBox, n, s.sceneBox, s.keys[0], s.vals[0]);
    s.launches++;
    if(tighten) { k_tighten_inst_boxes<<<cdiv(n, 128), 128, 0, st>>>(tighten->inst, tighten->meshSpheres, tighten->nMeshes, n, s.primBox); s.launches++; }
    radixSort(s, n, st);
    if(n >= 2) {
        k_hierarchy<<<cdiv(n - 1, 128), 128, 0, st>>>(s.keys[s.sortedBuf], n, s.bnodes, s.range, s.parent, s.flags);
        k_refit_binary<<<cdiv(n, 128), 128, 0, st>>>(n, s.bnodes, s.primBox, s.vals[s.sortedBuf], s.parent, s.flags);
        s.launches += 2;
    }
}
}  // namespace

namespace {
void meshSphere(LbvhScratch& s, const TriSource& src, float4* sphereOut, cudaStream_t st) {   // after k_tri_boxes filled s.sceneBox
    if(!sphereOut) return;
    RG_CUDA_OK(cudaMemsetAsync(&s.counters[12], 0, sizeof(uint32_t), st));
    if(src.nTri) k_mesh_sphere_reduce<<<cdiv(src.nTri, 256), 256, 0, st>>>((const float4*)src.vertices, src.indices, src.vtxOff, src.idxOff, src.nTri, s.sceneBox, &s.counters[12]);
    k_mesh_sphere_store<<<1, 32, 0, st>>>(s.sceneBox, &s.counters[12], src.nTri, sphereOut);
    s.launches += 2;
}
}  // namespace

void buildBlas(LbvhScratch& s, const TriSource& src, Node8* nodesBase, uint32_t nodeOffset, Tri* trisBase, uint32_t triOffset, uint32_t* nNodes,
               uint32_t* nTris, float* rootBoxOut, float4* sphereOut, cudaStream_t st) {
    const uint32_t n = src.nTri;
    *nNodes = 0; *nTris = 0;
    s.reserve(n ? n : 1);
    k_init_build<<<1, 32, 0, st>>>(s.sceneBox, s.counters);
    s.launches++;
    if(n == 0) { k_store_root_box<<<1, 32, 0, st>>>(s.sceneBox, rootBoxOut, 0); s.launches++; meshSphere(s, src, sphereOut, st); RG_CUDA_OK(cudaStreamSynchronize(st)); return; }
    k_tri_boxes<<<cdiv(n, 256), 256, 0, st>>>((const float4*)src.vertices, src.indices, src.vtxOff, src.idxOff, n, s.primBox, s.sceneBox);
    s.launches++;
    meshSphere(s, src, sphereOut, st);
    lbvhCommon(s, n, st);
    LeafSourceTri ls{(const float4*)src.vertices, src.indices, src.vtxOff, src.idxOff, trisBase};
    collapseGrid(s, n, nodesBase, nodeOffset, triOffset, ls, st);
    k_store_root_box<<<1, 32, 0, st>>>(s.sceneBox, rootBoxOut, n);
    s.launches++;
    uint32_t h[4] = {0, 0, 0, 0}, syncWords[2] = {0, 0};
    RG_CUDA_OK(cudaMemcpyAsync(h, s.counters, sizeof(h), cudaMemcpyDeviceToHost, st));
    RG_CUDA_OK(cudaMemcpyAsync(syncWords, s.gridSync, sizeof(syncWords), cudaMemcpyDeviceToHost, st));
    RG_CUDA_OK(cudaStreamSynchronize(st));   // the ONE host round trip of a BLAS build: the caller needs the node / triangle counts
    if(syncWords[1]) fprintf(stderr, "rgb200: BLAS collapse: grid barrier timed out\n");
    *nNodes = h[2]; *nTris = h[3];
}

void refitBlas(LbvhScratch& s, const TriSource& src, Node8* nodesBase, uint32_t nodeOffset, uint32_t nNodes, Tri* trisBase, uint32_t triOffset,
               float* rootBoxOut, float4* sphereOut, cudaStream_t st) {
    const uint32_t n = src.nTri;
    if(n == 0) return;
    k_init_build<<<1, 32, 0, st>>>(s.sceneBox, s.counters + 4);  // keep queue / node counters; counters+4 is scratch
    k_tri_boxes<<<cdiv(n, 256), 256, 0, st>>>((const float4*)src.vertices, src.indices, src.vtxOff, src.idxOff, n, s.primBox, s.sceneBox);
    s.launches += 2;
    meshSphere(s, src, sphereOut, st);
    if(n >= 2) {
        k_reset_flags<<<cdiv(n - 1, 256), 256, 0, st>>>(s.flags, n - 1);
        k_refit_binary<<<cdiv(n, 128), 128, 0, st>>>(n, s.bnodes, s.primBox, s.vals[s.sortedBuf], s.parent, s.flags);
        s.launches += 2;
    }
    LeafSourceTri ls{(const float4*)src.vertices, src.indices, src.vtxOff, src.idxOff, trisBase};
    k_requantize<LeafSourceTri><<<cdiv(nNodes, 64), 64, 0, st>>>(nNodes, nodesBase, nodeOffset, triOffset, s.wideRef, s.bnodes, s.range, s.primBox,
                                                                 s.vals[s.sortedBuf], ls);
    k_store_root_box<<<1, 32, 0, st>>>(s.sceneBox, rootBoxOut, n);
    s.launches += 2;
}

void buildTlas(LbvhScratch& s, const rg_instance* raw, uint32_t nInst, const uint32_t* meshRoots, uint32_t nMeshes, InstTrav* instTrav,
               InstShade* instShade, const float* meshBoxes, const float4* meshSpheres, Node8* tlasNodes, InstTrav* tlasLeavesOut, cudaStream_t st) {
    const uint32_t n = nInst;
    s.reserve(n ? n : 1);
    if(n == 0) { k_init_build<<<1, 32, 0, st>>>(s.sceneBox, s.counters); s.launches++; return; }
    if(n <= kTlasFusedMax) {   // one launch for the whole build
        TlasFusedArgs a{raw, n, meshRoots, nMeshes, instTrav, instShade, meshBoxes, meshSpheres, s.primBox, s.sceneBox, s.keys[0], s.vals[0], s.keys[1], s.vals[1],
                        s.bnodes, s.range, s.parent, s.flags, s.queue[0], s.queue[1], s.counters, s.wideRef, tlasNodes, tlasLeavesOut};
        k_tlas_fused<<<1, n <= 32 ? 64 : (n <= 256 ? 256 : 1024), 0, st>>>(a);
        s.launches++;
        s.sortedBuf = 1;
        return;
    }
    k_prepare_instances<<<cdiv(n, 128), 128, 0, st>>>(raw, n, meshRoots, nMeshes, instTrav, instShade, meshSpheres);
    k_init_build<<<1, 32, 0, st>>>(s.sceneBox, s.counters);
    k_inst_boxes<<<cdiv(n, 128), 128, 0, st>>>(instShade, meshBoxes, n, s.primBox, s.sceneBox);
    s.launches += 3;
    const TightenArgs tighten{instShade, meshSpheres, nMeshes};
    lbvhCommon(s, n, st, &tighten);
    LeafSourceInst ls{instTrav, tlasLeavesOut};
    collapseGrid(s, n, tlasNodes, 0, 0, ls, st);   // fully asynchronous
}

}  // namespace rg
