// rg_trace.cu -- persistent-thread ray generation + two-level wide-BVH traversal + Whitted shading for sm_100a.
//
// Replaces the reference's ray-tracing pipeline (raygun/render/raytracer.cpp:99 traceRaysKHR) and its four shaders:
//   raygen      resources/shaders/raygen.rgen:30-39, raygen.h:37-115
//   closest hit resources/shaders/closesthit.rchit:74-268
//   miss 0 / 1  resources/shaders/miss.rmiss:38-83, shadowMiss.rmiss:30-34
// B200 has no RT cores, so traceRayEXT becomes a software traversal of compressed 8-wide nodes (rg_types.cuh) and
// the shader recursion becomes an explicit per-lane stack of frames: each lane owns one pixel and walks that pixel's
// ray tree depth-first in the reference's order (shadow -> reflection -> refraction), with ONE mutable payload, so
// every stale-state effect of the GLSL (SURVEY.md 8a hazards 1-6) is reproduced.  Lanes that finish a pixel are
// refilled from a global work counter with warp vote / shuffle compaction, so a warp keeps traversing 32 live rays.
//
// Ray / triangle arithmetic is bit-identical to the oracle (explicitly rounded operations, never contracted):
// watertight Woop test, t preserved across the instance transform, ties resolved to the smallest (instance, primitive).
#include "rg_trace.cuh"

#include <cuda_fp16.h>

#include <cstdlib>

#include "../../include/rgb200.h"

#ifndef RG_LANE_CONTEXTS
#define RG_LANE_CONTEXTS 1
#endif
#ifndef RG_POSTPONE_THRESHOLD
#define RG_POSTPONE_THRESHOLD 12
#endif
#ifndef RG_TRAVERSE_IFIF
#define RG_TRAVERSE_IFIF 1
#endif

namespace rg {

namespace {

enum { RT_GENERIC = 0, RT_SHADOW_TRACE = 1, RT_SHADOW_INTERNAL = 2 };
enum { CNT_PRIMARY = 0, CNT_SHADOW = 1, CNT_REFLECT = 2, CNT_REFRACT = 3, CNT_SKY = 4, CNT_NODES = 5, CNT_TRIS = 6, CNT_INST = 7, CNT_GENHIT = 8, CNT_N = 9 };

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 normalize(V3 a) { const float r = 1.0f / sqrtf(dot(a, a)); return a * r; }
__device__ __forceinline__ V3 mix3(V3 a, V3 b, float t) { return a * (1.0f - t) + b * t; }
__device__ __forceinline__ float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
// GLSL min / max / clamp as GLM evaluates them ((y < x) ? y : x ...): NaN propagates exactly as in the oracle (RG_STRICT_IEEE)
__device__ __forceinline__ float glmin(float x, float y) { return (y < x) ? y : x; }
__device__ __forceinline__ float glmax(float x, float y) { return (x < y) ? y : x; }
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return glmin(glmax(x, lo), hi); }
__device__ __forceinline__ float glmod(float x, float y) { return x - y * floorf(x / y); }
__device__ __forceinline__ V3 reflect3(V3 I, V3 N) { return I - N * (dot(N, I) * 2.0f); }
__device__ __forceinline__ V3 refract3(V3 I, V3 N, float eta) {
    const float d = dot(N, I);
    const float k = 1.0f - eta * eta * (1.0f - d * d);
    if(!(k >= 0.0f)) return v3(0.0f, 0.0f, 0.0f);
    return eta * I - (eta * d + sqrtf(k)) * N;
}

// ------------------------------------------------------------------------------------------------ traversal
struct RayCtx {
    float ox, oy, oz, dx, dy, dz;
    float ix, iy, iz;     // 1/d (zero components clamped to +-1e-30)
    float Sx, Sy, Sz;     // Woop shear
    uint32_t octinv;      // bit 2/1/0 set when d.x/d.y/d.z >= 0
    int kx, ky, kz;
};

__device__ __forceinline__ float sel3(int k, float x, float y, float z) { return k == 0 ? x : (k == 1 ? y : z); }
__device__ __forceinline__ float __frcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

__device__ __forceinline__ void setupRay(RayCtx& r, float ox, float oy, float oz, float dx, float dy, float dz) {
    r.ox = ox; r.oy = oy; r.oz = oz; r.dx = dx; r.dy = dy; r.dz = dz;
    const float eps = 1e-30f;
    // approximate reciprocal (1 ulp): only the conservative slab test uses it; the slack below covers it
    r.ix = __frcp_approx(fabsf(dx) > eps ? dx : copysignf(eps, dx));
    r.iy = __frcp_approx(fabsf(dy) > eps ? dy : copysignf(eps, dy));
    r.iz = __frcp_approx(fabsf(dz) > eps ? dz : copysignf(eps, dz));
    r.octinv = (dx >= 0.0f ? 4u : 0u) | (dy >= 0.0f ? 2u : 0u) | (dz >= 0.0f ? 1u : 0u);
    const float ax = fabsf(dx), ay = fabsf(dy), az = fabsf(dz);
    const int kz = (ax > ay) ? ((ax > az) ? 0 : 2) : ((ay > az) ? 1 : 2);
    int kx = kz == 2 ? 0 : kz + 1;
    int ky = kx == 2 ? 0 : kx + 1;
    const float dkz = sel3(kz, dx, dy, dz);
    if(dkz < 0.0f) { const int t = kx; kx = ky; ky = t; }
    r.kx = kx; r.ky = ky; r.kz = kz;
    r.Sx = __fdiv_rn(sel3(kx, dx, dy, dz), dkz);
    r.Sy = __fdiv_rn(sel3(ky, dx, dy, dz), dkz);
    r.Sz = __frcp_rn(dkz);
}

template <int J>
__device__ __forceinline__ float byteF(uint32_t w) {  // 1 + b / 32768 for byte J of w: ONE PRMT puts b into mantissa bits 15..8 of 1.0f;
    return __uint_as_float(__byte_perm(w, 0x3F800000u, 0x7604u + (J << 4)));   // the affine map back to b is folded into the FFMA constants
}

// One child of a node: slab test on the quantised planes, then OR the child's bits into the hit mask.  Branch-free:
// bits4 / idx4 hold, per child byte, the bits to insert (0 for an empty slot) and where (see decodeMeta4).
template <int J>
__device__ __forceinline__ void childTest(uint32_t nx, uint32_t ny, uint32_t nz, uint32_t fx, uint32_t fy, uint32_t fz, uint32_t bits4, uint32_t idx4,
                                          float ax, float ay, float az, float onx, float ony, float onz, float ofx, float ofy, float ofz, float tmin,
                                          float tmax, uint32_t& hitmask) {
    const float tnx = fmaf(byteF<J>(nx), ax, onx), tny = fmaf(byteF<J>(ny), ay, ony), tnz = fmaf(byteF<J>(nz), az, onz);
    const float tfx = fmaf(byteF<J>(fx), ax, ofx), tfy = fmaf(byteF<J>(fy), ay, ofy), tfz = fmaf(byteF<J>(fz), az, ofz);
    const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tmin));
    const float tf = fminf(fminf(tfx, tfy), fminf(tfz, tmax));
    const uint32_t b = (bits4 >> (8 * J)) & 0xffu, i = (idx4 >> (8 * J)) & 0xffu;
    hitmask |= (tn <= tf) ? (b << i) : 0u;
}

// Four meta bytes at once (after Ylitie et al. 2017): inner children (bits 3 and 4 set) get bit index 24 + (slot ^ octinv),
// leaves keep their primitive offset; the bits to insert are meta >> 5 (1 for inner nodes, unary count for leaves).
__device__ __forceinline__ void decodeMeta4(uint32_t meta4, uint32_t octinv4, uint32_t& bits4, uint32_t& idx4) {
    const uint32_t isInner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
    const uint32_t innerMask4 = (isInner4 >> 4) * 0xffu;
    idx4 = (meta4 ^ (octinv4 & innerMask4)) & 0x1f1f1f1fu;
    bits4 = (meta4 >> 5) & 0x07070707u;
}

// Watertight ray / triangle test (Woop, Benthin, Wald 2013), no culling; operation order == oracle/orc_scene.cpp intersectTri.
__device__ __forceinline__ bool triTest(const RayCtx& r, const float4 p0, const float4 p1, const float4 p2, float tmin, float& tOut, float& uOut,
                                        float& vOut) {
    const float A0 = __fsub_rn(p0.x, r.ox), A1 = __fsub_rn(p0.y, r.oy), A2 = __fsub_rn(p0.z, r.oz);
    const float B0 = __fsub_rn(p1.x, r.ox), B1 = __fsub_rn(p1.y, r.oy), B2 = __fsub_rn(p1.z, r.oz);
    const float C0 = __fsub_rn(p2.x, r.ox), C1 = __fsub_rn(p2.y, r.oy), C2 = __fsub_rn(p2.z, r.oz);
    const float Akz = sel3(r.kz, A0, A1, A2), Bkz = sel3(r.kz, B0, B1, B2), Ckz = sel3(r.kz, C0, C1, C2);
    const float Ax = __fsub_rn(sel3(r.kx, A0, A1, A2), __fmul_rn(r.Sx, Akz)), Ay = __fsub_rn(sel3(r.ky, A0, A1, A2), __fmul_rn(r.Sy, Akz));
    const float Bx = __fsub_rn(sel3(r.kx, B0, B1, B2), __fmul_rn(r.Sx, Bkz)), By = __fsub_rn(sel3(r.ky, B0, B1, B2), __fmul_rn(r.Sy, Bkz));
    const float Cx = __fsub_rn(sel3(r.kx, C0, C1, C2), __fmul_rn(r.Sx, Ckz)), Cy = __fsub_rn(sel3(r.ky, C0, C1, C2), __fmul_rn(r.Sy, Ckz));
    float U = __fsub_rn(__fmul_rn(Cx, By), __fmul_rn(Cy, Bx));
    float V = __fsub_rn(__fmul_rn(Ax, Cy), __fmul_rn(Ay, Cx));
    float W = __fsub_rn(__fmul_rn(Bx, Ay), __fmul_rn(By, Ax));
    if(U == 0.0f || V == 0.0f || W == 0.0f) {  // rare: exact edge hit, redo in binary64 (products of floats are exact there)
        U = (float)__dsub_rn(__dmul_rn((double)Cx, (double)By), __dmul_rn((double)Cy, (double)Bx));
        V = (float)__dsub_rn(__dmul_rn((double)Ax, (double)Cy), __dmul_rn((double)Ay, (double)Cx));
        W = (float)__dsub_rn(__dmul_rn((double)Bx, (double)Ay), __dmul_rn((double)By, (double)Ax));
    }
    if((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    const float det = __fadd_rn(__fadd_rn(U, V), W);
    if(det == 0.0f) return false;
    const float Az = __fmul_rn(r.Sz, Akz), Bz = __fmul_rn(r.Sz, Bkz), Cz = __fmul_rn(r.Sz, Ckz);
    const float T = __fadd_rn(__fadd_rn(__fmul_rn(U, Az), __fmul_rn(V, Bz)), __fmul_rn(W, Cz));
    const float rcp = __frcp_rn(det);
    const float t = __fmul_rn(T, rcp);
    if(!(t > tmin)) return false;
    tOut = t; uOut = __fmul_rn(V, rcp); vOut = __fmul_rn(W, rcp);
    return true;
}

// Closest hit in [tmin, tmax] (exclusive).  hit.inst == kInvalid on miss.
template <bool COUNT>
__device__ __forceinline__ void traverse(const TraceParams& P, float ox, float oy, float oz, float dx, float dy, float dz, float tmin, float tmax,
                                         Hit& hit, uint32_t* cnt) {
    hit.t = tmax; hit.u = 0.0f; hit.v = 0.0f; hit.inst = kInvalid; hit.prim = kInvalid;
    if(P.nInst == 0) return;
    if(dx == 0.0f && dy == 0.0f && dz == 0.0f) return;       // zero direction (refract on total internal reflection): miss
    if(!(dx == dx && dy == dy && dz == dz && ox == ox && oy == oy && oz == oz)) return;  // NaN ray: miss

    RayCtx r;
    setupRay(r, ox, oy, oz, dx, dy, dz);
    uint2 stack[kStackSize];
    int sp = 0;
    uint2 ng = make_uint2(0u, 0x80000000u);
    uint2 tg = make_uint2(0u, 0u);
    bool inBlas = false;
    uint32_t curInst = kInvalid;
    const Node8* nodes = P.tlasNodes;

    while(true) {
#if RG_TRAVERSE_IFIF
        if((ng.y & 0xff000000u) && !tg.y) {   // one node per iteration, and only once this lane's pending primitives are done
#else
        if(ng.y & 0xff000000u) {
#endif
            const int bit = 31 - __clz(ng.y);
            ng.y &= ~(1u << bit);
            const uint32_t slot = (uint32_t)(bit - 24) ^ r.octinv;
            const uint32_t rel = __popc(ng.y & 0xffu & ((1u << slot) - 1u));
            if((ng.y & 0xff000000u) && sp < kStackSize) stack[sp++] = ng;
            const uint4* np = reinterpret_cast<const uint4*>(nodes + (ng.x + rel));
            const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
            if(COUNT) cnt[CNT_NODES]++;
            const float sx = __uint_as_float((n0.w & 0xffu) << 23), sy = __uint_as_float(((n0.w >> 8) & 0xffu) << 23),
                        sz = __uint_as_float(((n0.w >> 16) & 0xffu) << 23);
            const float px = __uint_as_float(n0.x) - r.ox, py = __uint_as_float(n0.y) - r.oy, pz = __uint_as_float(n0.z) - r.oz;
            // plane distance t = q * (2^e / d) + (p - o) / d.  The second term cancels against the first, so its rounding
            // error (relative to |(p - o) / d|, NOT to t) is what can make the slab test miss: widen near / far by that much.
            // Near and far use the same q * adj, so a flat child box (qlo == qhi) always keeps near <= far.
            // byteF gives v = 1 + q / 32768, so t = v * A + (b - A) with A = 32768 * 2^e / d.  Rounding: ulp(A) = 1/256 of one
            // quantisation step in t, plus the error of b = (p - o) / d; both are covered by the slack (relative to |b| and to a step).
            const float ax = sx * r.ix * 32768.0f, ay = sy * r.iy * 32768.0f, az = sz * r.iz * 32768.0f;
            const float bx = px * r.ix, by = py * r.iy, bz = pz * r.iz;
            const float kSlack = 7.2e-7f, kStep = 1.0f / (32768.0f * 64.0f);   // 1/64 of a quantisation step
            const float wx = fmaf(fabsf(bx), kSlack, fabsf(ax) * kStep), wy = fmaf(fabsf(by), kSlack, fabsf(ay) * kStep), wz = fmaf(fabsf(bz), kSlack, fabsf(az) * kStep);
            const float onx = (bx - ax) - wx, ony = (by - ay) - wy, onz = (bz - az) - wz;
            const float ofx = (bx - ax) + wx, ofy = (by - ay) + wy, ofz = (bz - az) + wz;
            const bool negx = r.dx < 0.0f, negy = r.dy < 0.0f, negz = r.dz < 0.0f;
            const uint32_t octinv4 = r.octinv * 0x01010101u;
            uint32_t hitmask = 0;
#pragma unroll 1
            for(int half = 0; half < 2; ++half) {   // slots 0..3, then 4..7: a rolled loop keeps the hot code small for the instruction cache
                const uint32_t lx = half ? n2.y : n2.x, ly = half ? n2.w : n2.z, lz = half ? n3.y : n3.x;
                const uint32_t hx = half ? n3.w : n3.z, hy = half ? n4.y : n4.x, hz = half ? n4.w : n4.z;
                const uint32_t nx = negx ? hx : lx, fx = negx ? lx : hx, ny = negy ? hy : ly, fy = negy ? ly : hy, nz = negz ? hz : lz, fz = negz ? lz : hz;
                uint32_t bits4, idx4;
                decodeMeta4(half ? n1.w : n1.z, octinv4, bits4, idx4);
                childTest<0>(nx, ny, nz, fx, fy, fz, bits4, idx4, ax, ay, az, onx, ony, onz, ofx, ofy, ofz, tmin, hit.t, hitmask);
                childTest<1>(nx, ny, nz, fx, fy, fz, bits4, idx4, ax, ay, az, onx, ony, onz, ofx, ofy, ofz, tmin, hit.t, hitmask);
                childTest<2>(nx, ny, nz, fx, fy, fz, bits4, idx4, ax, ay, az, onx, ony, onz, ofx, ofy, ofz, tmin, hit.t, hitmask);
                childTest<3>(nx, ny, nz, fx, fy, fz, bits4, idx4, ax, ay, az, onx, ony, onz, ofx, ofy, ofz, tmin, hit.t, hitmask);
            }
            ng = make_uint2(n1.x, (hitmask & 0xff000000u) | (n0.w >> 24));
            tg = make_uint2(n1.y, hitmask & 0x00ffffffu);
        }
#if !RG_TRAVERSE_IFIF
        else {
            tg = ng;
            ng = make_uint2(0u, 0u);
        }
#endif

#if RG_TRAVERSE_IFIF
        if(tg.y) {      // ONE primitive per iteration: lanes without pending primitives go on with their next node meanwhile
#else
        while(tg.y) {
#endif
            const int bit = __ffs(tg.y) - 1;
            tg.y &= tg.y - 1u;
            if(!inBlas) {
                const uint4* lp = reinterpret_cast<const uint4*>(P.tlasLeaves + (tg.x + bit));
                const uint4 l3 = __ldg(lp + 3);
                if(l3.x == kInvalid) continue;  // instance of an empty mesh
                if(sp + 6 > kStackSize) continue;  // stack exhausted: skip (never with sane scenes)
                if(COUNT) cnt[CNT_INST]++;
                if(tg.y) stack[sp++] = tg;
                if(ng.y & 0xff000000u) stack[sp++] = ng;
                if(l3.z) {
                    // pure translation (flagged by k_prepare_instances): the direction and everything derived from it stay; the
                    // oracle's ((1*ox + 0*oy) + 0*oz) + t is exactly ox + t
                    const uint4 l0 = __ldg(lp), l1 = __ldg(lp + 1), l2 = __ldg(lp + 2);
                    stack[sp++] = make_uint2(kInvalid, 0x1000u);
                    r.ox = __fadd_rn(ox, __uint_as_float(l0.w)); r.oy = __fadd_rn(oy, __uint_as_float(l1.w)); r.oz = __fadd_rn(oz, __uint_as_float(l2.w));
                } else {
                    const uint4 l0 = __ldg(lp), l1 = __ldg(lp + 1), l2 = __ldg(lp + 2);
                    // the world-space slab / shear constants ride on the stack while the instance is traversed
                    stack[sp++] = make_uint2(__float_as_uint(r.ix), __float_as_uint(r.iy));
                    stack[sp++] = make_uint2(__float_as_uint(r.iz), __float_as_uint(r.Sx));
                    stack[sp++] = make_uint2(__float_as_uint(r.Sy), __float_as_uint(r.Sz));
                    stack[sp++] = make_uint2(kInvalid, r.octinv | ((uint32_t)r.kx << 4) | ((uint32_t)r.ky << 6) | ((uint32_t)r.kz << 8));
                    // object-space ray: same operation order as the oracle (t is preserved, direction not normalised)
                    const float* w0 = reinterpret_cast<const float*>(&l0); const float* w1 = reinterpret_cast<const float*>(&l1);
                    const float* w2 = reinterpret_cast<const float*>(&l2);
                    const float oox = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w0[0], ox), __fmul_rn(w0[1], oy)), __fmul_rn(w0[2], oz)), w0[3]);
                    const float ooy = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w1[0], ox), __fmul_rn(w1[1], oy)), __fmul_rn(w1[2], oz)), w1[3]);
                    const float ooz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w2[0], ox), __fmul_rn(w2[1], oy)), __fmul_rn(w2[2], oz)), w2[3]);
                    const float odx = __fadd_rn(__fadd_rn(__fmul_rn(w0[0], dx), __fmul_rn(w0[1], dy)), __fmul_rn(w0[2], dz));
                    const float ody = __fadd_rn(__fadd_rn(__fmul_rn(w1[0], dx), __fmul_rn(w1[1], dy)), __fmul_rn(w1[2], dz));
                    const float odz = __fadd_rn(__fadd_rn(__fmul_rn(w2[0], dx), __fmul_rn(w2[1], dy)), __fmul_rn(w2[2], dz));
                    if(odx == 0.0f && ody == 0.0f && odz == 0.0f) { sp -= 4; if(ng.y & 0xff000000u) sp -= 1; if(tg.y) sp -= 1; continue; }
                    setupRay(r, oox, ooy, ooz, odx, ody, odz);
                }
                curInst = l3.y;
                inBlas = true;
                nodes = P.blasNodes;
                ng = make_uint2(l3.x, 0x80000000u);
                tg = make_uint2(0u, 0u);
#if !RG_TRAVERSE_IFIF
                break;
#endif
            } else {
#if RG_POSTPONE_THRESHOLD > 0 && !RG_TRAVERSE_IFIF
                // too few lanes of this warp are in the triangle loop and this lane still has child nodes to visit: put the
                // triangle group back (it goes to the stack) and test it later together with more lanes (after Ylitie et al. 2017)
                if((ng.y & 0xff000000u) && sp < kStackSize && __popc(__activemask()) < RG_POSTPONE_THRESHOLD) {
                    tg.y |= 1u << bit;
                    stack[sp++] = tg;
                    tg.y = 0u;
                    break;
                }
#endif
                const float4* tp = reinterpret_cast<const float4*>(P.tris + (tg.x + bit));
                const float4 p0 = __ldg(tp), p1 = __ldg(tp + 1), p2 = __ldg(tp + 2);
                if(COUNT) cnt[CNT_TRIS]++;
                float t, u, v;
                if(triTest(r, p0, p1, p2, tmin, t, u, v)) {
                    const uint32_t prim = __float_as_uint(p0.w);
                    if(t < hit.t || (t == hit.t && t < tmax && (curInst < hit.inst || (curInst == hit.inst && prim < hit.prim)))) {
                        hit.t = t; hit.u = u; hit.v = v; hit.inst = curInst; hit.prim = prim;
                    }
                }
            }
        }

#if RG_TRAVERSE_IFIF
        if(!(ng.y & 0xff000000u) && !tg.y) {
#else
        if(!(ng.y & 0xff000000u)) {
#endif
            bool done = false;
            while(true) {
                if(sp == 0) { done = true; break; }
                ng = stack[--sp];
                if(ng.x == kInvalid) {  // leave the instance: back to the world-space ray
                    inBlas = false; nodes = P.tlasNodes;
                    r.ox = ox; r.oy = oy; r.oz = oz;
                    if(!(ng.y & 0x1000u)) {   // a general instance: direction-derived constants come back from the stack
                        r.dx = dx; r.dy = dy; r.dz = dz;
                        r.octinv = ng.y & 7u; r.kx = (int)((ng.y >> 4) & 3u); r.ky = (int)((ng.y >> 6) & 3u); r.kz = (int)((ng.y >> 8) & 3u);
                        const uint2 c2 = stack[--sp], c1 = stack[--sp], c0 = stack[--sp];
                        r.ix = __uint_as_float(c0.x); r.iy = __uint_as_float(c0.y); r.iz = __uint_as_float(c1.x);
                        r.Sx = __uint_as_float(c1.y); r.Sy = __uint_as_float(c2.x); r.Sz = __uint_as_float(c2.y);
                    }
                    continue;
                }
                break;
            }
            if(done) break;
#if RG_TRAVERSE_IFIF
            if(!(ng.y & 0xff000000u)) { tg = ng; ng = make_uint2(0u, 0u); }   // a primitive group came off the stack
#endif
        }
    }
}

// ------------------------------------------------------------------------------------------------ shading
// raygen.h:37-67
__constant__ float2 c_aaOffsets[3][8] = {
    {{0.25f, 0.25f}, {-0.25f, -0.25f}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}},
    {{-0.125f, -0.375f}, {0.375f, -0.125f}, {-0.375f, 0.125f}, {0.125f, 0.375f}, {0, 0}, {0, 0}, {0, 0}, {0, 0}},
    {{0.0625f, -0.1875f}, {-0.0625f, 0.1875f}, {0.3125f, 0.0625f}, {-0.1875f, -0.3125f}, {-0.3125f, 0.3125f}, {-0.4375f, -0.0625f}, {0.1875f, 0.4375f}, {0.4375f, -0.4375f}},
};
__device__ __forceinline__ float2 aaOffset(int numSamples, int i) {
    const int n = numSamples < 8 ? numSamples : 8;
    if(n < 2) return make_float2(0.0f, 0.0f);
    return c_aaOffsets[n == 2 ? 0 : (n <= 4 ? 1 : 2)][i & 7];
}

// miss.rmiss:38-74 with the constants of :78; normalize(0) is kept 0 (SURVEY hazard 7)
__device__ __noinline__ V3 skyColor(V3 d, V3 lightDir, bool strict) {
    const bool zero = d.x == 0.0f && d.y == 0.0f && d.z == 0.0f;
    const V3 rayDir = (zero && !strict) ? v3(0, 0, 0) : normalize(d);
    const float y = fabsf(d.y + 1.5f) / 3.0f;
    const V3 sd = normalize(-lightDir) - rayDir;
    float sun = 1.0f - sqrtf(dot(sd, sd));
    sun = clampf(sun, 0.0f, 2.0f);
    float glow = clampf(sun, 0.0f, 1.0f);
    sun = powf(sun, 80.0f);
    sun *= 1000.0f;
    sun = clampf(sun, 0.0f, 16.0f);
    glow = powf(glow, 6.0f) * 1.0f;
    glow = powf(glow, y);
    glow = clampf(glow, 0.0f, 1.0f);
    sun *= powf(y * y, 1.0f / 1.65f);
    glow *= powf(y * y, 1.0f / 2.0f);
    sun += glow;
    const V3 sunColor = v3(1.0f, 0.6f, 0.05f) * sun;
    const float atmosphere = sqrtf(1.0f - y);
    float scatter = powf(4.0f - lightDir.y, 1.0f / 15.0f);
    scatter = 1.0f - clampf(scatter, 0.8f, 1.0f);
    const V3 scatterColor = mix3(v3(1.0f, 1.0f, 1.0f), v3(1.0f, 0.3f, 0.0f) * 1.5f, scatter);
    const V3 skyScatter = mix3(v3(0.2f, 0.4f, 0.8f), scatterColor, atmosphere / 1.3f);
    return sunColor + skyScatter;
}

// frame layout (words) in local memory
enum {
    F_ORG = 0, F_DIR = 3, F_N = 6, F_T = 9, F_DIFF = 10, F_SPEC = 13, F_TRANSP = 16, F_REFL = 17, F_ROUGH = 18, F_IOR = 19, F_EMIS = 20, F_FLAGS = 21,
    F_BASE = 22, F_RCOL = 25, F_RDEPTH = 28, F_RECDEPTH = 29, F_WORDS = 30
};
enum { FR_GEN = 0, FR_SHI = 1 };
enum { ST_SHADOW_RET = 0, ST_TRY_REFLECT = 1, ST_REFLECT_RET = 2, ST_TRY_REFRACT = 3, ST_REFRACT_RET_FRONT = 4, ST_REFRACT_RET_BACK = 5, ST_COMBINE = 6 };

__device__ __forceinline__ uint32_t f2h(float f) { return (uint32_t)__half_as_ushort(__float2half_rn(f)); }
__device__ __forceinline__ uint2 packHalf4(float x, float y, float z, float w) { return make_uint2(f2h(x) | (f2h(y) << 16), f2h(z) | (f2h(w) << 16)); }

#ifndef RG_POSTPONE_THRESHOLD
#define RG_POSTPONE_THRESHOLD 12
#endif
#ifndef RG_TRACE_MIN_BLOCKS
#define RG_TRACE_MIN_BLOCKS 8
#endif
// K = ray contexts per lane.  K == 1: the lane's state lives in registers.  K > 1: contexts live in local memory and every
// lane traverses its K pending rays back to back, so a warp waits for the slowest SUM of K rays rather than for the slowest
// single ray -- the cure for warps that idle on incoherent bounces (sphere-grid config).
template <bool COUNT, bool MULTI, int K>
__global__ void __launch_bounds__(128, RG_TRACE_MIN_BLOCKS) k_trace(const TraceParams P) {
    __shared__ float s_ubo[48];
    if(threadIdx.x < 48) s_ubo[threadIdx.x] = P.ubo[threadIdx.x];
    __syncthreads();
    const float* VI = s_ubo;       // viewInverse, column-major
    const float* PI = s_ubo + 16;  // projInverse
    const int numSamples = __float_as_int(s_ubo[35]);
    const V3 L = v3(s_ubo[36], s_ubo[37], s_ubo[38]);
    int maxRec = __float_as_int(s_ubo[39]);
    maxRec = maxRec > kMaxRecursions ? kMaxRecursions : maxRec;
    const bool strictIeee = (P.flags & RG_STRICT_IEEE) != 0;

    const uint32_t lane = threadIdx.x & 31;
    const uint32_t tilesX = (P.dw + 7) / 8, tilesY = (P.dh + 3) / 4;
    const uint32_t nTiles = tilesX * tilesY;
    // this rank's share: chunks of kChunkTiles tiles dealt round-robin (world == 1: everything)
    const uint32_t nChunks = (nTiles + kChunkTiles - 1) / kChunkTiles;
    const uint32_t myChunks = nChunks > P.rank ? (nChunks - P.rank + P.world - 1) / P.world : 0u;
    const uint32_t S = (uint32_t)numSamples;
    const uint32_t total = myChunks * kChunkTiles * 32u * S;   // one work item per pixel SAMPLE

    enum { CS_IDLE = 0, CS_RAY = 1, CS_HIT = 2 };
    float frames[K][kMaxFrames][F_WORDS];   // suspended shader invocations, per context
    float cRay[K][8], cPay[K][14];          // pending ray; payload (hv, depth, curIOR, refDepth, normal, rough, roughA, contrib)
    uint32_t cHit[K][5], cPix[K][6];        // closest hit; lx, ly, slot, pslot, sampleRays, sample
    int cSel[K], cStatus[K];                // rayType | missIndex << 2 | rayKind << 4 | recDepth << 8 | sp << 16; CS_*
#pragma unroll
    for(int k = 0; k < K; ++k) cStatus[k] = CS_IDLE;
    uint32_t cnt[CNT_N];
#pragma unroll
    for(int k = 0; k < CNT_N; ++k) cnt[k] = 0;

    // per-pixel state
    bool exhausted = false;
    uint32_t lx = 0, ly = 0;   // frame coordinates of the lane's pixel
    int sample = 0;
    uint32_t slot = 0, pslot = 0, sampleRays = 0;   // tile slot in this rank's share, pixel slot (slot * 32 + lane in tile), rays of this sample
    // payload (payload.h:29-39)
    V3 hv = v3(0, 0, 0), pNormal = v3(0, 0, 0), pRough = v3(0, 0, 0);
    float pRoughA = 0, pContrib = 0, depth = 0, curIOR = 1.0f, refDepth = 0;
    int recDepth = 0;
    int sp = 0;  // frames
    // pending ray
    V3 ro = v3(0, 0, 0), rd = v3(0, 0, 0);
    float rtmin = 0, rtmax = 0;
    int rayType = RT_GENERIC, missIndex = 0, rayKind = CNT_PRIMARY;
    // camera origin: viewInverse * (0,0,0,1) in GLM order (c0*0 + c1*0) + (c2*0 + c3*1)
    const V3 camO = v3((VI[0] * 0.0f + VI[4] * 0.0f) + (VI[8] * 0.0f + VI[12] * 1.0f), (VI[1] * 0.0f + VI[5] * 0.0f) + (VI[9] * 0.0f + VI[13] * 1.0f),
                       (VI[2] * 0.0f + VI[6] * 0.0f) + (VI[10] * 0.0f + VI[14] * 1.0f));

    auto primaryRay = [&](int i) {
        const float2 off = aaOffset(numSamples, i);
        const float pcx = (float)lx + 0.5f + off.x, pcy = (float)ly + 0.5f + off.y;
        const float ddx = __fdiv_rn(pcx, (float)P.W) * 2.0f - 1.0f, ddy = __fdiv_rn(pcy, (float)P.H) * 2.0f - 1.0f;
        // target = projInverse * (d.x, d.y, 1, 1); direction = viewInverse * (normalize(target.xyz), 0)
        const V3 tgt = v3((PI[0] * ddx + PI[4] * ddy) + (PI[8] + PI[12]), (PI[1] * ddx + PI[5] * ddy) + (PI[9] + PI[13]),
                          (PI[2] * ddx + PI[6] * ddy) + (PI[10] + PI[14]));
        const V3 nt = normalize(tgt);
        rd = v3((VI[0] * nt.x + VI[4] * nt.y) + (VI[8] * nt.z), (VI[1] * nt.x + VI[5] * nt.y) + (VI[9] * nt.z), (VI[2] * nt.x + VI[6] * nt.y) + (VI[10] * nt.z));
        ro = camO; rtmin = 0.001f; rtmax = 10000.0f;
        rayType = RT_GENERIC; missIndex = 0; rayKind = CNT_PRIMARY;
        hv = v3(0, 0, 0); pNormal = v3(0, 0, 0); pRough = v3(0, 0, 0); pRoughA = 0; pContrib = 0;
        depth = 0; refDepth = 0; curIOR = 1.0f; recDepth = 0; sp = 0;
    };

    auto storeCtx = [&](int k) {
        cRay[k][0] = ro.x; cRay[k][1] = ro.y; cRay[k][2] = ro.z; cRay[k][3] = rd.x; cRay[k][4] = rd.y; cRay[k][5] = rd.z; cRay[k][6] = rtmin; cRay[k][7] = rtmax;
        cPay[k][0] = hv.x; cPay[k][1] = hv.y; cPay[k][2] = hv.z; cPay[k][3] = depth; cPay[k][4] = curIOR; cPay[k][5] = refDepth;
        cPay[k][6] = pNormal.x; cPay[k][7] = pNormal.y; cPay[k][8] = pNormal.z; cPay[k][9] = pRough.x; cPay[k][10] = pRough.y; cPay[k][11] = pRough.z;
        cPay[k][12] = pRoughA; cPay[k][13] = pContrib;
        cPix[k][0] = lx; cPix[k][1] = ly; cPix[k][2] = slot; cPix[k][3] = pslot; cPix[k][4] = sampleRays; cPix[k][5] = (uint32_t)sample;
        cSel[k] = rayType | (missIndex << 2) | (rayKind << 4) | (recDepth << 8) | (sp << 16);
    };
    auto loadCtx = [&](int k) {
        ro = v3(cRay[k][0], cRay[k][1], cRay[k][2]); rd = v3(cRay[k][3], cRay[k][4], cRay[k][5]); rtmin = cRay[k][6]; rtmax = cRay[k][7];
        hv = v3(cPay[k][0], cPay[k][1], cPay[k][2]); depth = cPay[k][3]; curIOR = cPay[k][4]; refDepth = cPay[k][5];
        pNormal = v3(cPay[k][6], cPay[k][7], cPay[k][8]); pRough = v3(cPay[k][9], cPay[k][10], cPay[k][11]); pRoughA = cPay[k][12]; pContrib = cPay[k][13];
        lx = cPix[k][0]; ly = cPix[k][1]; slot = cPix[k][2]; pslot = cPix[k][3]; sampleRays = cPix[k][4]; sample = (int)cPix[k][5];
        const int sel = cSel[k];
        rayType = sel & 3; missIndex = (sel >> 2) & 3; rayKind = (sel >> 4) & 15; recDepth = (sel >> 8) & 255; sp = (sel >> 16) & 255;
    };

    while(true) {
        // ---- refill idle contexts (warp vote + prefix compaction over one atomic per context index)
        bool mine = false;
#pragma unroll 1
        for(int k = 0; k < K; ++k) {
            const uint32_t idle = __ballot_sync(0xffffffffu, cStatus[k] == CS_IDLE);
            if(idle && !exhausted && (idle == 0xffffffffu || __popc(idle) >= 8)) {
                const int n = __popc(idle), leader = __ffs(idle) - 1;
                uint32_t basew = 0;
                if((int)lane == leader) basew = atomicAdd(P.workCounter, (uint32_t)n);
                basew = __shfl_sync(0xffffffffu, basew, leader);
                if(cStatus[k] == CS_IDLE) {
                    const uint32_t w = basew + __popc(idle & ((1u << lane) - 1u));
                    if(w < total) {
                        // w = ((slot position * S) + sample) * 32 + lane: a warp works on one sample index of one 8x4 tile
                        const uint32_t l = w & 31u, pos = (w >> 5) / S;
                        sample = (int)((w >> 5) % S);
                        const uint32_t j = P.tileOrder ? __ldg(P.tileOrder + pos) : pos;
                        slot = j; pslot = j * 32u + l; sampleRays = 0;
                        const uint32_t tile = ((j / kChunkTiles) * P.world + P.rank) * kChunkTiles + (j % kChunkTiles);
                        lx = P.dx0 + (tile % tilesX) * 8u + (l & 7u); ly = P.dy0 + (tile / tilesX) * 4u + (l >> 3);   // frame coordinates
                        if(tile < nTiles && lx < P.dx0 + P.dw && ly < P.dy0 + P.dh) {
                            primaryRay(sample);
                            storeCtx(k);
                            cStatus[k] = CS_RAY;
                        }
                    }
                }
                if(basew + (uint32_t)n >= total) exhausted = true;
            }
            mine |= cStatus[k] != CS_IDLE;
        }
        if(!__any_sync(0xffffffffu, mine)) { if(exhausted) break; continue; }

        // ---- trace every pending ray of this lane, one after the other
#pragma unroll 1
        for(int k = 0; k < K; ++k) {
            if(cStatus[k] != CS_RAY) continue;
            Hit hit;
            const int kind = (cSel[k] >> 4) & 15;
            cnt[kind]++;
            cPix[k][4]++;
            traverse<COUNT>(P, cRay[k][0], cRay[k][1], cRay[k][2], cRay[k][3], cRay[k][4], cRay[k][5], cRay[k][6], cRay[k][7], hit, cnt);
            cHit[k][0] = __float_as_uint(hit.t); cHit[k][1] = __float_as_uint(hit.u); cHit[k][2] = __float_as_uint(hit.v); cHit[k][3] = hit.inst; cHit[k][4] = hit.prim;
            if(kind == CNT_PRIMARY && cPix[k][5] == 0u && P.idInst) {
                const int qx = (int)cPix[k][0] - P.sx0, qy = (int)cPix[k][1] - P.sy0;
                if(qx >= 0 && qy >= 0 && qx < P.sw && qy < P.sh) { P.idInst[qy * P.sw + qx] = hit.inst; P.idPrim[qy * P.sw + qx] = hit.prim; }
            }
            cStatus[k] = CS_HIT;
        }

        // ---- shade every finished ray: context index by context index, so the lanes of the warp run the hit / miss programs together
#pragma unroll 1
        for(int k = 0; k < K; ++k) {
            if(cStatus[k] != CS_HIT) continue;
            loadCtx(k);
            Hit hit;
            hit.t = __uint_as_float(cHit[k][0]); hit.u = __uint_as_float(cHit[k][1]); hit.v = __uint_as_float(cHit[k][2]); hit.inst = cHit[k][3]; hit.prim = cHit[k][4];
            const bool found = hit.inst != kInvalid;
            float (*fr)[F_WORDS] = frames[k];
            bool active = true;
                // ---- shade: closest hit or miss, then resume suspended frames until a new ray is issued
                bool issue = false;   // a new ray is pending
                if(found) {
                    // closesthit.rchit:96-109
                    const uint4 is3 = __ldg(reinterpret_cast<const uint4*>(P.instShade + hit.inst) + 3);
                    const uint32_t vtxOff = is3.x, idxOff = is3.y, matOff = is3.z;
                    const uint32_t i0 = __ldg(P.indices + idxOff + 3 * hit.prim), i1 = __ldg(P.indices + idxOff + 3 * hit.prim + 1),
                                   i2 = __ldg(P.indices + idxOff + 3 * hit.prim + 2);
                    const float4 v0p = __ldg(P.vertices + 2 * (size_t)(vtxOff + i0));
                    const float4 n0 = __ldg(P.vertices + 2 * (size_t)(vtxOff + i0) + 1), n1 = __ldg(P.vertices + 2 * (size_t)(vtxOff + i1) + 1),
                                 n2 = __ldg(P.vertices + 2 * (size_t)(vtxOff + i2) + 1);
                    const float4* mp = P.materials + 4 * (size_t)(matOff + __float_as_uint(v0p.w));
                    const float4 m0 = __ldg(mp), m1 = __ldg(mp + 1), m2 = __ldg(mp + 2), m3 = __ldg(mp + 3);
                    V3 diffuse = v3(m0.x, m0.y, m0.z), specular = v3(m1.x, m1.y, m1.z);
                    const float transparency = m0.w; float reflectivity = m1.w;
                    const float roughness = m2.x, ior = m2.y, emission = m3.x;
                    const uint32_t effectId = __float_as_uint(m2.z), rayConsumption = __float_as_uint(m2.w);

                    const float b0 = 1.0f - hit.u - hit.v;
                    const V3 origin = ro + rd * hit.t;                                        // :114
                    const V3 vn = v3(n0.x, n0.y, n0.z) * b0 + v3(n1.x, n1.y, n1.z) * hit.u + v3(n2.x, n2.y, n2.z) * hit.v;  // :117
                    const float4* ow = reinterpret_cast<const float4*>(P.instShade + hit.inst);
                    const float4 o0 = __ldg(ow), o1 = __ldg(ow + 1), o2 = __ldg(ow + 2);
                    V3 n = normalize(v3(o0.x * vn.x + o0.y * vn.y + o0.z * vn.z, o1.x * vn.x + o1.y * vn.y + o1.z * vn.z, o2.x * vn.x + o2.y * vn.y + o2.z * vn.z));  // :118-119

                    if(effectId == 1u) {  // gridEffect, :74-91
                        const float aa = (refDepth + hit.t + 8.0f) / 30.0f;
                        const float aa2 = aa / 2.0f;
                        float minmod = glmin(fabsf(glmod((origin.x + 1000.0f) * 10.0f + aa2, 20.0f) - aa2), fabsf(glmod((origin.z + 1000.0f) * 10.0f + aa2, 20.0f) - aa2));
                        if(minmod < aa2) {
                            minmod -= aa2 - (aa * aa) / 3.0f;
                            minmod *= 3.0f / (aa * aa);
                            const float f = mixf(aa / 10.0f, 1.0f, minmod);
                            diffuse = diffuse * f; specular = specular * f; reflectivity *= f;
                        }
                        if(glmod((origin.x + 1000.0f) * 5.0f, 20.0f) < 10.0f && glmod((origin.z + 1000.0f) * 5.0f, 20.0f) < 10.0f) reflectivity *= 1.5f;
                    }

                    if(rayType == RT_SHADOW_INTERNAL) {  // :125-152
                        const float thick = clampf(hit.t * (1.0f - transparency) * 10.0f, 0.0f, 1.0f);
                        const V3 nd = normalize(v3(1.1f - diffuse.x, 1.1f - diffuse.y, 1.1f - diffuse.z));
                        const V3 shadowCol = hv - mix3(v3(0, 0, 0), v3(nd.x + 0.1f, nd.y + 0.1f, nd.z + 0.1f), thick);
                        if(recDepth < maxRec) {
                            float* f = fr[sp++];
                            f[F_ORG] = shadowCol.x; f[F_ORG + 1] = shadowCol.y; f[F_ORG + 2] = shadowCol.z;
                            f[F_DIR] = rd.x; f[F_DIR + 1] = rd.y; f[F_DIR + 2] = rd.z;
                            f[F_N] = n.x; f[F_N + 1] = n.y; f[F_N + 2] = n.z;
                            f[F_IOR] = ior; f[F_FLAGS] = __int_as_float(FR_SHI); f[F_RECDEPTH] = __int_as_float(recDepth);
                            rayType = RT_SHADOW_TRACE; recDepth++;
                            ro = origin; rtmin = 0.01f; rtmax = 1000.0f; missIndex = 0; rayKind = CNT_SHADOW;   // rd unchanged (T3)
                            issue = true;
                        } else {
                            hv = shadowCol * 0.4f;
                        }
                    } else if(rayType == RT_SHADOW_TRACE) {  // :153-166
                        if(transparency > 0.0f) {
                            if(recDepth < maxRec) {   // T2: nothing to do after the child returns except recDepth--, which every
                                rayType = RT_SHADOW_INTERNAL; recDepth++;                 // resuming frame restores from its own copy
                                ro = origin; rtmin = 0.01f; rtmax = 1000.0f; missIndex = 1; rayKind = CNT_SHADOW;
                                issue = true;
                            }
                        } else {
                            hv = hv * mixf(0.4f, 0.8f, clampf(logf(hit.t) / 8.0f, 0.0f, 1.0f));
                        }
                    } else {  // RT_GENERIC, :168-268
                        if(COUNT) cnt[CNT_GENHIT]++;
                        const bool frontFacing = dot(-rd, n) > 0.0f;
                        if(!frontFacing) n = normalize(-n);
                        const float ndl = dot(-L, n);
                        V3 baseColor = diffuse * glmax(ndl, 0.2f);
                        float* f = fr[sp++];
                        f[F_ORG] = origin.x; f[F_ORG + 1] = origin.y; f[F_ORG + 2] = origin.z;
                        f[F_DIR] = rd.x; f[F_DIR + 1] = rd.y; f[F_DIR + 2] = rd.z;
                        f[F_N] = n.x; f[F_N + 1] = n.y; f[F_N + 2] = n.z;
                        f[F_T] = hit.t;
                        f[F_DIFF] = diffuse.x; f[F_DIFF + 1] = diffuse.y; f[F_DIFF + 2] = diffuse.z;
                        f[F_SPEC] = specular.x; f[F_SPEC + 1] = specular.y; f[F_SPEC + 2] = specular.z;
                        f[F_TRANSP] = transparency; f[F_REFL] = reflectivity; f[F_ROUGH] = roughness; f[F_IOR] = ior; f[F_EMIS] = emission;
                        f[F_RECDEPTH] = __int_as_float(recDepth);
                        int stage;
                        bool shadowPending = false;
                        if(ndl > 0.07f) {   // :188
                            if(recDepth < maxRec) {
                                hv = v3(1.0f, 1.0f, 1.0f);
                                rayType = RT_SHADOW_TRACE; recDepth++;
                                ro = origin; rd = -L; rtmin = 0.1f; rtmax = 1000.0f; missIndex = 1; rayKind = CNT_SHADOW;
                                shadowPending = true;
                            }
                            // recDepth >= max: shadowColor stays 1
                        } else {
                            const float sm = transparency * transparency;   // pow(t, 2)
                            const V3 sc = transparency < 1.0f ? mix3(v3(1, 1, 1), diffuse * sm, transparency) : v3(0.4f, 0.4f, 0.4f);
                            baseColor = baseColor * sc;
                        }
                        if(shadowPending) { stage = ST_SHADOW_RET; issue = true; }
                        else { baseColor = baseColor + diffuse * emission; stage = ST_TRY_REFLECT; }
                        f[F_BASE] = baseColor.x; f[F_BASE + 1] = baseColor.y; f[F_BASE + 2] = baseColor.z;
                        f[F_FLAGS] = __int_as_float(FR_GEN | (stage << 8) | ((frontFacing ? 1 : 0) << 16) | ((int)(rayConsumption & 0xffu) << 20));
                    }
                } else {
                    if(missIndex == 0) {  // miss.rmiss:76-83
                        const V3 sky = skyColor(rd, L, strictIeee);
                        hv = sky; depth = 10000.0f;
                        if(sp == 0 && recDepth == 0) { pRough = sky; pRoughA = 0.0f; }   // roughValue is only observable for a primary miss
                    } else {              // shadowMiss.rmiss:33
                        hv = v3(1.0f, 1.0f, 1.0f);
                    }
                }

                // ---- resume suspended frames (the code after each traceRayEXT returns)
                while(!issue) {
                    if(sp == 0) {   // raygen.h:105-111: the sample's trace returned
                        if(P.tileCost) atomicAdd(P.tileCost + slot, sampleRays);
                        V3 accColor = hv, accNormal = pNormal, accRough = pRough;
                        float accRoughA = pRoughA, accContrib = pContrib, accDepth = depth;
                        bool last = true;
                        if(S > 1u) {   // park this sample; the lane finishing the pixel's last sample sums all of them in order
                            float4* rec = P.sampleScratch + 3 * ((size_t)pslot * S + (uint32_t)sample);
                            __stcg(rec, make_float4(hv.x, hv.y, hv.z, pContrib));
                            __stcg(rec + 1, make_float4(pNormal.x, pNormal.y, pNormal.z, depth));
                            __stcg(rec + 2, make_float4(pRough.x, pRough.y, pRough.z, pRoughA));
                            __threadfence();
                            last = atomicAdd(P.sampleDone + pslot, 1u) == S - 1u;
                            if(last) {
                                __threadfence();
                                P.sampleDone[pslot] = 0u;   // ready for the next frame
                                accColor = v3(0, 0, 0); accNormal = v3(0, 0, 0); accRough = v3(0, 0, 0); accRoughA = 0; accContrib = 0; accDepth = 0;
                                const float4* all = P.sampleScratch + 3 * (size_t)pslot * S;
                                for(uint32_t i = 0; i < S; ++i) {   // raygen.h:105-111 in the loop's order
                                    const float4 a = __ldcg(all + 3 * i), b = __ldcg(all + 3 * i + 1), c = __ldcg(all + 3 * i + 2);
                                    accColor = accColor + v3(a.x, a.y, a.z); accNormal = accNormal + v3(b.x, b.y, b.z);
                                    accRough = accRough + v3(c.x, c.y, c.z); accRoughA += c.w; accContrib += a.w; accDepth += b.w;
                                }
                            }
                        }
                        if(last) {      // raygen.h:114 + raygen.rgen:35-38
                            const float inv = (float)numSamples;
                            const uint2 ob = packHalf4(accColor.x / inv, accColor.y / inv, accColor.z / inv, accContrib / inv);
                            const uint2 on = packHalf4(accNormal.x / inv, accNormal.y / inv, accNormal.z / inv, (logf(accDepth) * 0.25f) / inv);
                            const uint2 orr = packHalf4(accRough.x / inv, accRough.y / inv, accRough.z / inv, accRoughA / inv);
                            // the pixel goes to every GPU whose post-chain rectangle contains it (own images or peer memory over NVLink)
        #pragma unroll
                            for(int q = 0; q < (MULTI ? kMaxPeers : 1); ++q) {   // static indices: the targets stay in the constant bank
                                if((uint32_t)q < P.nTargets) {
                                    const int qx = (int)lx - P.targets[q].x0, qy = (int)ly - P.targets[q].y0;
                                    if(qx >= 0 && qy >= 0 && qx < P.targets[q].w && qy < P.targets[q].h) {
                                        const size_t o = (size_t)qy * P.targets[q].w + qx;
                                        P.targets[q].base[o] = ob; P.targets[q].normal[o] = on; P.targets[q].rough[o] = orr;
                                    }
                                }
                            }
                        }
                        active = false;
                        break;
                    }
                    float* f = fr[sp - 1];
                    const int flags = __float_as_int(f[F_FLAGS]);
                    recDepth = __float_as_int(f[F_RECDEPTH]);   // undoes every recDepth++ / += rayConsumption below this frame
                    if((flags & 0xff) == FR_SHI) {   // closesthit.rchit:136-146, after T3 returned
                        V3 shadowCol = v3(f[F_ORG], f[F_ORG + 1], f[F_ORG + 2]);
                        if(depth < 1000.0f) {
                            hv = hv * shadowCol;
                        } else {
                            const V3 D = v3(f[F_DIR], f[F_DIR + 1], f[F_DIR + 2]), n = v3(f[F_N], f[F_N + 1], f[F_N + 2]);
                            const V3 dir = refract3(D, n, f[F_IOR]);
                            const float dp = dot(L, dir);
                            const float dp2 = dp * dp;
                            shadowCol = shadowCol * (dp2 * dp2 * dp + 0.75f);   // pow(x, 5)
                            cnt[CNT_SKY]++;   // T4: cull mask 0 -> always miss 0
                            const V3 sky = skyColor(-dir, L, strictIeee);
                            depth = 10000.0f;
                            hv = shadowCol + sky * 0.1f;
                        }
                        sp--;
                        continue;
                    }
                    int stage = (flags >> 8) & 0xff;
                    const bool frontFacing = ((flags >> 16) & 1) != 0;
                    const int rc = (flags >> 20) & 0xff;
                    if(stage == ST_SHADOW_RET) {   // :196-197 then :254-255
                        const V3 diffuse = v3(f[F_DIFF], f[F_DIFF + 1], f[F_DIFF + 2]);
                        V3 base = v3(f[F_BASE], f[F_BASE + 1], f[F_BASE + 2]) * hv + diffuse * f[F_EMIS];
                        f[F_BASE] = base.x; f[F_BASE + 1] = base.y; f[F_BASE + 2] = base.z;
                        stage = ST_TRY_REFLECT;
                    }
                    if(stage == ST_TRY_REFLECT) {  // :207-221
                        if(recDepth < maxRec && f[F_REFL] > 0.0f) {
                            const V3 D = v3(f[F_DIR], f[F_DIR + 1], f[F_DIR + 2]), n = v3(f[F_N], f[F_N + 1], f[F_N + 2]);
                            recDepth += rc; refDepth += f[F_T];
                            ro = v3(f[F_ORG], f[F_ORG + 1], f[F_ORG + 2]); rd = reflect3(D, n); rtmin = 0.01f; rtmax = 1000.0f;
                            rayType = RT_GENERIC; missIndex = 0; rayKind = CNT_REFLECT;
                            f[F_FLAGS] = __int_as_float((flags & ~0xff00) | (ST_REFLECT_RET << 8));
                            issue = true;
                            break;
                        }
                        f[F_RCOL] = 1.0f; f[F_RCOL + 1] = 1.0f; f[F_RCOL + 2] = 1.0f; f[F_RDEPTH] = 0.0f;
                        stage = ST_TRY_REFRACT;
                    }
                    if(stage == ST_REFLECT_RET) {
                        f[F_RCOL] = hv.x * f[F_SPEC]; f[F_RCOL + 1] = hv.y * f[F_SPEC + 1]; f[F_RCOL + 2] = hv.z * f[F_SPEC + 2];
                        f[F_RDEPTH] = depth;
                        stage = ST_TRY_REFRACT;
                    }
                    V3 refractColor = v3(1.0f, 1.0f, 1.0f);
                    if(stage == ST_TRY_REFRACT) {  // :224-251
                        if(recDepth < maxRec && f[F_TRANSP] > 0.0f) {
                            const V3 D = v3(f[F_DIR], f[F_DIR + 1], f[F_DIR + 2]), n = v3(f[F_N], f[F_N + 1], f[F_N + 2]);
                            const float ior = f[F_IOR];
                            const float eta = frontFacing ? curIOR / ior : ior / 1.0f;
                            recDepth++;
                            curIOR = frontFacing ? ior : 1.0f;
                            ro = v3(f[F_ORG], f[F_ORG + 1], f[F_ORG + 2]); rd = refract3(D, n, eta); rtmin = 0.01f; rtmax = 1000.0f;
                            rayType = RT_GENERIC; missIndex = 0; rayKind = CNT_REFRACT;
                            f[F_FLAGS] = __int_as_float((flags & ~0xff00) | ((frontFacing ? ST_REFRACT_RET_FRONT : ST_REFRACT_RET_BACK) << 8));
                            issue = true;
                            break;
                        }
                        stage = ST_COMBINE;
                    } else if(stage == ST_REFRACT_RET_FRONT) {
                        refractColor = hv;
                    } else if(stage == ST_REFRACT_RET_BACK) {
                        const V3 diffuse = v3(f[F_DIFF], f[F_DIFF + 1], f[F_DIFF + 2]);
                        refractColor = mix3(v3(1, 1, 1), diffuse, logf(1.0f + f[F_T])) * hv;
                    }
                    // combine, :254-267
                    {
                        const V3 base = v3(f[F_BASE], f[F_BASE + 1], f[F_BASE + 2]);
                        const V3 reflectColor = v3(f[F_RCOL], f[F_RCOL + 1], f[F_RCOL + 2]);
                        const float transparency = f[F_TRANSP], reflectivity = f[F_REFL], roughness = f[F_ROUGH];
                        const float totalContrib = glmax(transparency, reflectivity);
                        float weight = reflectivity / (transparency + reflectivity);
                        if(!strictIeee && (transparency + reflectivity) == 0.0f) weight = 0.0f;   // SURVEY hazard 8
                        const V3 roughCol = mix3(refractColor, reflectColor, weight);
                        hv = mix3(base, roughCol, totalContrib);
                        if(recDepth == 0) {
                            hv = base;
                            pNormal = v3(f[F_N], f[F_N + 1], f[F_N + 2]);
                            pRough = roughCol; pRoughA = glmin((f[F_RDEPTH] / 50.0f) * roughness, roughness / 2.1f);
                            pContrib = totalContrib;
                        }
                        depth = f[F_T];
                        sp--;
                    }
                }
            storeCtx(k);
            cStatus[k] = active ? CS_RAY : CS_IDLE;
        }
    }

    // ---- ray counters: warp reduce, one atomic per warp and counter
#pragma unroll
    for(int k = 0; k < CNT_N; ++k) {
        if(!COUNT && k >= CNT_NODES) break;
        unsigned long long v = cnt[k];
#pragma unroll
        for(int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if(lane == 0 && v) atomicAdd(P.counters + k, v);
    }
}

__global__ void k_trace_rays(const TraceParams P, const float* __restrict__ rays8, uint32_t n, float* __restrict__ tuv, uint32_t* __restrict__ instPrim) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const float* q = rays8 + 8 * (size_t)i;
    Hit hit;
    uint32_t cnt[CNT_N];
    traverse<false>(P, q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7], hit, cnt);
    tuv[3 * i] = hit.t; tuv[3 * i + 1] = hit.u; tuv[3 * i + 2] = hit.v;
    instPrim[2 * i] = hit.inst; instPrim[2 * i + 1] = hit.prim;
}

}  // namespace

uint32_t traceShareTiles(uint32_t dw, uint32_t dh, uint32_t rank, uint32_t world) {
    const uint32_t nTiles = ((dw + 7) / 8) * ((dh + 3) / 4);
    const uint32_t nChunks = (nTiles + kChunkTiles - 1) / kChunkTiles;
    const uint32_t myChunks = nChunks > rank ? (nChunks - rank + world - 1) / world : 0u;
    return myChunks * kChunkTiles;
}

namespace {
// Counting sort of the tile slots into 8 cost classes (relative to the mean), most expensive class first.
__global__ void __launch_bounds__(1024) k_order_tiles(uint32_t* cost, uint32_t n, uint32_t* order) {
    __shared__ unsigned long long sSum;
    __shared__ uint32_t sCount[8], sBase[8];
    if(threadIdx.x == 0) sSum = 0ull;
    if(threadIdx.x < 8) sCount[threadIdx.x] = 0u;
    __syncthreads();
    unsigned long long local = 0;
    for(uint32_t i = threadIdx.x; i < n; i += blockDim.x) local += cost[i];
    atomicAdd(&sSum, local);
    __syncthreads();
    const float mean = fmaxf((float)sSum / (float)(n ? n : 1u), 1.0f);
    auto classOf = [&](uint32_t c) {   // 7 = >= 8x mean ... 0 = < mean / 8
        const float r = (float)c / mean;
        int k = 3 + (int)floorf(log2f(fmaxf(r, 1e-6f)));
        return k < 0 ? 0 : (k > 7 ? 7 : k);
    };
    for(uint32_t i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&sCount[classOf(cost[i])], 1u);
    __syncthreads();
    if(threadIdx.x == 0) { uint32_t run = 0; for(int k = 7; k >= 0; --k) { sBase[k] = run; run += sCount[k]; } }
    __syncthreads();
    for(uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t c = cost[i];
        order[atomicAdd(&sBase[classOf(c)], 1u)] = i;
        cost[i] = 0u;
    }
}
}  // namespace

void launchOrderTiles(uint32_t* cost, uint32_t nSlots, uint32_t* order, cudaStream_t stream) {
    if(nSlots) k_order_tiles<<<1, 1024, 0, stream>>>(cost, nSlots, order);
}

template <bool COUNT, bool MULTI, int K>
static void launchTraceK(const TraceParams& p, int numSms, cudaStream_t stream) {
    // persistent grid: a multiple of the SM count; resident CTAs per SM limited by registers / local memory
    int perSm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_trace<COUNT, MULTI, K>, 128, 0);
    if(perSm < 1) perSm = 1;
    k_trace<COUNT, MULTI, K><<<numSms * perSm, 128, 0, stream>>>(p);
}

void launchTrace(const TraceParams& p, int numSms, cudaStream_t stream) {
    static int contexts = [] { const char* e = getenv("RGB200_LANE_CONTEXTS"); const int v = e ? atoi(e) : RG_LANE_CONTEXTS; return v == 2 || v == 4 ? v : 1; }();
    if(p.flags & RG_COUNT_TRAVERSAL) { launchTraceK<true, true, 1>(p, numSms, stream); return; }
    const bool multi = p.nTargets > 1;
    if(contexts == 4) { if(multi) launchTraceK<false, true, 4>(p, numSms, stream); else launchTraceK<false, false, 4>(p, numSms, stream); }
    else if(contexts == 2) { if(multi) launchTraceK<false, true, 2>(p, numSms, stream); else launchTraceK<false, false, 2>(p, numSms, stream); }
    else { if(multi) launchTraceK<false, true, 1>(p, numSms, stream); else launchTraceK<false, false, 1>(p, numSms, stream); }
}

void launchTraceRays(const TraceParams& p, const float* rays8, uint32_t n, float* tuv, uint32_t* instPrim, cudaStream_t stream) {
    if(n == 0) return;
    k_trace_rays<<<(n + 127) / 128, 128, 0, stream>>>(p, rays8, n, tuv, instPrim);
}

}  // namespace rg
