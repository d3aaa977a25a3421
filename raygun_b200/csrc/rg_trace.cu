// rg_trace.cu -- persistent-thread ray generation + two-level wide-BVH traversal + Whitted shading for sm_100a.
//
// Replaces the reference's ray-tracing pipeline (raygun/render/raytracer.cpp:99 traceRaysKHR) and its four shaders:
//   raygen      resources/shaders/raygen.rgen:30-39, raygen.h:37-115
//   closest hit resources/shaders/closesthit.rchit:74-268
//   miss 0 / 1  resources/shaders/miss.rmiss:38-83, shadowMiss.rmiss:30-34
// B200 has no RT cores, so traceRayEXT becomes a software traversal of compressed 8-wide nodes (rg_types.cuh) and
// the shader recursion becomes an explicit stack of frames per pixel sample ("context"): its ray tree is walked
// depth-first in the reference's order (shadow -> reflection -> refraction), with ONE mutable payload, so every
// stale-state effect of the GLSL (SURVEY.md 8a hazards 1-6) is reproduced.  Every warp of the persistent grid runs a pool
// of contexts with ray / hit queues in shared memory: lanes take the next ray the moment theirs is done (warp vote +
// prefix rank), and hits are shaded 32 at a time (k_trace_pool); coherent scenes use one context per lane (k_trace_lanes).
//
// Ray / triangle arithmetic is bit-identical to the oracle (explicitly rounded operations, never contracted):
// watertight Woop test, t preserved across the instance transform, ties resolved to the smallest (instance, primitive).
#include "rg_trace.cuh"

#include <cuda_fp16.h>

#include <cstdlib>

#include "../../include/rgb200.h"

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
#ifndef RG_INLINE_MATH
#define RG_MATH_FN __device__ __noinline__   // ONE copy of every library expansion: the kernels are sensitive to their instruction-cache
#else                                        // footprint (an all-inlined build, +224 instructions, is 4 % slower on C2); results are the same
#define RG_MATH_FN __device__ __forceinline__
#endif
RG_MATH_FN float rsqrtIeee(float x) { return 1.0f / sqrtf(x); }
RG_MATH_FN float powShared(float x, float y) { return powf(x, y); }
RG_MATH_FN float divShared(float x, float y) { return x / y; }
RG_MATH_FN float logShared(float x) { return logf(x); }
RG_MATH_FN float sqrtShared(float x) { return sqrtf(x); }
__device__ __forceinline__ V3 normalize(V3 a) { const float r = rsqrtIeee(dot(a, a)); return a * r; }
__device__ __forceinline__ V3 mix3(V3 a, V3 b, float t) { return a * (1.0f - t) + b * t; }
__device__ __forceinline__ float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
// GLSL min / max / clamp as GLM evaluates them ((y < x) ? y : x ...): NaN propagates exactly as in the oracle (RG_STRICT_IEEE)
__device__ __forceinline__ float glmin(float x, float y) { return (y < x) ? y : x; }
__device__ __forceinline__ float glmax(float x, float y) { return (x < y) ? y : x; }
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return glmin(glmax(x, lo), hi); }
__device__ __forceinline__ float glmod(float x, float y) { return x - y * floorf(divShared(x, y)); }
__device__ __forceinline__ V3 reflect3(V3 I, V3 N) { return I - N * (dot(N, I) * 2.0f); }
__device__ __forceinline__ V3 refract3(V3 I, V3 N, float eta) {
    const float d = dot(N, I);
    const float k = 1.0f - eta * eta * (1.0f - d * d);
    if(!(k >= 0.0f)) return v3(0.0f, 0.0f, 0.0f);
    return eta * I - (eta * d + sqrtShared(k)) * N;
}

// n / d and n % d for any n < 2^32 with m = floor(2^32 / d) (0xffffffff for d == 1): the high product is the quotient or one below it.
// The work-item decode runs for every tile sample; two runtime divisions there are ~50 instructions of the loop every warp keeps in
// the instruction caches, this is ~12.
__device__ __forceinline__ uint32_t divMagic(uint32_t n, uint32_t d, uint32_t m, uint32_t& rem) {
    uint32_t q = __umulhi(n, m), r = n - q * d;
    if(r >= d) { ++q; r -= d; }
    rem = r;
    return q;
}

// ------------------------------------------------------------------------------------------------ traversal
struct RayCtx {
    float ox, oy, oz;
    float ix, iy, iz;     // 1/d (zero components clamped to +-1e-30)
    float Sx, Sy, Sz;     // Woop shear
    uint32_t octinv;      // bit 2/1/0 set when 1/d.x, 1/d.y, 1/d.z > 0
    int kx, ky, kz;       // kz == kNoShear: shear constants not computed yet
};

// component k of (x, y, z) as two predicated selects: written in PTX because the compiler turns the ?: chain into divergent branches
// (8 reconvergence barriers per triangle test; the axes differ from lane to lane)
__device__ __forceinline__ float sel3(int k, float x, float y, float z) {
    float r;
    asm("{\n\t.reg .pred p0, p1;\n\tsetp.eq.s32 p0, %1, 0;\n\tsetp.eq.s32 p1, %1, 1;\n\tselp.f32 %0, %3, %4, p1;\n\tselp.f32 %0, %2, %0, p0;\n\t}"
        : "=f"(r) : "r"(k), "f"(x), "f"(y), "f"(z));
    return r;
}
__device__ __forceinline__ float __frcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// The ray constants of the slab test (every node) ...
__device__ __forceinline__ void setupSlab(RayCtx& r, float ox, float oy, float oz, float dx, float dy, float dz) {
    r.ox = ox; r.oy = oy; r.oz = oz;
    const float eps = 1e-30f;
    // approximate reciprocal (1 ulp): only the conservative slab test uses it; the slack below covers it
    r.ix = __frcp_approx(fabsf(dx) > eps ? dx : copysignf(eps, dx));
    r.iy = __frcp_approx(fabsf(dy) > eps ? dy : copysignf(eps, dy));
    r.iz = __frcp_approx(fabsf(dz) > eps ? dz : copysignf(eps, dz));
    // by the sign BIT of the (clamped) direction, i.e. of 1/d: the near / far plane choice of the slab test must agree with it (d = -0)
    r.octinv = (__float_as_int(r.ix) >= 0 ? 4u : 0u) | (__float_as_int(r.iy) >= 0 ? 2u : 0u) | (__float_as_int(r.iz) >= 0 ? 1u : 0u);
}
// ... and those of the triangle test (Woop shear; two IEEE divisions).  Only needed inside an instance, so they are computed when
// the first instance is entered (kz == kNoShear until then): rays that miss every instance box never pay for them, and a ray that
// enters a rotated / scaled instance pays once (for the object-space direction) instead of twice.
constexpr int kNoShear = 3;
// CALLS: the two IEEE divisions through the shared routine (pool kernel: C3 44.1 -> 43.1 ms) or expanded in place (lanes kernel: a
// call inside its traversal loop costs 2 %); measured per kernel, same results.
template <bool CALLS>
__device__ __forceinline__ void setupShear(RayCtx& r, float dx, float dy, float dz) {
    const float ax = fabsf(dx), ay = fabsf(dy), az = fabsf(dz);
    const int kz = (ax > ay) ? ((ax > az) ? 0 : 2) : ((ay > az) ? 1 : 2);
    int kx = kz == 2 ? 0 : kz + 1;
    int ky = kx == 2 ? 0 : kx + 1;
    const float dkz = sel3(kz, dx, dy, dz);
    if(dkz < 0.0f) { const int t = kx; kx = ky; ky = t; }
    r.kx = kx; r.ky = ky; r.kz = kz;
    r.Sx = CALLS ? divShared(sel3(kx, dx, dy, dz), dkz) : __fdiv_rn(sel3(kx, dx, dy, dz), dkz);
    r.Sy = CALLS ? divShared(sel3(ky, dx, dy, dz), dkz) : __fdiv_rn(sel3(ky, dx, dy, dz), dkz);
    r.Sz = __frcp_rn(dkz);
}

__device__ __forceinline__ uint32_t bitsel(uint32_t m, uint32_t a, uint32_t b) {   // (m & a) | (~m & b), opaque to the optimiser
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(d) : "r"(m), "r"(a), "r"(b));
    return d;
}

template <int J>
__device__ __forceinline__ float byteF(uint32_t w) {  // 1 + b / 32768 for byte J of w: ONE PRMT puts b into mantissa bits 15..8 of 1.0f;
    // (the constant is the FIRST operand so that the selector is the instruction's immediate: PRMT Rd, Rconst, 0x32x0, Rw)
    return __uint_as_float(__byte_perm(0x3F800000u, w, 0x3240u + (J << 4)));   // the affine map back to b is folded into the FFMA constants
}

#ifdef RG_FFMA2   // measured: no gain; moving the byte decode of one or two axes from PRMT (ALU pipe) to I2F.U8 (conversion pipe) gains nothing
// either (C2 4.34 / 4.33 / 4.29 ms): the node step is bound by the number of instructions issued, not by one pipe
// d = a * b + c on two binary32 values at once (sm_100a FFMA2; b is broadcast): near and far plane of one axis share the multiplier
__device__ __forceinline__ void fma2(float a0, float a1, float b, float c0, float c1, float& d0, float& d1) {
    asm("{\n\t.reg .b64 va, vb, vc, vd;\n\tmov.b64 va, {%2, %3};\n\tmov.b64 vb, {%4, %4};\n\tmov.b64 vc, {%5, %6};\n\t"
        "fma.rn.f32x2 vd, va, vb, vc;\n\tmov.b64 {%0, %1}, vd;\n\t}"
        : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b), "f"(c0), "f"(c1));
}
#endif

// One child of a node: slab test on the quantised planes; a hit sets nibble POS of the hit mask.  J: byte of the plane words.
template <int J, int POS>
__device__ __forceinline__ void childTest(uint32_t nx, uint32_t ny, uint32_t nz, uint32_t fx, uint32_t fy, uint32_t fz, float ax, float ay, float az, float onx,
                                          float ony, float onz, float ofx, float ofy, float ofz, float tmin, float tmax, uint32_t& hn) {
#ifdef RG_FFMA2
    float tnx, tny, tnz, tfx, tfy, tfz;
    fma2(byteF<J>(nx), byteF<J>(fx), ax, onx, ofx, tnx, tfx);
    fma2(byteF<J>(ny), byteF<J>(fy), ay, ony, ofy, tny, tfy);
    fma2(byteF<J>(nz), byteF<J>(fz), az, onz, ofz, tnz, tfz);
#else
    const float tnx = fmaf(byteF<J>(nx), ax, onx), tny = fmaf(byteF<J>(ny), ay, ony), tnz = fmaf(byteF<J>(nz), az, onz);
    const float tfx = fmaf(byteF<J>(fx), ax, ofx), tfy = fmaf(byteF<J>(fy), ay, ofy), tfz = fmaf(byteF<J>(fz), az, ofz);
#endif
    const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tmin));
    const float tf = fminf(fminf(tfx, tfy), fminf(tfz, tmax));
    if(tn <= tf) hn |= 0xFu << (4 * POS);
#ifdef RG_EXP_EXTRA_ALU   // experiment: sensitivity of the kernel to ALU-pipe instructions in the child test (3 more per child; results unchanged)
    if(__byte_perm(nx, fx, 0x5140 + J) == 0xdeadbeefu) hn ^= 1u;
#endif
}

#if RG_HALF_SLAB
__device__ __forceinline__ __half2 asHalf2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ uint32_t madlo(uint32_t a, uint32_t b, uint32_t c) {   // a * b + c as ONE IMAD (FMA pipe), opaque to the optimiser
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// TWO children of a node (positions 2 PAIR and 2 PAIR + 1: bytes SEL of the plane words) in packed binary16 arithmetic: every
// instruction serves both.  One PRMT puts a plane byte q into the low byte of each half: the SUBNORMAL 2^-24 q, exact, and HFMA2 takes
// subnormal operands at full rate; the node-local scaling of travNode keeps every plane distance inside binary16's range.  The near
// planes use the low halves of cx / cy / cz (rounding errors given away downwards), the far planes the high halves (upwards).  The
// results land in the low / high half of one mask = nibbles PAIR and PAIR + 4 of the hit mask (rg_types.cuh nibbleOfPos).
// SASS per pair: 6 PRMT + 6 HFMA2 + 2 HMNMX2 + 2 VHMNMX (three inputs) + HSET2 + LOP3 = 18, against 2 x 20 in binary32.
template <int SEL, int PAIR>
__device__ __forceinline__ void pairTest(uint32_t nx, uint32_t ny, uint32_t nz, uint32_t fx, uint32_t fy, uint32_t fz, __half2 axy, __half2 azz, __half2 cx,
                                         __half2 cy, __half2 cz, __half2 tmn, __half2 tmx, uint32_t& hn) {
    const __half2 ax = __low2half2(axy), ay = __high2half2(axy);   // operand swizzles of HFMA2, no instructions
    const __half2 tnx = __hfma2(asHalf2(__byte_perm(nx, 0u, SEL)), ax, __low2half2(cx)), tny = __hfma2(asHalf2(__byte_perm(ny, 0u, SEL)), ay, __low2half2(cy)),
                  tnz = __hfma2(asHalf2(__byte_perm(nz, 0u, SEL)), azz, __low2half2(cz));
    const __half2 tfx = __hfma2(asHalf2(__byte_perm(fx, 0u, SEL)), ax, __high2half2(cx)), tfy = __hfma2(asHalf2(__byte_perm(fy, 0u, SEL)), ay, __high2half2(cy)),
                  tfz = __hfma2(asHalf2(__byte_perm(fz, 0u, SEL)), azz, __high2half2(cz));
    const __half2 tn = __hmax2(__hmax2(tnx, tny), __hmax2(tnz, tmn));   // a NaN operand (never with finite scenes) is dropped: more hits, never fewer
    const __half2 tf = __hmin2(__hmin2(tfx, tfy), __hmin2(tfz, tmx));
    hn |= __hle2_mask(tn, tf) & (0x000F000Fu << (4 * PAIR));
}
#endif

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {   // default mode: nibble bit 3 = replicate the byte's sign
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

__device__ __noinline__ float edge64(float a, float b, float c, float d) {   // a * b - c * d, one rounding at the end of the binary64 evaluation
    return (float)__dsub_rn(__dmul_rn((double)a, (double)b), __dmul_rn((double)c, (double)d));
}
// Watertight ray / triangle test (Woop, Benthin, Wald 2013), no culling; operation order == oracle/orc_scene.cpp intersectTri.
__device__ __forceinline__ bool triTest(const RayCtx& r, const float4 p0, const float4 p1, const float4 p2, float tmin, float& tOut, float& uOut,
                                        float& vOut) {
    const float A0 = __fsub_rn(p0.x, r.ox), A1 = __fsub_rn(p0.y, r.oy), A2 = __fsub_rn(p0.z, r.oz);
    const float B0 = __fsub_rn(p1.x, r.ox), B1 = __fsub_rn(p1.y, r.oy), B2 = __fsub_rn(p1.z, r.oz);
    const float C0 = __fsub_rn(p2.x, r.ox), C1 = __fsub_rn(p2.y, r.oy), C2 = __fsub_rn(p2.z, r.oz);
    // the nine axis selects of the three vertices share six predicates (one asm block: 6 ISETP + 18 SEL instead of 18 + 18)
    float Akx, Aky, Akz, Bkx, Bky, Bkz, Ckx, Cky, Ckz;
    asm("{\n\t.reg .pred x0, x1, y0, y1, z0, z1;\n\t"
        "setp.eq.s32 x0, %9, 0;\n\tsetp.eq.s32 x1, %9, 1;\n\tsetp.eq.s32 y0, %10, 0;\n\tsetp.eq.s32 y1, %10, 1;\n\t"
        "setp.eq.s32 z0, %11, 0;\n\tsetp.eq.s32 z1, %11, 1;\n\t"
        "selp.f32 %0, %13, %14, x1;\n\tselp.f32 %0, %12, %0, x0;\n\tselp.f32 %1, %13, %14, y1;\n\tselp.f32 %1, %12, %1, y0;\n\t"
        "selp.f32 %2, %13, %14, z1;\n\tselp.f32 %2, %12, %2, z0;\n\t"
        "selp.f32 %3, %16, %17, x1;\n\tselp.f32 %3, %15, %3, x0;\n\tselp.f32 %4, %16, %17, y1;\n\tselp.f32 %4, %15, %4, y0;\n\t"
        "selp.f32 %5, %16, %17, z1;\n\tselp.f32 %5, %15, %5, z0;\n\t"
        "selp.f32 %6, %19, %20, x1;\n\tselp.f32 %6, %18, %6, x0;\n\tselp.f32 %7, %19, %20, y1;\n\tselp.f32 %7, %18, %7, y0;\n\t"
        "selp.f32 %8, %19, %20, z1;\n\tselp.f32 %8, %18, %8, z0;\n\t}"
        : "=&f"(Akx), "=&f"(Aky), "=&f"(Akz), "=&f"(Bkx), "=&f"(Bky), "=&f"(Bkz), "=&f"(Ckx), "=&f"(Cky), "=&f"(Ckz)
        : "r"(r.kx), "r"(r.ky), "r"(r.kz), "f"(A0), "f"(A1), "f"(A2), "f"(B0), "f"(B1), "f"(B2), "f"(C0), "f"(C1), "f"(C2));
    const float Ax = __fsub_rn(Akx, __fmul_rn(r.Sx, Akz)), Ay = __fsub_rn(Aky, __fmul_rn(r.Sy, Akz));
    const float Bx = __fsub_rn(Bkx, __fmul_rn(r.Sx, Bkz)), By = __fsub_rn(Bky, __fmul_rn(r.Sy, Bkz));
    const float Cx = __fsub_rn(Ckx, __fmul_rn(r.Sx, Ckz)), Cy = __fsub_rn(Cky, __fmul_rn(r.Sy, Ckz));
    float U = __fsub_rn(__fmul_rn(Cx, By), __fmul_rn(Cy, Bx));
    float V = __fsub_rn(__fmul_rn(Ax, Cy), __fmul_rn(Ay, Cx));
    float W = __fsub_rn(__fmul_rn(Bx, Ay), __fmul_rn(By, Ax));
    if(U == 0.0f || V == 0.0f || W == 0.0f) {  // rare: exact edge hit, redo in binary64 (products of floats are exact there); out of line
        U = edge64(Cx, By, Cy, Bx); V = edge64(Ax, Cy, Ay, Cx); W = edge64(Bx, Ay, By, Ax);
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

// The world-space ray of the traversal in flight: in registers (one ray per lane), or in the warp pool in shared memory.
struct WorldRayRegs {
    float o[3], d[3], tm;
    __device__ __forceinline__ float ox() const { return o[0]; }
    __device__ __forceinline__ float oy() const { return o[1]; }
    __device__ __forceinline__ float oz() const { return o[2]; }
    __device__ __forceinline__ float dx() const { return d[0]; }
    __device__ __forceinline__ float dy() const { return d[1]; }
    __device__ __forceinline__ float dz() const { return d[2]; }
    __device__ __forceinline__ float tmax() const { return tm; }
};
struct WorldRayPool {   // structure-of-arrays: component k of context c at p[k * kPoolCtx]
    const float* p;
    __device__ __forceinline__ float ox() const { return p[0]; }
    __device__ __forceinline__ float oy() const { return p[kPoolCtx]; }
    __device__ __forceinline__ float oz() const { return p[2 * kPoolCtx]; }
    __device__ __forceinline__ float dx() const { return p[3 * kPoolCtx]; }
    __device__ __forceinline__ float dy() const { return p[4 * kPoolCtx]; }
    __device__ __forceinline__ float dz() const { return p[5 * kPoolCtx]; }
    __device__ __forceinline__ float tmax() const { return p[7 * kPoolCtx]; }
};

// Where the traversal stack of a ray lives.  StackLocal: the thread's local memory (through L1).  StackHybrid<D>: the first D entries in
// shared memory (entry i of thread t at sh[i * 128 + t]: conflict-free 64-bit accesses, no L1 lines, no write-backs of dead entries), the rest
// -- deep trees only -- in local memory.
struct StackLocal {
    uint2* p;
    __device__ __forceinline__ void put(int i, uint2 v) const { p[i] = v; }
    __device__ __forceinline__ uint2 get(int i) const { return p[i]; }
};
template <int D>
struct StackHybrid {
    uint2* sh;   // + threadIdx.x already
    uint2* p;    // local part: entries D..kStackSize-1
    __device__ __forceinline__ void put(int i, uint2 v) const { if(i < D) sh[i * 128] = v; else p[i - D] = v; }
    __device__ __forceinline__ uint2 get(int i) const { return i < D ? sh[i * 128] : p[i - D]; }
};

// Resumable closest-hit traversal.  The state of one ray lives in registers (+ a stack in local memory) and travStep advances
// it by ONE node test and / or ONE primitive test, so a warp can hand a finished lane its next ray at any step, and can leave
// the traversal loop to shade while some lanes are still on their way.
struct Trav {
    RayCtx r;            // the ray in the space being traversed (world in the TLAS, object inside an instance)
    uint2 ng, tg;        // current node group (child base, hit bits << 24 | imask) and primitive group (base, hit bits)
    int sp;
    uint32_t curInst;    // instance being traversed; kInvalid while in the TLAS
};

// Closest hit in (tmin, tmax).  hit.inst == kInvalid on miss.
// LAZY: leave the triangle-test constants to the first instance entry (setupShear); otherwise compute them for the world-space direction now
// (the pool scheduler starts rays with more lanes than it has at an instance entry)
template <bool LAZY>
__device__ __forceinline__ void travInit(const TraceParams& P, Trav& T, Hit& hit, float ox, float oy, float oz, float dx, float dy, float dz, float tmax) {
    hit.t = tmax; hit.u = 0.0f; hit.v = 0.0f; hit.inst = kInvalid; hit.prim = kInvalid;
    T.sp = 0;
    T.tg = make_uint2(0u, 0u);
    T.curInst = kInvalid;
    // nothing to traverse (the first travStep reports a miss): empty scene, zero direction (refract on total internal reflection), NaN ray
    // (|dx| + |dy| + |dz| > 0 is false for the zero direction and for a NaN component alike; no product, so nothing underflows)
    const bool none = P.nInst == 0 || !(fabsf(dx) + fabsf(dy) + fabsf(dz) > 0.0f) || !(ox == ox && oy == oy && oz == oz);
    T.ng = make_uint2(0u, none ? 0u : 0x80000000u);
    T.r.kx = 0; T.r.ky = 0; T.r.kz = kNoShear; T.r.Sx = 0.0f; T.r.Sy = 0.0f; T.r.Sz = 0.0f;
    if(!none) { setupSlab(T.r, ox, oy, oz, dx, dy, dz); if(!LAZY) setupShear<true>(T.r, dx, dy, dz); }
}

// The three kinds of traversal work.  travNode: the 8 children of the next node of the lane's node group; leaves T.ng / T.tg = the hit
// children / primitives of that node (primitives the lane still held are parked on its stack: postponed tests).
template <bool COUNT, class ST>
__device__ __forceinline__ void travNode(const TraceParams& P, Trav& T, const ST& stack, const Hit& hit, float tmin, uint32_t* cnt) {
    RayCtx& r = T.r;
    if(T.tg.y && T.sp < kStackSize) stack.put(T.sp++, T.tg);
    uint2 ng = T.ng;
    const int bit = 31 - __clz(ng.y);
    ng.y &= ~(1u << bit);
    const uint32_t slot = (uint32_t)(bit - 24) ^ r.octinv;
    const uint32_t rel = __popc(ng.y & 0xffu & ((1u << slot) - 1u));
    if((ng.y & 0xff000000u) && T.sp < kStackSize) stack.put(T.sp++, ng);
    const Node8* nodes = T.curInst != kInvalid ? P.blasNodes : P.tlasNodes;
    const uint4* np = reinterpret_cast<const uint4*>(nodes + (ng.x + rel));
    const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
    if(COUNT) cnt[CNT_NODES]++;
    // 2^e per axis from the exponent bytes (x: one shift; the bit of ey that lands in the sign is dropped by |.|, an operand modifier)
    const float sx = fabsf(__uint_as_float(n0.w << 23)), sy = __uint_as_float(((n0.w >> 8) & 0xffu) << 23),
                sz = __uint_as_float(((n0.w >> 16) & 0xffu) << 23);
    const float px = __uint_as_float(n0.x) - r.ox, py = __uint_as_float(n0.y) - r.oy, pz = __uint_as_float(n0.z) - r.oz;
#if RG_HALF_SLAB
    // Plane distance of byte q on one axis: t(q) = q A + B with A = 2^e / d (per quantisation step) and B = (p - o) / d.  The children are
    // tested in binary16, two per instruction (pairTest), on t'(q) = (t(q) - t0) s -- a monotone map, so the slab test is unchanged:
    //   s   power of two that brings the STEEPEST axis to |A| s in [2^-10, 2^-9): a node spans < 1/2 in t' on every axis, and the half
    //       operand 2^24 A s stays below 2^15;
    //   t0  = B of the axis with the smallest |A| (the axis along which the ray crosses the node fastest): for a ray that hits the node
    //       every |B - t0| is then at most the sum of two axis spans, so the roundings of binary16 (relative 2^-11: the multiplier,
    //       the addend, the result) move a plane by < 1/2 quantisation step of ITS OWN axis -- the price is a child box that looks
    //       ~1 % larger; a ray that misses the node by far has large |B - t0| and errors to match, but a gap as large.
    // Everything the roundings can do is bounded by e (per axis: relative to |c| and |A|, plus the binary32 error of B, which is relative
    // to |B| and cancels -- the slack of the binary32 formulation -- plus the subnormal spacing) and given away: near planes use c - e,
    // far planes c + e, so near <= far holds for a flat child (qlo == qhi) too.  tests/test_slab_half_model.py restates this arithmetic
    // in numpy and checks it against binary64 over millions of random nodes and rays (incl. axis-parallel and far-away ones).
    // Overflow to +-inf only happens in the addends (planes > 65 504 node spans away: a true miss); q * finite never makes a NaN.
    const float ax = sx * r.ix, ay = sy * r.iy, az = sz * r.iz;
    const float bx = px * r.ix, by = py * r.iy, bz = pz * r.iz;
    const float aax = fabsf(ax), aay = fabsf(ay), aaz = fabsf(az);
    const uint32_t me = min(max(__float_as_uint(fmaxf(aax, fmaxf(aay, aaz))) & 0x7f800000u, 0x0A000000u), 0x79000000u);   // clamps: degenerate nodes / rays only
    const float S = __uint_as_float(0x86000000u - me), s = __uint_as_float(0x7A000000u - me);   // 2^(14 - E), 2^(14 - E - 24); E = exponent of max |A|
    float t0 = aax <= aay ? bx : by;
    t0 = aaz < fminf(aax, aay) ? bz : t0;
    const float t0s = t0 * s;
    const float asx = ax * S, asy = ay * S, asz = az * S;
    const float csx = fmaf(bx, s, -t0s), csy = fmaf(by, s, -t0s), csz = fmaf(bz, s, -t0s);   // (B - t0) s, exactly 0 on the axis of t0
    // 1.02 x 2^-11 (rounding of the addend), 1.03 x 2^-27 (of the multiplier, times q <= 255), 2 x 2^-25 (subnormal spacing), error of B.
    // The rounding of the RESULT needs no allowance: it is monotone and applied to both sides of near <= far alike.
    const float kRelC = 5.0e-4f, kRelA = 7.7e-9f, kAbs = 6.0e-8f, kSlack = 7.3e-7f;
    const float e0 = fmaf(fabsf(t0s), kSlack, kAbs);
    const float ex = fmaf(fabsf(csx), kRelC, fmaf(fabsf(asx), kRelA, e0)), ey = fmaf(fabsf(csy), kRelC, fmaf(fabsf(asy), kRelA, e0)),
                ez = fmaf(fabsf(csz), kRelC, fmaf(fabsf(asz), kRelA, e0));
    const __half2 haxy = __floats2half2_rn(asx, asy), hazz = __floats2half2_rn(asz, asz);
    const __half2 hcx = __floats2half2_rn(csx - ex, csx + ex), hcy = __floats2half2_rn(csy - ey, csy + ey), hcz = __floats2half2_rn(csz - ez, csz + ez);
    // (tmin, hit.t) on the same scale, rounded outwards
    const float tn0 = fmaf(tmin, s, -t0s), tf0 = fmaf(hit.t, s, -t0s);
    const float tn1 = fmaf(fabsf(tn0), -kRelC, tn0) - kAbs, tf1 = fmaf(fabsf(tf0), kRelC, tf0) + kAbs;
    const __half2 htmn = __floats2half2_rn(tn1, tn1), htmx = __floats2half2_rn(tf1, tf1);
    const uint32_t vm = n1.w;
    // nibbles 0..n-1 of vm are the occupied ones and pair i holds nibbles i and i + 4: all four pairs are in use from n = 4 on, and
    // nodes nearly always have 8 children, so no pair is skipped (empty positions hold an inverted box and never hit)
    uint32_t hn = 0;
#if RG_PLANE_DIFF
    // near / far plane words by the sign of the direction: lo + s * (hi - lo) with s in {0, 1}, the difference words come from the builder
    const uint32_t sx1 = __float_as_uint(r.ix) >> 31, sy1 = __float_as_uint(r.iy) >> 31, sz1 = __float_as_uint(r.iz) >> 31;
    const uint32_t sx0 = madlo(sx1, 0xffffffffu, 1u), sy0 = madlo(sy1, 0xffffffffu, 1u), sz0 = madlo(sz1, 0xffffffffu, 1u);   // 1 - s
    {
        const uint32_t nx = madlo(sx1, n3.z, n2.x), fx = madlo(sx0, n3.z, n2.x), ny = madlo(sy1, n4.x, n2.z), fy = madlo(sy0, n4.x, n2.z),
                       nz = madlo(sz1, n4.z, n3.x), fz = madlo(sz0, n4.z, n3.x);
        pairTest<0x4140, 0>(nx, ny, nz, fx, fy, fz, haxy, hazz, hcx, hcy, hcz, htmn, htmx, hn);
        pairTest<0x4342, 1>(nx, ny, nz, fx, fy, fz, haxy, hazz, hcx, hcy, hcz, htmn, htmx, hn);
    }
    {
        const uint32_t nx = madlo(sx1, n3.w, n2.y), fx = madlo(sx0, n3.w, n2.y), ny = madlo(sy1, n4.y, n2.w), fy = madlo(sy0, n4.y, n2.w),
                       nz = madlo(sz1, n4.w, n3.y), fz = madlo(sz0, n4.w, n3.y);
        pairTest<0x4140, 2>(nx, ny, nz, fx, fy, fz, haxy, hazz, hcx, hcy, hcz, htmn, htmx, hn);
        pairTest<0x4342, 3>(nx, ny, nz, fx, fy, fz, haxy, hazz, hcx, hcy, hcz, htmn, htmx, hn);
    }
#else
    const uint32_t mx = (uint32_t)(__float_as_int(r.ix) >> 31), my = (uint32_t)(__float_as_int(r.iy) >> 31), mz = (uint32_t)(__float_as_int(r.iz) >> 31);
    {
        const uint32_t nx = bitsel(mx, n3.z, n2.x), fx = bitsel(mx, n2.x, n3.z), ny = bitsel(my, n4.x, n2.z), fy = bitsel(my, n2.z, n4.x),
                       nz = bitsel(mz, n4.z, n3.x), fz = bitsel(mz, n3.x, n4.z);
        pairTest<0x4140, 0>(nx, ny, nz, fx, fy, fz, haxy, hazz, hcx, hcy, hcz, htmn, htmx, hn);
        pairTest<0x4342, 1>(nx, ny, nz, fx, fy, fz, haxy, hazz, hcx, hcy, hcz, htmn, htmx, hn);
    }
    {
        const uint32_t nx = bitsel(mx, n3.w, n2.y), fx = bitsel(mx, n2.y, n3.w), ny = bitsel(my, n4.y, n2.w), fy = bitsel(my, n2.w, n4.y),
                       nz = bitsel(mz, n4.w, n3.y), fz = bitsel(mz, n3.y, n4.w);
        pairTest<0x4140, 2>(nx, ny, nz, fx, fy, fz, haxy, hazz, hcx, hcy, hcz, htmn, htmx, hn);
        pairTest<0x4342, 3>(nx, ny, nz, fx, fy, fz, haxy, hazz, hcx, hcy, hcz, htmn, htmx, hn);
    }
#endif
#else
    // plane distance t = q * (2^e / d) + (p - o) / d.  The second term cancels against the first, so its rounding
    // error (relative to |(p - o) / d|, NOT to t) is what can make the slab test miss: widen near / far by that much.
    // Near and far use the same q * adj, so a flat child box (qlo == qhi) always keeps near <= far.
    // byteF gives v = 1 + q / 32768, so t = v * A + (b - A) with A = 32768 * 2^e / d.  Rounding: ulp(A) = 1/256 of one
    // quantisation step in t, plus the error of b = (p - o) / d; both are covered by the slack (relative to |b| and to a step).
    const float ax = sx * r.ix, ay = sy * r.iy, az = sz * r.iz;   // the node's exponents carry the factor 32768 (kExpBias)
    const float bx = px * r.ix, by = py * r.iy, bz = pz * r.iz;
    const float kSlack = 7.2e-7f, kStep = 1.0f / (32768.0f * 64.0f);   // 1/64 of a quantisation step
    const float wx = fmaf(fabsf(bx), kSlack, fabsf(ax) * kStep), wy = fmaf(fabsf(by), kSlack, fabsf(ay) * kStep), wz = fmaf(fabsf(bz), kSlack, fabsf(az) * kStep);
    const float onx = (bx - ax) - wx, ony = (by - ay) - wy, onz = (bz - az) - wz;
    const float ofx = (bx - ax) + wx, ofy = (by - ay) + wy, ofz = (bz - az) + wz;
    // near / far plane words by the sign of the direction: one LOP3 each on the packed words BEFORE the byte decode
    // (a plain ?: lets the compiler select after decoding both, which doubles the PRMTs)
    const uint32_t mx = (uint32_t)(__float_as_int(r.ix) >> 31), my = (uint32_t)(__float_as_int(r.iy) >> 31), mz = (uint32_t)(__float_as_int(r.iz) >> 31);
    const uint32_t vm = n1.w;
    // The children sit in positions 0..n-1 (rg_types.cuh): two at a time, stop at the first empty pair (vm nibbles of
    // occupied positions are non-zero).  A hit child sets its nibble of hn.
    uint32_t hn = 0;
    {
        const uint32_t nx = bitsel(mx, n3.z, n2.x), fx = bitsel(mx, n2.x, n3.z), ny = bitsel(my, n4.x, n2.z), fy = bitsel(my, n2.z, n4.x),
                       nz = bitsel(mz, n4.z, n3.x), fz = bitsel(mz, n3.x, n4.z);
        childTest<0, 0>(nx, ny, nz, fx, fy, fz, ax, ay, az, onx, ony, onz, ofx, ofy, ofz, tmin, hit.t, hn);
        childTest<1, 1>(nx, ny, nz, fx, fy, fz, ax, ay, az, onx, ony, onz, ofx, ofy, ofz, tmin, hit.t, hn);
        if(vm >= 0x100u) {
            childTest<2, 2>(nx, ny, nz, fx, fy, fz, ax, ay, az, onx, ony, onz, ofx, ofy, ofz, tmin, hit.t, hn);
            childTest<3, 3>(nx, ny, nz, fx, fy, fz, ax, ay, az, onx, ony, onz, ofx, ofy, ofz, tmin, hit.t, hn);
        }
    }
    if(vm >= 0x10000u) {
        const uint32_t nx = bitsel(mx, n3.w, n2.y), fx = bitsel(mx, n2.y, n3.w), ny = bitsel(my, n4.y, n2.w), fy = bitsel(my, n2.w, n4.y),
                       nz = bitsel(mz, n4.w, n3.y), fz = bitsel(mz, n3.y, n4.w);
        childTest<0, 4>(nx, ny, nz, fx, fy, fz, ax, ay, az, onx, ony, onz, ofx, ofy, ofz, tmin, hit.t, hn);
        childTest<1, 5>(nx, ny, nz, fx, fy, fz, ax, ay, az, onx, ony, onz, ofx, ofy, ofz, tmin, hit.t, hn);
        if(vm >= 0x1000000u) {
            childTest<2, 6>(nx, ny, nz, fx, fy, fz, ax, ay, az, onx, ony, onz, ofx, ofy, ofz, tmin, hit.t, hn);
            childTest<3, 7>(nx, ny, nz, fx, fy, fz, ax, ay, az, onx, ony, onz, ofx, ofy, ofz, tmin, hit.t, hn);
        }
    }
#endif
    // Internal children that were hit -> bits 24 + (code ^ octant) of the node group, so that the highest bit is the nearest child:
    // ONE table look-up per four children (PRMT: selector nibble = code ^ octant picks the byte 1 << nibble; nibble 8 yields 0),
    // then the eight distinct one-hot bytes are summed into the top byte by a multiplication.
    const uint32_t innerNib = ((vm >> 3) & 0x11111111u) * 15u;
    const uint32_t sel = bitsel(hn & innerNib, n1.z ^ (r.octinv * 0x11111111u), 0x88888888u);
    const uint32_t onehot = prmt(0x08040201u, 0x80402010u, sel) | prmt(0x08040201u, 0x80402010u, sel >> 16);
    T.ng = make_uint2(n1.x, ((onehot * 0x01010101u) & 0xff000000u) | (n0.w >> 24));
    T.tg = make_uint2(n1.y, hn & vm & 0x77777777u);   // primitive k of position j: bit 4 j + k
}

// True when the ray o + t d (t >= 0) certainly misses the sphere (centre sph.xyz, squared radius sph.w) that contains every triangle of
// a mesh: the instance need not be entered.  Conservative: the discriminant B^2 - A C of |o + t d - c|^2 = r^2 carries a rounding
// error of a few ulp of A |c - o|^2, which is given away (4e-6 A |c - o|^2), so a grazing ray always enters; closest hits are unchanged.
__device__ __forceinline__ bool missesSphere(const float4 sph, float ox, float oy, float oz, float dx, float dy, float dz) {
    if(!(sph.w < 3.0e38f)) return false;   // no useful sphere for this mesh (flat or boxy)
    const float cx = sph.x - ox, cy = sph.y - oy, cz = sph.z - oz;
    const float A = fmaf(dx, dx, fmaf(dy, dy, dz * dz)), B = fmaf(cx, dx, fmaf(cy, dy, cz * dz)), L = fmaf(cx, cx, fmaf(cy, cy, cz * cz));
    const float C = L - sph.w;
    if(C > 0.0f && B < 0.0f) return true;               // origin outside, centre behind the ray
    return fmaf(B, B, -A * C) + 4.0e-6f * A * L < 0.0f;   // no real root
}

// travPrim: ONE primitive of the lane's primitive group -- a triangle test inside an instance, entering an instance in the TLAS.
// SPH: test the mesh's bounding sphere before entering an instance (pool scheduler: incoherent rays over many instances; the lanes
// kernel leaves it out -- on the example scene the extra load and test in its hot loop cost 5 % and reject nothing)
template <bool COUNT, bool SPH, class WR, class ST>
__device__ __forceinline__ void travPrim(const TraceParams& P, Trav& T, const ST& stack, Hit& hit, const WR& wray, float tmin, uint32_t* cnt) {
    RayCtx& r = T.r;
    const uint32_t bit = (uint32_t)(__ffs(T.tg.y) - 1);
    T.tg.y &= T.tg.y - 1u;
    const uint32_t primIdx = (T.tg.x & ~kPrimGroupBit) + bit - (bit >> 2);   // bit 4 j + k -> element 3 j + k of the group (rg_types.cuh)
    if(T.curInst == kInvalid) {
        const uint4* lp = reinterpret_cast<const uint4*>(P.tlasLeaves + primIdx);
        const uint4 l3 = __ldg(lp + 3), l0 = __ldg(lp), l1 = __ldg(lp + 1), l2 = __ldg(lp + 2);   // all in flight at once
        const uint4 l4 = SPH ? __ldg(lp + 4) : make_uint4(0u, 0u, 0u, 0x7f7fffffu);               // the mesh's bounding sphere in object space (w = r^2)
        // instance of an empty mesh, or stack exhausted (never with sane scenes): skip
        if(l3.x != kInvalid && T.sp + 6 <= kStackSize) {
            if(COUNT) cnt[CNT_INST]++;
            const float ox = wray.ox(), oy = wray.oy(), oz = wray.oz();
            bool enter = true, shear;
            float sdx = wray.dx(), sdy = wray.dy(), sdz = wray.dz();   // direction inside the instance
            const float4 sph = make_float4(__uint_as_float(l4.x), __uint_as_float(l4.y), __uint_as_float(l4.z), __uint_as_float(l4.w));
            if(l3.z) {
                // pure translation (flagged by the instance preparation): the direction and everything derived from it stay;
                // the oracle's ((1*ox + 0*oy) + 0*oz) + t is exactly ox + t
                const float tox = __fadd_rn(ox, __uint_as_float(l0.w)), toy = __fadd_rn(oy, __uint_as_float(l1.w)), toz = __fadd_rn(oz, __uint_as_float(l2.w));
                if(SPH && missesSphere(sph, tox, toy, toz, sdx, sdy, sdz)) return;
                if(T.tg.y) stack.put(T.sp++, T.tg);
                if(T.ng.y & 0xff000000u) stack.put(T.sp++, T.ng);
                stack.put(T.sp++, make_uint2(kInvalid, 0x1000u));
                r.ox = tox; r.oy = toy; r.oz = toz;
                shear = r.kz == kNoShear;   // first instance of this ray
            } else {
                // object-space ray: same operation order as the oracle (t is preserved, direction not normalised)
                const float* w0 = reinterpret_cast<const float*>(&l0); const float* w1 = reinterpret_cast<const float*>(&l1);
                const float* w2 = reinterpret_cast<const float*>(&l2);
                const float dx = sdx, dy = sdy, dz = sdz;
                const float oox = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w0[0], ox), __fmul_rn(w0[1], oy)), __fmul_rn(w0[2], oz)), w0[3]);
                const float ooy = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w1[0], ox), __fmul_rn(w1[1], oy)), __fmul_rn(w1[2], oz)), w1[3]);
                const float ooz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w2[0], ox), __fmul_rn(w2[1], oy)), __fmul_rn(w2[2], oz)), w2[3]);
                sdx = __fadd_rn(__fadd_rn(__fmul_rn(w0[0], dx), __fmul_rn(w0[1], dy)), __fmul_rn(w0[2], dz));
                sdy = __fadd_rn(__fadd_rn(__fmul_rn(w1[0], dx), __fmul_rn(w1[1], dy)), __fmul_rn(w1[2], dz));
                sdz = __fadd_rn(__fadd_rn(__fmul_rn(w2[0], dx), __fmul_rn(w2[1], dy)), __fmul_rn(w2[2], dz));
                shear = true;
                if((sdx == 0.0f && sdy == 0.0f && sdz == 0.0f) || (SPH && missesSphere(sph, oox, ooy, ooz, sdx, sdy, sdz))) {
                    enter = false;
                } else {
                    if(T.tg.y) stack.put(T.sp++, T.tg);
                    if(T.ng.y & 0xff000000u) stack.put(T.sp++, T.ng);
                    // the world-space slab / shear constants ride on the stack while the instance is traversed
                    stack.put(T.sp++, make_uint2(__float_as_uint(r.ix), __float_as_uint(r.iy)));
                    stack.put(T.sp++, make_uint2(__float_as_uint(r.iz), __float_as_uint(r.Sx)));
                    stack.put(T.sp++, make_uint2(__float_as_uint(r.Sy), __float_as_uint(r.Sz)));
                    stack.put(T.sp++, make_uint2(kInvalid, r.octinv | ((uint32_t)r.kx << 4) | ((uint32_t)r.ky << 6) | ((uint32_t)r.kz << 8)));
                    setupSlab(r, oox, ooy, ooz, sdx, sdy, sdz);
                }
            }
            if(enter && shear) setupShear<SPH>(r, sdx, sdy, sdz);   // SPH: the pool kernel
            if(enter) {
                T.curInst = l3.y;
                T.ng = make_uint2(l3.x, 0x80000000u);
                T.tg = make_uint2(0u, 0u);
            }
        }
        return;
    } else {
        const float4* tp = reinterpret_cast<const float4*>(P.tris + primIdx);
        const float4 p0 = __ldg(tp), p1 = __ldg(tp + 1), p2 = __ldg(tp + 2);
        if(COUNT) cnt[CNT_TRIS]++;
        float t, u, v;
        if(triTest(r, p0, p1, p2, tmin, t, u, v)) {
            const uint32_t prim = __float_as_uint(p0.w);
            if(t < hit.t || (t == hit.t && t < wray.tmax() && (T.curInst < hit.inst || (T.curInst == hit.inst && prim < hit.prim)))) {
                hit.t = t; hit.u = u; hit.v = v; hit.inst = T.curInst; hit.prim = prim;
            }
        }
    }
}

// travPop: the lane has neither nodes nor primitives at hand: next entry of its stack.  True when the stack is empty (ray done).
template <class WR, class ST>
__device__ __forceinline__ bool travPop(Trav& T, const ST& stack, const WR& wray) {
    RayCtx& r = T.r;
    while(true) {
        if(T.sp == 0) return true;
        T.ng = stack.get(--T.sp);
        if(T.ng.x != kInvalid) break;
        // leave the instance: back to the world-space ray
        T.curInst = kInvalid;
        r.ox = wray.ox(); r.oy = wray.oy(); r.oz = wray.oz();
        if(!(T.ng.y & 0x1000u)) {   // a general instance: direction-derived constants come back from the stack
            r.octinv = T.ng.y & 7u; r.kx = (int)((T.ng.y >> 4) & 3u); r.ky = (int)((T.ng.y >> 6) & 3u); r.kz = (int)((T.ng.y >> 8) & 3u);
            const uint2 c2 = stack.get(--T.sp), c1 = stack.get(--T.sp), c0 = stack.get(--T.sp);
            r.ix = __uint_as_float(c0.x); r.iy = __uint_as_float(c0.y); r.iz = __uint_as_float(c1.x);
            r.Sx = __uint_as_float(c1.y); r.Sy = __uint_as_float(c2.x); r.Sz = __uint_as_float(c2.y);
        }
    }
    if(T.ng.x & kPrimGroupBit) { T.tg = T.ng; T.ng = make_uint2(0u, 0u); }   // a primitive group came off the stack
    return false;
}

// One step of the fixed order node -> primitives -> pop (a lane goes on with nodes only once its pending primitives are done);
// true when the traversal is complete.  wray: the world-space ray given to travInit (needed when an instance is entered or left).
// The primitives a lane has at hand (up to three triangles of a leaf, or the instances of a TLAS leaf until one is entered) are done
// in a tight loop, not one per step: measured against one per step (C2 4.19 -> 4.02 ms, C5 14.86 -> 14.16 ms), against looping over
// the nodes too (4.05) and against looping over the nodes only (4.25).  RG_PRIM_LOOP_POOL: the same for the pool kernel.
#ifndef RG_PRIM_LOOP_POOL
#define RG_PRIM_LOOP_POOL 0
#endif
template <bool COUNT, bool SPH, class WR, class ST>
__device__ __forceinline__ bool travStep(const TraceParams& P, Trav& T, const ST& stack, Hit& hit, const WR& wray, float tmin, uint32_t* cnt) {
    if((T.ng.y & 0xff000000u) && !T.tg.y) travNode<COUNT>(P, T, stack, hit, tmin, cnt);
    if(!SPH || RG_PRIM_LOOP_POOL) { while(T.tg.y) travPrim<COUNT, SPH>(P, T, stack, hit, wray, tmin, cnt); }
    else if(T.tg.y) travPrim<COUNT, SPH>(P, T, stack, hit, wray, tmin, cnt);
    if(!(T.ng.y & 0xff000000u) && !T.tg.y) return travPop(T, stack, wray);
    return false;
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

// miss.rmiss:38-74 with the constants of :78; normalize(0) is kept 0 (SURVEY hazard 7).  Of the shader's six pow() calls four need
// no pow: the integer powers are multiplication chains (x^80 = x^64 * x^16: 7 products, within 40 ulp of the exact power where the
// sun term is not clamped away, far below the binary16 store), pow(x, 0.5) is the correctly rounded square root, and the scatter
// colour depends on the light direction only -- the host evaluates it once per frame with the oracle's own libm (TraceParams).
__device__ __noinline__ V3 skyColor(V3 d, V3 lightDir, V3 scatterColor, bool strict) {
    const bool zero = d.x == 0.0f && d.y == 0.0f && d.z == 0.0f;
    const V3 rayDir = (zero && !strict) ? v3(0, 0, 0) : normalize(d);
    const float y = divShared(fabsf(d.y + 1.5f), 3.0f);
    const V3 sd = normalize(-lightDir) - rayDir;
    float sun = 1.0f - sqrtShared(dot(sd, sd));
    sun = clampf(sun, 0.0f, 2.0f);
    float glow = clampf(sun, 0.0f, 1.0f);
    const float s2 = sun * sun, s4 = s2 * s2, s8 = s4 * s4, s16 = s8 * s8, s32 = s16 * s16, s64 = s32 * s32;
    sun = s64 * s16;               // pow(sun, 80)
    sun *= 1000.0f;
    sun = clampf(sun, 0.0f, 16.0f);
    const float g2 = glow * glow, g4 = g2 * g2;
    glow = (g4 * g2) * 1.0f;       // pow(glow, 6)
    glow = powShared(glow, y);
    glow = clampf(glow, 0.0f, 1.0f);
    sun *= powShared(y * y, 1.0f / 1.65f);
    glow *= sqrtShared(y * y);          // pow(y * y, 1 / 2)
    sun += glow;
    const V3 sunColor = v3(1.0f, 0.6f, 0.05f) * sun;
    const float atmosphere = sqrtShared(1.0f - y);
    const V3 skyScatter = mix3(v3(0.2f, 0.4f, 0.8f), scatterColor, divShared(atmosphere, 1.3f));
    return sunColor + skyScatter;
}

enum { FR_GEN = 0, FR_SHI = 1 };
enum { ST_SHADOW_RET = 0, ST_TRY_REFLECT = 1, ST_REFLECT_RET = 2, ST_TRY_REFRACT = 3, ST_REFRACT_RET_FRONT = 4, ST_REFRACT_RET_BACK = 5, ST_COMBINE = 6 };

__device__ __forceinline__ uint32_t f2h(float f) { return (uint32_t)__half_as_ushort(__float2half_rn(f)); }
__device__ __forceinline__ uint2 packHalf4(float x, float y, float z, float w) { return make_uint2(f2h(x) | (f2h(y) << 16), f2h(z) | (f2h(w) << 16)); }

#ifndef RG_TRACE_MIN_BLOCKS
#define RG_TRACE_MIN_BLOCKS 5   // pool kernel: 96 registers (a few spills) x 20 warps / SM; swept 3..6 in round 2: C3 56.3 / 46.5 / 45.0 / 45.9 ms
#endif
#ifndef RG_FETCH_THRESHOLD
#define RG_FETCH_THRESHOLD 4     // lanes without a ray before the warp fetches from its ray queue (swept 2..16 on C3)
#endif
#ifndef RG_EXIT_THRESHOLD
#define RG_EXIT_THRESHOLD 28     // lanes without a ray, with the ray queue empty, before the warp leaves the traversal loop to shade (swept 12..31)
#endif
#ifndef RG_POOL_LIVE
#define RG_POOL_LIVE kPoolCtx    // contexts of a warp's pool that are actually used (<= kPoolCtx)
#endif
#ifndef RG_GRAB_MAX
#define RG_GRAB_MAX 32           // consecutive work items (samples of one 8x4 tile) a warp takes per atomic
#endif
#ifndef RG_REFILL_THRESHOLD
#define RG_REFILL_THRESHOLD 16   // free contexts before new work items are fetched (consecutive items = neighbouring pixels)
#endif

// Per-warp pool of ray-tree contexts.  A context = one pixel SAMPLE whose ray tree is walked depth-first in the shader's order
// (shadow -> reflection -> refraction) with ONE mutable payload, so every stale-state effect of the GLSL (SURVEY.md 8a hazards
// 1-6) is reproduced.  The hot part of a context (pending ray, closest hit, payload) lives in shared memory, structure-of-arrays
// so that 32 lanes holding 32 different contexts hit different banks; the suspended shader invocations (frames) live in global
// memory (kCtxQuads float4 per context, L2 resident).  Three byte queues hold context ids: rays waiting for traversal, hits
// waiting for shading, free contexts.
struct WarpPool {
    float ray[8][kPoolCtx];      // o.xyz, d.xyz, tmin, tmax
    uint32_t hit[5][kPoolCtx];   // t, u, v, instance, primitive
    float pay[6][kPoolCtx];      // hitValue.xyz, depth, curIOR, refDepth
    uint32_t sel[kPoolCtx];      // rayType | missIndex << 2 | rayKind << 4 | recDepth << 8 | frames << 16
    uint32_t pix[4][kPoolCtx];   // lx | ly << 16, pixel slot, sample, rays of this sample
    uint8_t rayQ[kPoolCtx], hitQ[kPoolCtx], freeQ[kPoolCtx];
};
constexpr uint32_t kNoCtx = 0xffu;
constexpr uint32_t kQMask = kPoolCtx - 1;
static_assert((kPoolCtx & (kPoolCtx - 1)) == 0 && kPoolCtx >= 32 && kPoolCtx <= 128, "kPoolCtx: power of two, 32..128");

// frame layout: 8 float4 per suspended invocation
enum { Q_ORG_T = 0, Q_DIR_FLAGS = 1, Q_N_RECDEPTH = 2, Q_DIFF_TRANSP = 3, Q_SPEC_REFL = 4, Q_BASE_ROUGH = 5, Q_RCOL_RDEPTH = 6, Q_IOR_EMIS = 7 };

// What the shaders do between two traceRayEXT calls of one pixel sample: the closest-hit or miss program of the ray that just
// came back, then the code after every traceRayEXT that returned (suspended frames, innermost first) up to the next traceRayEXT.
// Shared by both trace kernels; they differ only in where the state lives.  h: closest hit of the ray (ro, rd).  On return 1 the
// next ray is in (ro, rd, rtmin, rtmax) / (rayType, missIndex, rayKind); on return 2 the sample ended and its pixel share is stored.
// fr: the context's frames, 8 float4 each, + 2 float4 for the payload members that are only observable at recDepth 0.
// Where the frames of a context live: local memory (one context per lane) or the global pool, which is read and written
// through L2 only (ld/st.global.cg) so that the frames of 64 contexts per warp do not evict the BVH from L1.
struct FramesLocal {
    float4* p;
    __device__ __forceinline__ FramesLocal at(int i) const { return FramesLocal{p + i}; }
    __device__ __forceinline__ const float4 ld(int i) const { return p[i]; }
    __device__ __forceinline__ void st(int i, float4 v) const { p[i] = v; }
};
struct FramesPool {
    float4* p;
    __device__ __forceinline__ FramesPool at(int i) const { return FramesPool{p + i}; }
#ifdef RG_POOL_FRAMES_L1
    __device__ __forceinline__ const float4 ld(int i) const { return p[i]; }
    __device__ __forceinline__ void st(int i, float4 v) const { p[i] = v; }
#else
    __device__ __forceinline__ const float4 ld(int i) const { return __ldcg(p + i); }
    __device__ __forceinline__ void st(int i, float4 v) const { __stcg(p + i, v); }
#endif
};
struct ShadeConsts { V3 L; int maxRec; bool strictIeee; int numSamples; uint32_t S; };
constexpr uint32_t kMaxRaysPerSample = 1u << 20;   // watchdogs: far above anything a scene can need, they only turn a bug into a wrong
constexpr uint32_t kMaxStepsPerRay = 1u << 18;     // image instead of a hung GPU
struct PixInfo { uint32_t xy, pslot, sample, rays; };   // lx | ly << 16, pixel slot, sample index, rays traced for this sample so far
// SEQ: the numSamples samples of a pixel run one after the other in the same lane (lanes kernel, large frames): the sums of
// raygen.h:105-111 are carried in three quads behind the context's frames -- no scratch, no atomics, no fences.
template <bool COUNT, bool MULTI, bool SEQ, class FR>
__device__ __forceinline__ int shadeContext(const TraceParams& P, const ShadeConsts& K, const Hit& h, V3& ro, V3& rd, float& rtmin, float& rtmax, V3& hv,
                                            float& depth, float& curIOR, float& refDepth, int& rayType, int& missIndex, int& rayKind, int& recDepth, int& sp,
                                            const FR fr, const PixInfo& pix, uint32_t* skyLookups, uint32_t* cntT) {
    const bool found = h.inst != kInvalid;
    if(rayKind == CNT_PRIMARY && P.idInst && pix.sample == 0u) {
        const uint32_t xy = pix.xy;
        const int qx = (int)(xy & 0xffffu) - P.sx0, qy = (int)(xy >> 16) - P.sy0;
        if(qx >= 0 && qy >= 0 && qx < P.sw && qy < P.sh) { P.idInst[qy * P.sw + qx] = h.inst; P.idPrim[qy * P.sw + qx] = h.prim; }
    }
    bool issue = false;   // a new ray is pending
    if(found) {
        // closesthit.rchit:96-109
        const uint4 is3 = __ldg(reinterpret_cast<const uint4*>(P.instShade + h.inst) + 3);
        const uint32_t vtxOff = is3.x, idxOff = is3.y, matOff = is3.z;
        const uint32_t i0 = __ldg(P.indices + idxOff + 3 * h.prim), i1 = __ldg(P.indices + idxOff + 3 * h.prim + 1),
                       i2 = __ldg(P.indices + idxOff + 3 * h.prim + 2);
        const float4 v0p = __ldg(P.vertices + 2 * (size_t)(vtxOff + i0));
        const float4 n0 = __ldg(P.vertices + 2 * (size_t)(vtxOff + i0) + 1), n1 = __ldg(P.vertices + 2 * (size_t)(vtxOff + i1) + 1),
                     n2 = __ldg(P.vertices + 2 * (size_t)(vtxOff + i2) + 1);
        const float4* mp = P.materials + 4 * (size_t)(matOff + __float_as_uint(v0p.w));
        const float4 m0 = __ldg(mp), m1 = __ldg(mp + 1), m2 = __ldg(mp + 2), m3 = __ldg(mp + 3);
        V3 diffuse = v3(m0.x, m0.y, m0.z), specular = v3(m1.x, m1.y, m1.z);
        const float transparency = m0.w; float reflectivity = m1.w;
        const float roughness = m2.x, ior = m2.y, emission = m3.x;
        const uint32_t effectId = __float_as_uint(m2.z);
        uint32_t rayConsumption = __float_as_uint(m2.w);   // gpu_material.def: 1..5; clamped so that every reflection advances recDepth (frame stack bound)
        rayConsumption = rayConsumption < 1u ? 1u : (rayConsumption > (uint32_t)kMaxRecursions ? (uint32_t)kMaxRecursions : rayConsumption);

        const float b0 = 1.0f - h.u - h.v;
        const V3 origin = ro + rd * h.t;                                        // :114
        const V3 vn = v3(n0.x, n0.y, n0.z) * b0 + v3(n1.x, n1.y, n1.z) * h.u + v3(n2.x, n2.y, n2.z) * h.v;  // :117
        const float4* ow = reinterpret_cast<const float4*>(P.instShade + h.inst);
        const float4 o0 = __ldg(ow), o1 = __ldg(ow + 1), o2 = __ldg(ow + 2);
        V3 n = normalize(v3(o0.x * vn.x + o0.y * vn.y + o0.z * vn.z, o1.x * vn.x + o1.y * vn.y + o1.z * vn.z, o2.x * vn.x + o2.y * vn.y + o2.z * vn.z));  // :118-119

        if(effectId == 1u) {  // gridEffect, :74-91
            const float aa = divShared(refDepth + h.t + 8.0f, 30.0f);
            const float aa2 = aa / 2.0f;
            float minmod = glmin(fabsf(glmod((origin.x + 1000.0f) * 10.0f + aa2, 20.0f) - aa2), fabsf(glmod((origin.z + 1000.0f) * 10.0f + aa2, 20.0f) - aa2));
            if(minmod < aa2) {
                minmod -= aa2 - divShared(aa * aa, 3.0f);
                minmod *= divShared(3.0f, aa * aa);
                const float f = mixf(divShared(aa, 10.0f), 1.0f, minmod);
                diffuse = diffuse * f; specular = specular * f; reflectivity *= f;
            }
            if(glmod((origin.x + 1000.0f) * 5.0f, 20.0f) < 10.0f && glmod((origin.z + 1000.0f) * 5.0f, 20.0f) < 10.0f) reflectivity *= 1.5f;
        }

        if(rayType == RT_SHADOW_INTERNAL) {  // :125-152
            const float thick = clampf(h.t * (1.0f - transparency) * 10.0f, 0.0f, 1.0f);
            const V3 nd = normalize(v3(1.1f - diffuse.x, 1.1f - diffuse.y, 1.1f - diffuse.z));
            const V3 shadowCol = hv - mix3(v3(0, 0, 0), v3(nd.x + 0.1f, nd.y + 0.1f, nd.z + 0.1f), thick);
            if(recDepth < K.maxRec) {
                const FR f = fr.at(8 * sp++);
                f.st(Q_ORG_T, make_float4(shadowCol.x, shadowCol.y, shadowCol.z, 0.0f));
                f.st(Q_DIR_FLAGS, make_float4(rd.x, rd.y, rd.z, __int_as_float(FR_SHI)));
                f.st(Q_N_RECDEPTH, make_float4(n.x, n.y, n.z, __int_as_float(recDepth)));
                f.st(Q_IOR_EMIS, make_float4(ior, 0.0f, 0.0f, 0.0f));
                rayType = RT_SHADOW_TRACE; recDepth++;
                ro = origin; rtmin = 0.01f; rtmax = 1000.0f; missIndex = 0; rayKind = CNT_SHADOW;   // rd unchanged (T3)
                issue = true;
            } else {
                hv = shadowCol * 0.4f;
            }
        } else if(rayType == RT_SHADOW_TRACE) {  // :153-166
            if(transparency > 0.0f) {
                if(recDepth < K.maxRec) {   // T2: nothing to do after the child returns except recDepth--, which every
                    rayType = RT_SHADOW_INTERNAL; recDepth++;                 // resuming frame restores from its own copy
                    ro = origin; rtmin = 0.01f; rtmax = 1000.0f; missIndex = 1; rayKind = CNT_SHADOW;
                    issue = true;
                }
            } else {
                hv = hv * mixf(0.4f, 0.8f, clampf(logShared(h.t) / 8.0f, 0.0f, 1.0f));
            }
        } else {  // RT_GENERIC, :168-268
            if(COUNT) cntT[CNT_GENHIT]++;
            const bool frontFacing = dot(-rd, n) > 0.0f;
            if(!frontFacing) n = normalize(-n);
            const float ndl = dot(-K.L, n);
            V3 baseColor = diffuse * glmax(ndl, 0.2f);
            const FR f = fr.at(8 * sp++);
            const V3 rdIn = rd;
            int stage;
            bool shadowPending = false;
            if(ndl > 0.07f) {   // :188
                if(recDepth < K.maxRec) {
                    hv = v3(1.0f, 1.0f, 1.0f);
                    rayType = RT_SHADOW_TRACE;
                    ro = origin; rd = -K.L; rtmin = 0.1f; rtmax = 1000.0f; missIndex = 1; rayKind = CNT_SHADOW;
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
            const int flags = FR_GEN | (stage << 8) | ((frontFacing ? 1 : 0) << 16) | ((int)(rayConsumption & 0xffu) << 20);
            f.st(Q_ORG_T, make_float4(origin.x, origin.y, origin.z, h.t));
            f.st(Q_DIR_FLAGS, make_float4(rdIn.x, rdIn.y, rdIn.z, __int_as_float(flags)));
            f.st(Q_N_RECDEPTH, make_float4(n.x, n.y, n.z, __int_as_float(recDepth)));
            f.st(Q_DIFF_TRANSP, make_float4(diffuse.x, diffuse.y, diffuse.z, transparency));
            f.st(Q_SPEC_REFL, make_float4(specular.x, specular.y, specular.z, reflectivity));
            f.st(Q_BASE_ROUGH, make_float4(baseColor.x, baseColor.y, baseColor.z, roughness));
            f.st(Q_IOR_EMIS, make_float4(ior, emission, 0.0f, 0.0f));
            if(shadowPending) recDepth++;
        }
    } else {
        if(missIndex == 0) {  // miss.rmiss:76-83
            const V3 sky = skyColor(rd, K.L, v3(P.scatterColor[0], P.scatterColor[1], P.scatterColor[2]), K.strictIeee);
            hv = sky; depth = 10000.0f;
            if(sp == 0 && recDepth == 0) fr.st(kMaxFrames * 8 + 1, make_float4(sky.x, sky.y, sky.z, 0.0f));   // roughValue is only observable for a primary miss
        } else {              // shadowMiss.rmiss:33
            hv = v3(1.0f, 1.0f, 1.0f);
        }
    }

    // ---- resume suspended frames (the code after each traceRayEXT returns)
    bool ended = false;
    if(pix.rays > kMaxRaysPerSample) { sp = 0; issue = false; }   // watchdog: state machine stuck (never happens); end the sample
    while(!issue) {
        if(sp == 0) {   // raygen.h:105-111: the sample's trace returned
            const uint32_t xy = pix.xy, pslot = pix.pslot, sample = pix.sample;
            const uint32_t lx = xy & 0xffffu, ly = xy >> 16;
            if(P.tileCost) atomicAdd(P.tileCost + (pslot >> 5), pix.rays);
            const float4 cold0 = fr.ld(kMaxFrames * 8), cold1 = fr.ld(kMaxFrames * 8 + 1);   // normal + reflectContribution, roughValue
            V3 accColor = hv, accNormal = v3(cold0.x, cold0.y, cold0.z), accRough = v3(cold1.x, cold1.y, cold1.z);
            float accRoughA = cold1.w, accContrib = cold0.w, accDepth = depth;
            bool last = true;
            if(SEQ && K.S > 1u) {   // raygen.h:105-111: acc = 0; acc += sample i, in the loop's order
                float4 s0 = make_float4(0, 0, 0, 0), s1 = s0, s2 = s0;
                if(sample) { s0 = fr.ld(kCtxQuads); s1 = fr.ld(kCtxQuads + 1); s2 = fr.ld(kCtxQuads + 2); }
                s0 = make_float4(s0.x + hv.x, s0.y + hv.y, s0.z + hv.z, s0.w + cold0.w);
                s1 = make_float4(s1.x + cold0.x, s1.y + cold0.y, s1.z + cold0.z, s1.w + depth);
                s2 = make_float4(s2.x + cold1.x, s2.y + cold1.y, s2.z + cold1.z, s2.w + cold1.w);
                last = sample == K.S - 1u;
                if(last) {
                    accColor = v3(s0.x, s0.y, s0.z); accContrib = s0.w; accNormal = v3(s1.x, s1.y, s1.z); accDepth = s1.w;
                    accRough = v3(s2.x, s2.y, s2.z); accRoughA = s2.w;
                } else { fr.st(kCtxQuads, s0); fr.st(kCtxQuads + 1, s1); fr.st(kCtxQuads + 2, s2); }
            } else if(K.S > 1u) {   // park this sample; the lane finishing the pixel's last sample sums all of them in order
                float4* rec = P.sampleScratch + 3 * ((size_t)pslot * K.S + sample);
                __stcg(rec, make_float4(hv.x, hv.y, hv.z, cold0.w));
                __stcg(rec + 1, make_float4(cold0.x, cold0.y, cold0.z, depth));
                __stcg(rec + 2, cold1);
                // release: the three stores above are visible (in L2) before the count.  No acquire fence on the reading side: the
                // records are read past L1 (ld.global.cg) by loads that cannot issue before the atomic's result is known, and a
                // gpu-scope acquire / __threadfence would invalidate the SM's whole L1 (CCTL.IVALL) at the end of EVERY sample.
                uint32_t done;
                asm volatile("atom.release.gpu.global.add.u32 %0, [%1], 1;" : "=r"(done) : "l"(P.sampleDone + pslot) : "memory");
                last = done == K.S - 1u;
                if(last) {
                    P.sampleDone[pslot] = 0u;   // ready for the next frame
                    accColor = v3(0, 0, 0); accNormal = v3(0, 0, 0); accRough = v3(0, 0, 0); accRoughA = 0; accContrib = 0; accDepth = 0;
                    const float4* all = P.sampleScratch + 3 * (size_t)pslot * K.S;
                    for(uint32_t i = 0; i < K.S; ++i) {   // raygen.h:105-111 in the loop's order
                        const float4 a = __ldcg(all + 3 * i), b = __ldcg(all + 3 * i + 1), cc = __ldcg(all + 3 * i + 2);
                        accColor = accColor + v3(a.x, a.y, a.z); accNormal = accNormal + v3(b.x, b.y, b.z);
                        accRough = accRough + v3(cc.x, cc.y, cc.z); accRoughA += cc.w; accContrib += a.w; accDepth += b.w;
                    }
                }
            }
            if(last) {      // raygen.h:114 + raygen.rgen:35-38
                // x / numSamples: for a power of two the product with the reciprocal is the same correctly rounded value.  (The other
                // sample counts take one rolled loop over the twelve values: this code runs once per pixel but sits in the instruction
                // caches of every warp.)
                const float inv = (float)K.numSamples;
                float q[12] = {accColor.x, accColor.y, accColor.z, accContrib, accNormal.x, accNormal.y, accNormal.z, logShared(accDepth) * 0.25f,
                               accRough.x, accRough.y, accRough.z, accRoughA};
                if((K.S & (K.S - 1u)) == 0u) {
                    const float rinv = 1.0f / inv;
#pragma unroll
                    for(int k = 0; k < 12; ++k) q[k] *= rinv;
                } else {
#pragma unroll 1
                    for(int k = 0; k < 12; ++k) q[k] = divShared(q[k], inv);
                }
                const uint2 ob = packHalf4(q[0], q[1], q[2], q[3]), on = packHalf4(q[4], q[5], q[6], q[7]), orr = packHalf4(q[8], q[9], q[10], q[11]);
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
            ended = true;
            break;
        }
        const FR f = fr.at(8 * (sp - 1));
        const float4 qDir = f.ld(Q_DIR_FLAGS), qN = f.ld(Q_N_RECDEPTH), qOrg = f.ld(Q_ORG_T), qIor = f.ld(Q_IOR_EMIS);
        const int flags = __float_as_int(qDir.w);
        recDepth = __float_as_int(qN.w);   // undoes every recDepth++ / += rayConsumption below this frame
        if((flags & 0xff) == FR_SHI) {   // closesthit.rchit:136-146, after T3 returned
            V3 shadowCol = v3(qOrg.x, qOrg.y, qOrg.z);
            if(depth < 1000.0f) {
                hv = hv * shadowCol;
            } else {
                const V3 dir = refract3(v3(qDir.x, qDir.y, qDir.z), v3(qN.x, qN.y, qN.z), qIor.x);
                const float dp = dot(K.L, dir);
                const float dp2 = dp * dp;
                shadowCol = shadowCol * (dp2 * dp2 * dp + 0.75f);   // pow(x, 5)
                (*skyLookups)++;   // T4: cull mask 0 -> always miss 0
                const V3 sky = skyColor(-dir, K.L, v3(P.scatterColor[0], P.scatterColor[1], P.scatterColor[2]), K.strictIeee);
                depth = 10000.0f;
                hv = shadowCol + sky * 0.1f;
            }
            sp--;
            continue;
        }
        const float4 qDiff = f.ld(Q_DIFF_TRANSP), qSpec = f.ld(Q_SPEC_REFL);
        int stage = (flags >> 8) & 0xff;
        // the reflection result is stored when the refraction ray is issued: earlier stages compute it below (and never read the slot)
        float4 qBase = f.ld(Q_BASE_ROUGH), qRcol = stage >= ST_REFRACT_RET_FRONT ? f.ld(Q_RCOL_RDEPTH) : make_float4(1.0f, 1.0f, 1.0f, 0.0f);
        const bool frontFacing = ((flags >> 16) & 1) != 0;
        const int rc = (flags >> 20) & 0xff;
        const V3 D = v3(qDir.x, qDir.y, qDir.z), n = v3(qN.x, qN.y, qN.z);
        if(stage == ST_SHADOW_RET) {   // :196-197 then :254-255
            const V3 diffuse = v3(qDiff.x, qDiff.y, qDiff.z);
            const V3 base = v3(qBase.x, qBase.y, qBase.z) * hv + diffuse * qIor.y;
            qBase.x = base.x; qBase.y = base.y; qBase.z = base.z;
            stage = ST_TRY_REFLECT;
        }
        if(stage == ST_TRY_REFLECT) {  // :207-221
            if(recDepth < K.maxRec && qSpec.w > 0.0f) {
                recDepth += rc; refDepth += qOrg.w;
                ro = v3(qOrg.x, qOrg.y, qOrg.z); rd = reflect3(D, n); rtmin = 0.01f; rtmax = 1000.0f;
                rayType = RT_GENERIC; missIndex = 0; rayKind = CNT_REFLECT;
                f.st(Q_DIR_FLAGS, make_float4(qDir.x, qDir.y, qDir.z, __int_as_float((flags & ~0xff00) | (ST_REFLECT_RET << 8))));
                f.st(Q_BASE_ROUGH, qBase);
                issue = true;
                break;
            }
            qRcol = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
            stage = ST_TRY_REFRACT;
        }
        if(stage == ST_REFLECT_RET) {
            qRcol = make_float4(hv.x * qSpec.x, hv.y * qSpec.y, hv.z * qSpec.z, depth);
            stage = ST_TRY_REFRACT;
        }
        V3 refractColor = v3(1.0f, 1.0f, 1.0f);
        if(stage == ST_TRY_REFRACT) {  // :224-251
            if(recDepth < K.maxRec && qDiff.w > 0.0f) {
                const float ior = qIor.x;
                const float eta = frontFacing ? divShared(curIOR, ior) : ior / 1.0f;
                recDepth++;
                curIOR = frontFacing ? ior : 1.0f;
                ro = v3(qOrg.x, qOrg.y, qOrg.z); rd = refract3(D, n, eta); rtmin = 0.01f; rtmax = 1000.0f;
                rayType = RT_GENERIC; missIndex = 0; rayKind = CNT_REFRACT;
                f.st(Q_DIR_FLAGS, make_float4(qDir.x, qDir.y, qDir.z,
                                              __int_as_float((flags & ~0xff00) | ((frontFacing ? ST_REFRACT_RET_FRONT : ST_REFRACT_RET_BACK) << 8))));
                f.st(Q_BASE_ROUGH, qBase); f.st(Q_RCOL_RDEPTH, qRcol);
                issue = true;
                break;
            }
            stage = ST_COMBINE;
        } else if(stage == ST_REFRACT_RET_FRONT) {
            refractColor = hv;
        } else if(stage == ST_REFRACT_RET_BACK) {
            const V3 diffuse = v3(qDiff.x, qDiff.y, qDiff.z);
            refractColor = mix3(v3(1, 1, 1), diffuse, logShared(1.0f + qOrg.w)) * hv;
        }
        // combine, :254-267
        {
            const V3 base = v3(qBase.x, qBase.y, qBase.z);
            const V3 reflectColor = v3(qRcol.x, qRcol.y, qRcol.z);
            const float transparency = qDiff.w, reflectivity = qSpec.w, roughness = qBase.w;
            const float totalContrib = glmax(transparency, reflectivity);
            float weight = divShared(reflectivity, transparency + reflectivity);
            if(!K.strictIeee && (transparency + reflectivity) == 0.0f) weight = 0.0f;   // SURVEY hazard 8
            const V3 roughCol = mix3(refractColor, reflectColor, weight);
            hv = mix3(base, roughCol, totalContrib);
            if(recDepth == 0) {
                hv = base;
                fr.st(kMaxFrames * 8, make_float4(n.x, n.y, n.z, totalContrib));
                fr.st(kMaxFrames * 8 + 1, make_float4(roughCol.x, roughCol.y, roughCol.z, glmin(divShared(qRcol.w, 50.0f) * roughness, divShared(roughness, 2.1f))));
            }
            depth = qOrg.w;
            sp--;
        }
    }
    return ended ? 2 : 1;
}


// The warp is a small scheduler over its pool (the decoupling the per-lane recursion lacked: with one context per lane a warp
// waited for its longest ray and shaded with a handful of lanes, BASELINE config 3 ran at 7 of 32 lanes):
//   TRAVERSE  lanes hold rays in flight (registers + stack).  A lane whose ray is done parks the hit in the pool, and as soon as
//             RG_FETCH_THRESHOLD lanes are empty they take the next rays from the ray queue (vote + prefix rank).  The loop is
//             left when 32 hits wait, or when the ray queue is dry and RG_EXIT_THRESHOLD lanes idle; rays in flight stay in flight.
//   SHADE     up to 32 waiting hits are shaded together, one per lane: hit / miss program, then the code after the traceRayEXT
//             that returned, up to the next traceRayEXT (whose ray goes to the ray queue) or the end of the sample.
//   REFILL    free contexts take the next work items (pixel samples) from the global counter.
template <bool COUNT, bool MULTI>
__global__ void __launch_bounds__(128, RG_TRACE_MIN_BLOCKS) k_trace_pool(const TraceParams P) {
    __shared__ float s_ubo[48];
    __shared__ WarpPool s_pool[4];
    __shared__ uint32_t s_cnt[5][128];    // per lane: rays by kind + sky lookups
    const uint32_t tid = threadIdx.x;
    if(tid < 48) s_ubo[tid] = P.ubo[tid];
#pragma unroll
    for(int k = 0; k < 5; ++k) s_cnt[k][tid] = 0u;
    __syncthreads();
    const float* VI = s_ubo;       // viewInverse, column-major
    const float* PI = s_ubo + 16;  // projInverse
    const int numSamples = __float_as_int(s_ubo[35]);
    const V3 L = v3(s_ubo[36], s_ubo[37], s_ubo[38]);
    int maxRec = __float_as_int(s_ubo[39]);
    maxRec = maxRec > kMaxRecursions ? kMaxRecursions : maxRec;
    ShadeConsts K;
    K.L = L; K.maxRec = maxRec; K.strictIeee = (P.flags & RG_STRICT_IEEE) != 0; K.numSamples = numSamples; K.S = (uint32_t)numSamples;

    const uint32_t lane = tid & 31;
    const uint32_t ltMask = (1u << lane) - 1u;
    const uint32_t tilesX = (P.dw + 7) / 8, tilesY = (P.dh + 3) / 4;
    const uint32_t nTiles = tilesX * tilesY;
    // this rank's share: chunks of kChunkTiles tiles dealt round-robin (world == 1: everything)
    const uint32_t nChunks = (nTiles + kChunkTiles - 1) / kChunkTiles;
    const uint32_t myChunks = nChunks > P.rank ? (nChunks - P.rank + P.world - 1) / P.world : 0u;
    const uint32_t S = (uint32_t)numSamples;
    const uint32_t total = myChunks * kChunkTiles * 32u * S;   // one work item per pixel SAMPLE

    WarpPool& W = s_pool[tid >> 5];
    float4* const ctxMem = P.ctxPool + (size_t)(blockIdx.x * 4u + (tid >> 5)) * kPoolCtx * kCtxQuads;
    for(uint32_t i = lane; i < (uint32_t)kPoolCtx; i += 32u) W.freeQ[i] = (uint8_t)i;
    __syncwarp();
    // warp-uniform scheduler state
    uint32_t rayHead = 0, rayCount = 0, hitHead = 0, hitCount = 0, freeHead = 0, freeCount = RG_POOL_LIVE;
    bool exhausted = false;
    // the lane's ray in flight
    uint32_t myCtx = kNoCtx, steps = 0;
    float tmin = 0.0f;
    uint2 stack[kStackSize];
    Trav T;
    T.ng = make_uint2(0u, 0u); T.tg = make_uint2(0u, 0u); T.sp = 0; T.curInst = kInvalid;
    Hit hit;
    hit.t = 0.0f; hit.u = 0.0f; hit.v = 0.0f; hit.inst = kInvalid; hit.prim = kInvalid;
    uint32_t cntT[CNT_N];
    if(COUNT) {
#pragma unroll
        for(int k = 0; k < CNT_N; ++k) cntT[k] = 0;
    }
    // camera origin: viewInverse * (0,0,0,1) in GLM order (c0*0 + c1*0) + (c2*0 + c3*1)
    const V3 camO = v3((VI[0] * 0.0f + VI[4] * 0.0f) + (VI[8] * 0.0f + VI[12] * 1.0f), (VI[1] * 0.0f + VI[5] * 0.0f) + (VI[9] * 0.0f + VI[13] * 1.0f),
                       (VI[2] * 0.0f + VI[6] * 0.0f) + (VI[10] * 0.0f + VI[14] * 1.0f));

    uint32_t guard = 0;
    while(true) {
        if(++guard > (1u << 22)) { if(lane == 0) atomicAdd(P.counters + 15, 1ull); break; }   // scheduler watchdog (never trips; a trip shows in rg_timings)
        const uint32_t nActive = __popc(__ballot_sync(0xffffffffu, myCtx != kNoCtx));

        // ---- REFILL: free contexts take new work items
        if(!exhausted && freeCount && (freeCount >= RG_REFILL_THRESHOLD || (rayCount == 0u && hitCount == 0u))) {
            const uint32_t n = freeCount < (uint32_t)RG_GRAB_MAX ? freeCount : (uint32_t)RG_GRAB_MAX;
            uint32_t basew = 0;
            if(lane == 0) basew = atomicAdd(P.workCounter, n);
            basew = __shfl_sync(0xffffffffu, basew, 0);
            const uint32_t c = lane < n ? W.freeQ[(freeHead + lane) & kQMask] : kNoCtx;
            freeHead += n; freeCount -= n;
            bool ok = false;
            const uint32_t w = basew + lane;
            if(c != kNoCtx && w < total) {
                // w = ((slot position * S) + sample) * 32 + pixel in tile: consecutive items = one sample index of one 8x4 tile
                uint32_t sample, tcol;
                const uint32_t l = w & 31u, pos = divMagic(w >> 5, S, P.magicS, sample);
                const uint32_t j = P.tileOrder ? __ldg(P.tileOrder + pos) : pos;
                const uint32_t tile = ((j / kChunkTiles) * P.world + P.rank) * kChunkTiles + (j % kChunkTiles);
                const uint32_t trow = divMagic(tile, tilesX, P.magicTilesX, tcol);
                const uint32_t lx = P.dx0 + tcol * 8u + (l & 7u), ly = P.dy0 + trow * 4u + (l >> 3);   // frame coordinates
                if(tile < nTiles && lx < P.dx0 + P.dw && ly < P.dy0 + P.dh) {
                    ok = true;
                    W.pix[0][c] = lx | (ly << 16); W.pix[1][c] = j * 32u + l; W.pix[2][c] = sample; W.pix[3][c] = 1u;
                    // raygen.h:80-100
                    const float2 off = aaOffset(numSamples, (int)sample);
                    const float pcx = (float)lx + 0.5f + off.x, pcy = (float)ly + 0.5f + off.y;
                    const float ddx = divShared(pcx, (float)P.W) * 2.0f - 1.0f, ddy = divShared(pcy, (float)P.H) * 2.0f - 1.0f;
                    // target = projInverse * (d.x, d.y, 1, 1); direction = viewInverse * (normalize(target.xyz), 0)
                    const V3 tgt = v3((PI[0] * ddx + PI[4] * ddy) + (PI[8] + PI[12]), (PI[1] * ddx + PI[5] * ddy) + (PI[9] + PI[13]),
                                      (PI[2] * ddx + PI[6] * ddy) + (PI[10] + PI[14]));
                    const V3 nt = normalize(tgt);
                    W.ray[0][c] = camO.x; W.ray[1][c] = camO.y; W.ray[2][c] = camO.z;
                    W.ray[3][c] = (VI[0] * nt.x + VI[4] * nt.y) + (VI[8] * nt.z);
                    W.ray[4][c] = (VI[1] * nt.x + VI[5] * nt.y) + (VI[9] * nt.z);
                    W.ray[5][c] = (VI[2] * nt.x + VI[6] * nt.y) + (VI[10] * nt.z);
                    W.ray[6][c] = 0.001f; W.ray[7][c] = 10000.0f;
                    W.pay[0][c] = 0.0f; W.pay[1][c] = 0.0f; W.pay[2][c] = 0.0f; W.pay[3][c] = 0.0f; W.pay[4][c] = 1.0f; W.pay[5][c] = 0.0f;
                    W.sel[c] = (uint32_t)(RT_GENERIC | (0 << 2) | (CNT_PRIMARY << 4));   // recDepth 0, no frames
                    float4* cold = ctxMem + (size_t)c * kCtxQuads + kMaxFrames * 8;
                    __stcg(cold, make_float4(0, 0, 0, 0)); __stcg(cold + 1, make_float4(0, 0, 0, 0));
                    s_cnt[CNT_PRIMARY][tid]++;
                }
            }
            const uint32_t mOk = __ballot_sync(0xffffffffu, ok), mBack = __ballot_sync(0xffffffffu, c != kNoCtx && !ok);
            if(ok) W.rayQ[(rayHead + rayCount + __popc(mOk & ltMask)) & kQMask] = (uint8_t)c;
            rayCount += __popc(mOk);
            if(c != kNoCtx && !ok) W.freeQ[(freeHead + freeCount + __popc(mBack & ltMask)) & kQMask] = (uint8_t)c;   // padding pixels: try the next item
            freeCount += __popc(mBack);
            if(basew + n >= total) exhausted = true;
            __syncwarp();
            continue;
        }

        // ---- SHADE one batch: closest hit or miss, then resume suspended frames until a new ray is issued
        if(hitCount >= 32u || (hitCount && rayCount == 0u && nActive <= 32u - RG_EXIT_THRESHOLD)) {
            const uint32_t nb = hitCount < 32u ? hitCount : 32u;
            const uint32_t c = lane < nb ? W.hitQ[(hitHead + lane) & kQMask] : kNoCtx;
            hitHead += nb; hitCount -= nb;
            int outcome = 0;   // 1: a new ray is pending, 2: the sample ended
            if(c != kNoCtx) {
                float4* const fr = ctxMem + (size_t)c * kCtxQuads;
                Hit h;
                h.t = __uint_as_float(W.hit[0][c]); h.u = __uint_as_float(W.hit[1][c]); h.v = __uint_as_float(W.hit[2][c]); h.inst = W.hit[3][c]; h.prim = W.hit[4][c];
                V3 ro = v3(W.ray[0][c], W.ray[1][c], W.ray[2][c]), rd = v3(W.ray[3][c], W.ray[4][c], W.ray[5][c]);
                float rtmin = 0.0f, rtmax = 0.0f;
                V3 hv = v3(W.pay[0][c], W.pay[1][c], W.pay[2][c]);
                float depth = W.pay[3][c], curIOR = W.pay[4][c], refDepth = W.pay[5][c];
                const uint32_t sel = W.sel[c];
                int rayType = (int)(sel & 3u), missIndex = (int)((sel >> 2) & 3u), rayKind = (int)((sel >> 4) & 15u), recDepth = (int)((sel >> 8) & 255u);
                int sp = (int)((sel >> 16) & 255u);
                PixInfo pix;
                pix.xy = W.pix[0][c]; pix.pslot = W.pix[1][c]; pix.sample = W.pix[2][c]; pix.rays = W.pix[3][c];
                const bool ended = shadeContext<COUNT, MULTI, false>(P, K, h, ro, rd, rtmin, rtmax, hv, depth, curIOR, refDepth, rayType, missIndex, rayKind, recDepth, sp,
                                                              FramesPool{fr}, pix, &s_cnt[CNT_SKY][tid], cntT) == 2;
                if(ended) {
                    outcome = 2;
                } else {   // the next traceRayEXT of this sample
                    outcome = 1;
                    s_cnt[rayKind][tid]++;
                    W.pix[3][c]++;
                    W.ray[0][c] = ro.x; W.ray[1][c] = ro.y; W.ray[2][c] = ro.z; W.ray[3][c] = rd.x; W.ray[4][c] = rd.y; W.ray[5][c] = rd.z;
                    W.ray[6][c] = rtmin; W.ray[7][c] = rtmax;
                    W.pay[0][c] = hv.x; W.pay[1][c] = hv.y; W.pay[2][c] = hv.z; W.pay[3][c] = depth; W.pay[4][c] = curIOR; W.pay[5][c] = refDepth;
                    W.sel[c] = (uint32_t)(rayType | (missIndex << 2) | (rayKind << 4) | (recDepth << 8) | (sp << 16));
                }
            }
            const uint32_t mNew = __ballot_sync(0xffffffffu, outcome == 1), mEnd = __ballot_sync(0xffffffffu, outcome == 2);
            if(outcome == 1) W.rayQ[(rayHead + rayCount + __popc(mNew & ltMask)) & kQMask] = (uint8_t)c;
            rayCount += __popc(mNew);
            if(outcome == 2) W.freeQ[(freeHead + freeCount + __popc(mEnd & ltMask)) & kQMask] = (uint8_t)c;
            freeCount += __popc(mEnd);
            __syncwarp();
            continue;
        }

        if(nActive == 0u && rayCount == 0u) {   // nothing in flight, nothing to trace, nothing to shade (or it was shaded above)
            if(exhausted || freeCount == 0u) break;
            continue;   // not exhausted: the refill above takes whatever is free
        }

        // ---- TRAVERSE
        while(true) {
            const uint32_t mEmpty = __ballot_sync(0xffffffffu, myCtx == kNoCtx);
            uint32_t nEmpty = __popc(mEmpty);
            if(rayCount && nEmpty >= RG_FETCH_THRESHOLD) {   // hand the next rays to the empty lanes
                const uint32_t take = nEmpty < rayCount ? nEmpty : rayCount;
                const uint32_t r = __popc(mEmpty & ltMask);
                if(myCtx == kNoCtx && r < take) {
                    myCtx = W.rayQ[(rayHead + r) & kQMask];
                    tmin = W.ray[6][myCtx];
                    steps = 0;
                    travInit<false>(P, T, hit, W.ray[0][myCtx], W.ray[1][myCtx], W.ray[2][myCtx], W.ray[3][myCtx], W.ray[4][myCtx], W.ray[5][myCtx], W.ray[7][myCtx]);
                }
                rayHead += take; rayCount -= take; nEmpty -= take;
            }
            if(nEmpty == 32u || hitCount >= 32u) break;
            if(rayCount == 0u && nEmpty >= RG_EXIT_THRESHOLD && hitCount) break;
            if(!exhausted && freeCount >= RG_REFILL_THRESHOLD && rayCount == 0u) break;
            bool done = false;
            // (two steps per round of this loop -- half the votes and queue bookkeeping -- were measured: C3 43.2 -> 55.4 ms)
            if(myCtx != kNoCtx) done = travStep<COUNT, true>(P, T, StackLocal{stack}, hit, WorldRayPool{&W.ray[0][myCtx]}, tmin, cntT) || ++steps > kMaxStepsPerRay;
            const uint32_t mDone = __ballot_sync(0xffffffffu, done);
            if(mDone) {
                if(done) {   // park the closest hit; the lane is free for the next ray
                    W.hit[0][myCtx] = __float_as_uint(hit.t); W.hit[1][myCtx] = __float_as_uint(hit.u); W.hit[2][myCtx] = __float_as_uint(hit.v);
                    W.hit[3][myCtx] = hit.inst; W.hit[4][myCtx] = hit.prim;
                    W.hitQ[(hitHead + hitCount + __popc(mDone & ltMask)) & kQMask] = (uint8_t)myCtx;
                    myCtx = kNoCtx;
                }
                hitCount += __popc(mDone);
            }
        }
        __syncwarp();
    }

    // ---- ray counters: warp reduce, one atomic per warp and counter
#pragma unroll
    for(int k = 0; k < CNT_N; ++k) {
        if(!COUNT && k >= CNT_NODES) break;
        unsigned long long v = k < CNT_NODES ? (unsigned long long)s_cnt[k][tid] : (unsigned long long)cntT[k];
#pragma unroll
        for(int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if(lane == 0 && v) atomicAdd(P.counters + k, v);
    }
}


#ifndef RG_LANES_REFILL
#define RG_LANES_REFILL 32       // idle lanes before a warp of the lanes kernel fetches new work items (swept 2..32: C2 7.9 ms at 2, 6.6 at 8, 5.4 at 20-28, 5.2 at 32)
#endif
#ifndef RG_LANES_SHSTACK
#define RG_LANES_SHSTACK 0       // lanes kernel: traversal stack entries per ray kept in shared memory (0: all of them in local memory).  Measured: 8 / 12 / 16
                                 // entries C2 4.21 ms against 3.56 (two predicated accesses per push / pop, and L1 loses what the carve-out takes): off
#endif
#ifndef RG_LANES_MIN_BLOCKS
#define RG_LANES_MIN_BLOCKS 7   // 72 registers (~150 B of spills); swept 5..8 with the binary16 node test: C2 3.74 / 3.68 / 3.55 / 3.66 ms, C5 13.19 / 12.93 / 12.48 / 12.73 ms
#endif
// The second scheduler, for COHERENT workloads: one context per lane, state in registers, frames in local memory.  A warp takes 32
// consecutive work items (one sample index of one 8x4 tile), so its lanes trace neighbouring rays and then run the same shader
// stage together; every lane traverses its ray to completion, then all shade.  Lanes whose sample ended are refilled from the
// global counter 8 or more at a time.  On the example scene (BASELINE config 2) this keeps 21 of 32 lanes busy at 8 CTAs / SM and
// beats the pool scheduler; on incoherent bounces (config 3) a warp waits for its longest ray and the pool wins.  rg_render
// picks per scene by timing both (rg_api.cu).
template <bool COUNT, bool MULTI, bool SEQ>
__global__ void __launch_bounds__(128, RG_LANES_MIN_BLOCKS) k_trace_lanes(const TraceParams P) {
    static_assert(!SEQ || (RG_LANES_REFILL == 32 && RG_GRAB_MAX == 32), "SEQ: a warp owns one whole tile at a time");
    __shared__ float s_ubo[48];
    __shared__ uint32_t s_cnt[5][128];    // per lane: rays by kind + sky lookups
    const uint32_t tid = threadIdx.x;
    if(tid < 48) s_ubo[tid] = P.ubo[tid];
#pragma unroll
    for(int k = 0; k < 5; ++k) s_cnt[k][tid] = 0u;
    __syncthreads();
    const float* VI = s_ubo;       // viewInverse, column-major
    const float* PI = s_ubo + 16;  // projInverse
    const int numSamples = __float_as_int(s_ubo[35]);
    int maxRec = __float_as_int(s_ubo[39]);
    maxRec = maxRec > kMaxRecursions ? kMaxRecursions : maxRec;
    ShadeConsts K;
    K.L = v3(s_ubo[36], s_ubo[37], s_ubo[38]); K.maxRec = maxRec; K.strictIeee = (P.flags & RG_STRICT_IEEE) != 0; K.numSamples = numSamples; K.S = (uint32_t)numSamples;

    const uint32_t lane = tid & 31;
    const uint32_t tilesX = (P.dw + 7) / 8, tilesY = (P.dh + 3) / 4;
    const uint32_t nTiles = tilesX * tilesY;
    const uint32_t nChunks = (nTiles + kChunkTiles - 1) / kChunkTiles;
    const uint32_t myChunks = nChunks > P.rank ? (nChunks - P.rank + P.world - 1) / P.world : 0u;
    const uint32_t S = (uint32_t)numSamples;
    // one work item per pixel SAMPLE; SEQ: per pixel (its samples follow each other in the lane that took it)
    const uint32_t total = myChunks * kChunkTiles * 32u * (SEQ ? 1u : S);

    float4 frames[kCtxQuads + (SEQ ? 3 : 0)];   // suspended shader invocations (local memory) [+ the pixel's sums]
#if RG_LANES_SHSTACK
    __shared__ uint2 s_stack[RG_LANES_SHSTACK * 128];
    uint2 stackLocal[kStackSize - RG_LANES_SHSTACK];
    const StackHybrid<RG_LANES_SHSTACK> stack{s_stack + tid, stackLocal};
#else
    uint2 stackLocal[kStackSize];
    const StackLocal stack{stackLocal};
#endif
    uint32_t cntT[CNT_N];
    if(COUNT) {
#pragma unroll
        for(int k = 0; k < CNT_N; ++k) cntT[k] = 0;
    }
    bool exhausted = false, busy = false;
    bool pixValid = false;   // SEQ: the lane holds a pixel whose samples are not all traced yet
    PixInfo pix;
    pix.xy = 0; pix.pslot = 0; pix.sample = 0; pix.rays = 0;
    V3 hv = v3(0, 0, 0), ro = v3(0, 0, 0), rd = v3(0, 0, 0);
    float depth = 0, curIOR = 1.0f, refDepth = 0, rtmin = 0, rtmax = 0;
    int recDepth = 0, sp = 0, rayType = RT_GENERIC, missIndex = 0, rayKind = CNT_PRIMARY;
    const V3 camO = v3((VI[0] * 0.0f + VI[4] * 0.0f) + (VI[8] * 0.0f + VI[12] * 1.0f), (VI[1] * 0.0f + VI[5] * 0.0f) + (VI[9] * 0.0f + VI[13] * 1.0f),
                       (VI[2] * 0.0f + VI[6] * 0.0f) + (VI[10] * 0.0f + VI[14] * 1.0f));
#ifdef RG_DEBUG_TAIL   // developer build: how long every warp of the persistent grid had work (printed by rg_get_timings under RGB200_DEBUG_TAIL)
    unsigned long long tWarp0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tWarp0));
#endif
    // raygen.h:80-100: the primary ray of sample `sample` of the lane's pixel (pix.xy)
    auto startSample = [&](uint32_t sample) {
        const uint32_t lx = pix.xy & 0xffffu, ly = pix.xy >> 16;
        busy = true;
        pix.sample = sample; pix.rays = 1u;
        const float2 off = aaOffset(numSamples, (int)sample);
        const float pcx = (float)lx + 0.5f + off.x, pcy = (float)ly + 0.5f + off.y;
        const float ddx = divShared(pcx, (float)P.W) * 2.0f - 1.0f, ddy = divShared(pcy, (float)P.H) * 2.0f - 1.0f;
        // target = projInverse * (d.x, d.y, 1, 1); direction = viewInverse * (normalize(target.xyz), 0)
        const V3 tgt = v3((PI[0] * ddx + PI[4] * ddy) + (PI[8] + PI[12]), (PI[1] * ddx + PI[5] * ddy) + (PI[9] + PI[13]),
                          (PI[2] * ddx + PI[6] * ddy) + (PI[10] + PI[14]));
        const V3 nt = normalize(tgt);
        rd = v3((VI[0] * nt.x + VI[4] * nt.y) + (VI[8] * nt.z), (VI[1] * nt.x + VI[5] * nt.y) + (VI[9] * nt.z), (VI[2] * nt.x + VI[6] * nt.y) + (VI[10] * nt.z));
        ro = camO; rtmin = 0.001f; rtmax = 10000.0f;
        rayType = RT_GENERIC; missIndex = 0; rayKind = CNT_PRIMARY;
        hv = v3(0, 0, 0); depth = 0; refDepth = 0; curIOR = 1.0f; recDepth = 0; sp = 0;
        frames[kMaxFrames * 8] = make_float4(0, 0, 0, 0); frames[kMaxFrames * 8 + 1] = make_float4(0, 0, 0, 0);
        s_cnt[CNT_PRIMARY][tid]++;
    };

    while(true) {
        uint32_t idle = __ballot_sync(0xffffffffu, !busy);
        if(SEQ && idle == 0xffffffffu && __any_sync(0xffffffffu, pixValid)) {
            // the whole warp finished sample i of its tile: sample i + 1 of the same 32 pixels, again in lock-step
            if(pixValid) {
                if(pix.sample + 1u < S) startSample(pix.sample + 1u);
                else pixValid = false;
            }
            idle = __ballot_sync(0xffffffffu, !busy);
        }
        // ---- refill idle lanes (warp vote + prefix compaction over one atomic)
        while(!exhausted && __popc(idle) >= RG_LANES_REFILL) {
            // at most RG_GRAB_MAX consecutive work items per grab: the samples of one (possibly very expensive) tile spread over several warps
            const int nIdle = __popc(idle), n = nIdle < RG_GRAB_MAX ? nIdle : RG_GRAB_MAX, leader = __ffs(idle) - 1;
            uint32_t basew = 0;
            if((int)lane == leader) basew = atomicAdd(P.workCounter, (uint32_t)n);
            basew = __shfl_sync(0xffffffffu, basew, leader);
            const uint32_t rank = __popc(idle & ((1u << lane) - 1u));
            const bool served = !busy && rank < (uint32_t)n;
            if(served) {
                const uint32_t w = basew + rank;
                if(w < total) {
                    // w = ((slot position * S) + sample) * 32 + pixel in tile: consecutive items = one sample index of one 8x4 tile
                    uint32_t sample = 0u, tcol;
                    const uint32_t l = w & 31u, pos = SEQ ? (w >> 5) : divMagic(w >> 5, S, P.magicS, sample);
                    const uint32_t j = P.tileOrder ? __ldg(P.tileOrder + pos) : pos;
                    const uint32_t tile = ((j / kChunkTiles) * P.world + P.rank) * kChunkTiles + (j % kChunkTiles);
                    const uint32_t trow = divMagic(tile, tilesX, P.magicTilesX, tcol);
                    const uint32_t lx = P.dx0 + tcol * 8u + (l & 7u), ly = P.dy0 + trow * 4u + (l >> 3);   // frame coordinates
                    if(tile < nTiles && lx < P.dx0 + P.dw && ly < P.dy0 + P.dh) {
                        pix.xy = lx | (ly << 16); pix.pslot = j * 32u + l;
                        pixValid = SEQ;
                        startSample(sample);
                    }
                }
            }
            if(basew + (uint32_t)n >= total) exhausted = true;
            idle &= ~__ballot_sync(0xffffffffu, served);   // padding pixels stay idle until the next round of the outer loop
        }
        if(!__any_sync(0xffffffffu, busy)) { if(exhausted) break; continue; }

        if(busy) {
            // ---- trace the lane's ray to completion
            Hit hit;
            Trav T;
            const WorldRayRegs wr{{ro.x, ro.y, ro.z}, {rd.x, rd.y, rd.z}, rtmax};
            travInit<true>(P, T, hit, ro.x, ro.y, ro.z, rd.x, rd.y, rd.z, rtmax);
for(uint32_t steps = 0; !travStep<COUNT, false>(P, T, stack, hit, wr, rtmin, cntT) && steps < kMaxStepsPerRay; ++steps) {}
            // ---- shade: hit / miss program, then the frames that resume, up to the next traceRayEXT
            const int outcome = shadeContext<COUNT, MULTI, SEQ>(P, K, hit, ro, rd, rtmin, rtmax, hv, depth, curIOR, refDepth, rayType, missIndex, rayKind, recDepth, sp,
                                                                FramesLocal{frames}, pix, &s_cnt[CNT_SKY][tid], cntT);
            if(outcome == 1) { s_cnt[rayKind][tid]++; pix.rays++; }
            else busy = false;
        }
    }

#ifdef RG_DEBUG_TAIL
    if(lane == 0) {
        unsigned long long tWarp1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tWarp1));
        atomicMax(P.counters + 10, tWarp1); atomicAdd(P.counters + 11, tWarp1 - tWarp0); atomicMax(P.counters + 12, ~tWarp0); atomicAdd(P.counters + 13, 1ull);
    }
#endif
    // ---- ray counters: warp reduce, one atomic per warp and counter
#pragma unroll
    for(int k = 0; k < CNT_N; ++k) {
        if(!COUNT && k >= CNT_NODES) break;
        unsigned long long v = k < CNT_NODES ? (unsigned long long)s_cnt[k][tid] : (unsigned long long)cntT[k];
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
    Trav T;
    uint2 stack[kStackSize];
    uint32_t cnt[CNT_N];
    const WorldRayRegs wr{{q[0], q[1], q[2]}, {q[3], q[4], q[5]}, q[7]};
    travInit<true>(P, T, hit, q[0], q[1], q[2], q[3], q[4], q[5], q[7]);
    for(uint32_t steps = 0; !travStep<false, true>(P, T, StackLocal{stack}, hit, wr, q[6], cnt) && steps < kMaxStepsPerRay; ++steps) {}
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
// Counting sort of the tile slots into 8 cost classes (relative to the mean cost = rays of the frame / slots, from the trace kernel's own
// ray counters), most expensive class first: a histogram pass and a scatter pass over all SMs.  work[0..7] = class counts, work[8..15] =
// scatter cursors (both zero on entry; the scatter pass leaves them zero again).
__device__ __forceinline__ int costClass(uint32_t c, float mean) {   // 7 = >= 8x mean ... 0 = < mean / 8
    const float r = (float)c / mean;
    const int k = 3 + (int)floorf(log2f(fmaxf(r, 1e-6f)));
    return k < 0 ? 0 : (k > 7 ? 7 : k);
}
__device__ __forceinline__ float meanCost(const unsigned long long* counters, uint32_t n) {
    const unsigned long long rays = counters[CNT_PRIMARY] + counters[CNT_SHADOW] + counters[CNT_REFLECT] + counters[CNT_REFRACT];
    return fmaxf((float)rays / (float)(n ? n : 1u), 1.0f);
}
__global__ void __launch_bounds__(256) k_order_hist(const uint32_t* __restrict__ cost, uint32_t n, const unsigned long long* __restrict__ counters, uint32_t* __restrict__ work) {
    __shared__ uint32_t sCount[8];
    if(threadIdx.x < 8) sCount[threadIdx.x] = 0u;
    __syncthreads();
    const float mean = meanCost(counters, n);
    for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) atomicAdd(&sCount[costClass(cost[i], mean)], 1u);
    __syncthreads();
    if(threadIdx.x < 8 && sCount[threadIdx.x]) atomicAdd(work + threadIdx.x, sCount[threadIdx.x]);
}
__global__ void __launch_bounds__(256) k_order_scatter(uint32_t* __restrict__ cost, uint32_t n, const unsigned long long* __restrict__ counters, uint32_t* __restrict__ work,
                                                       uint32_t* __restrict__ order, uint32_t* __restrict__ done) {
    __shared__ uint32_t sBase[8];
    if(threadIdx.x == 0) { uint32_t run = 0; for(int k = 7; k >= 0; --k) { sBase[k] = run; run += work[k]; } }
    __syncthreads();
    const float mean = meanCost(counters, n);
    const uint32_t lane = threadIdx.x & 31u;
    for(uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; i0 < n; i0 += gridDim.x * blockDim.x) {   // warp-uniform trip count
        const uint32_t i = i0 + lane;
        const int cls = i < n ? costClass(cost[i], mean) : -1;
#pragma unroll
        for(int k = 0; k < 8; ++k) {   // one atomic per warp and class
            const uint32_t m = __ballot_sync(0xffffffffu, cls == k);
            if(!m) continue;
            uint32_t at = 0;
            if(lane == (uint32_t)(__ffs(m) - 1)) at = atomicAdd(work + 8 + k, (uint32_t)__popc(m));
            at = __shfl_sync(0xffffffffu, at, __ffs(m) - 1);
            if(cls == k) order[sBase[k] + at + __popc(m & ((1u << lane) - 1u))] = i;
        }
        if(i < n) cost[i] = 0u;
    }
    // the last block to finish clears the histogram and the cursors for the next frame
    __syncthreads();
    if(threadIdx.x == 0) {
        __threadfence();
        if(atomicAdd(done, 1u) == gridDim.x - 1u) { for(int k = 0; k < 16; ++k) work[k] = 0u; *done = 0u; }
    }
}
}  // namespace

void launchOrderTiles(uint32_t* cost, uint32_t nSlots, uint32_t* order, const unsigned long long* counters, uint32_t* work, int numSms, cudaStream_t stream) {
    if(!nSlots) return;
    const int g = (int)((nSlots + 255u) / 256u) < numSms * 2 ? (int)((nSlots + 255u) / 256u) : numSms * 2;
    k_order_hist<<<g, 256, 0, stream>>>(cost, nSlots, counters, work);
    k_order_scatter<<<g, 256, 0, stream>>>(cost, nSlots, counters, work, order, work + 16);
}

template <bool COUNT, bool MULTI, bool POOL, bool SEQ>
static int tracePerSm() {   // persistent grid: a multiple of the SM count; resident CTAs per SM limited by registers / shared memory
    static int perSm = [] {
        int v = 0;
        if(POOL) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k_trace_pool<COUNT, MULTI>, 128, 0);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k_trace_lanes<COUNT, MULTI, SEQ>, 128, 0);
        if(const char* e = getenv(POOL ? "RGB200_POOL_CTAS" : "RGB200_LANES_CTAS")) { const int w = atoi(e); if(w >= 1 && w < v) v = w; }   // developer knob: occupancy sweeps
        if(const char* e = getenv(POOL ? "RGB200_POOL_CARVEOUT" : "RGB200_LANES_CARVEOUT")) {   // developer knob: shared-memory carve-out in percent (the rest is L1)
            if(POOL) cudaFuncSetAttribute(k_trace_pool<COUNT, MULTI>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e));
            else cudaFuncSetAttribute(k_trace_lanes<COUNT, MULTI, SEQ>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e));
        }
        return v < 1 ? 1 : v;
    }();
    return perSm;
}

template <bool COUNT, bool MULTI>
static void launchTraceK(const TraceParams& p, int numSms, bool pool, bool seq, cudaStream_t stream) {
    if(pool) k_trace_pool<COUNT, MULTI><<<numSms * tracePerSm<COUNT, MULTI, true, false>(), 128, 0, stream>>>(p);
    else if(seq) k_trace_lanes<COUNT, MULTI, true><<<numSms * tracePerSm<COUNT, MULTI, false, true>(), 128, 0, stream>>>(p);
    else k_trace_lanes<COUNT, MULTI, false><<<numSms * tracePerSm<COUNT, MULTI, false, false>(), 128, 0, stream>>>(p);
}

size_t tracePoolBytes(int numSms) {   // frames of every context of every warp of the largest persistent pool grid
    int perSm = tracePerSm<false, false, true, false>();
    if(tracePerSm<false, true, true, false>() > perSm) perSm = tracePerSm<false, true, true, false>();
    if(tracePerSm<true, true, true, false>() > perSm) perSm = tracePerSm<true, true, true, false>();
    return sizeof(float4) * (size_t)numSms * perSm * 4u * kPoolCtx * kCtxQuads;
}

int traceLanesWarps(int numSms) { return numSms * tracePerSm<false, false, false, true>() * 4; }

void launchTrace(const TraceParams& p, int numSms, bool pool, bool seq, cudaStream_t stream) {
    if(p.flags & RG_COUNT_TRAVERSAL) { launchTraceK<true, true>(p, numSms, pool, seq, stream); return; }
    if(p.nTargets > 1) launchTraceK<false, true>(p, numSms, pool, seq, stream); else launchTraceK<false, false>(p, numSms, pool, seq, stream);
}

void launchTraceRays(const TraceParams& p, const float* rays8, uint32_t n, float* tuv, uint32_t* instPrim, cudaStream_t stream) {
    if(n == 0) return;
    k_trace_rays<<<(n + 127) / 128, 128, 0, stream>>>(p, rays8, n, tuv, instPrim);
}

}  // namespace rg
