// rg_post.cu -- the compute post passes as image kernels for sm_100a.
//
// Replaces the five Vulkan compute pipelines dispatched by raygun/render/raytracer.cpp:106-144 and the blit of
// raygun/render/render_system.cpp:130-144:
//   rough_prepare  resources/shaders/rough_prepare.comp:28-54
//   rough_blur_h/v resources/shaders/rough_blur.h:23-40 (10 x H then V)
//   postprocess    resources/shaders/postprocess.comp:28-50
//   fxaa           resources/shaders/fxaa.comp:29-36 + fxaa.h:617-1134 (preset 39, green as luma)
//   blit           rgba16f -> 8-bit, nearest
// Storage semantics are kept: every image store rounds to binary16, the transition image is R8_SNORM, an
// out-of-bounds load is 0 and an out-of-bounds store is dropped.  All images are pitch = width of the rectangle
// this context renders (region + halo); at a true frame border the rectangle border coincides with it.
#include "rg_post.cuh"

#include <cuda_fp16.h>

#include "../../include/rgb200.h"

namespace rg {

namespace {

struct F4 { float x, y, z, w; };
__device__ __forceinline__ F4 unpackHalf4(uint2 v) {
    const __half2 a = *reinterpret_cast<const __half2*>(&v.x), b = *reinterpret_cast<const __half2*>(&v.y);
    const float2 fa = __half22float2(a), fb = __half22float2(b);
    return F4{fa.x, fa.y, fb.x, fb.y};
}
__device__ __forceinline__ uint32_t f2h(float f) { return (uint32_t)__half_as_ushort(__float2half_rn(f)); }
__device__ __forceinline__ uint2 packHalf4(float x, float y, float z, float w) { return make_uint2(f2h(x) | (f2h(y) << 16), f2h(z) | (f2h(w) << 16)); }

__device__ __forceinline__ F4 loadOob0(const uint2* __restrict__ img, int x, int y, int W, int H) {
    if(x < 0 || y < 0 || x >= W || y >= H) return F4{0, 0, 0, 0};
    return unpackHalf4(__ldg(img + (size_t)y * W + x));
}
__device__ __forceinline__ F4 texel(const uint2* __restrict__ img, int x, int y, int W, int H) {
    x = min(max(x, 0), W - 1); y = min(max(y, 0), H - 1);
    return unpackHalf4(__ldg(img + (size_t)y * W + x));
}
__device__ __forceinline__ float texelG(const uint2* __restrict__ img, int x, int y, int W, int H) {
    x = min(max(x, 0), W - 1); y = min(max(y, 0), H - 1);
    const uint32_t lo = __ldg(reinterpret_cast<const uint32_t*>(img + (size_t)y * W + x));
    return __half2float(__ushort_as_half((unsigned short)(lo >> 16)));
}
__device__ __forceinline__ float dist4(F4 a, F4 b) {
    const float dx = b.x - a.x, dy = b.y - a.y, dz = b.z - a.z, dw = b.w - a.w;
    return sqrtf((dx * dx + dy * dy) + (dz * dz + dw * dw));
}
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
__device__ __forceinline__ float fromSnorm8(signed char v) { return fmaxf((float)v / 127.0f, -1.0f); }

// rough_prepare.comp:28-54.  Besides the three images the shader writes, the kernel appends every pixel the blur passes will touch
// (transition weight >= 0.001 after the SNORM8 store, i.e. a stored byte > 0) to `list` (warp vote + one atomic per warp): the 20 blur
// sub-passes then run over that list -- on the example scene 11 % of the pixels -- instead of scanning the whole transition image 20 times.
// listCount[parity] is this frame's counter; the other one is cleared for the next frame.
__global__ void k_rough_prepare(const uint2* __restrict__ rough, const uint2* __restrict__ normal, int W, int H, signed char* __restrict__ trans,
                                uint2* __restrict__ roughA, uint2* __restrict__ roughB, uint32_t* __restrict__ list, uint32_t* __restrict__ listCount, int parity) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if(x == 0 && y == 0) listCount[parity ^ 1] = 0u;
    const bool inside = x < W && y < H;
    bool active = false;
    const size_t i = (size_t)y * W + x;
    if(inside) {
        const uint2 rraw = __ldg(rough + i);
        const F4 r = unpackHalf4(rraw);
        const F4 ru = loadOob0(rough, x, y + 1, W, H), rd = loadOob0(rough, x, y - 1, W, H), rl = loadOob0(rough, x + 1, y, W, H), rr = loadOob0(rough, x - 1, y, W, H);
        const F4 n = unpackHalf4(__ldg(normal + i));
        const F4 nu = loadOob0(normal, x, y + 1, W, H), nd = loadOob0(normal, x, y - 1, W, H), nl = loadOob0(normal, x + 1, y, W, H), nr = loadOob0(normal, x - 1, y, W, H);
        const float uf = fminf(ru.w, r.w) * clampf(1.0f - dist4(nu, n) * 10.0f, 0.0f, 1.0f);
        const float df = fminf(rd.w, r.w) * clampf(1.0f - dist4(nd, n) * 10.0f, 0.0f, 1.0f);
        const float lf = fminf(rl.w, r.w) * clampf(1.0f - dist4(nl, n) * 10.0f, 0.0f, 1.0f);
        const float rf = fminf(rr.w, r.w) * clampf(1.0f - dist4(nr, n) * 10.0f, 0.0f, 1.0f);
        float t = fminf(fminf(fminf(uf, df), lf), rf);
        t = (t == t) ? clampf(t, -1.0f, 1.0f) : 0.0f;
        const signed char t8 = (signed char)__float2int_rn(t * 127.0f);
        trans[i] = t8;
        roughA[i] = rraw;
        roughB[i] = rraw;
        active = t8 > 0;   // fromSnorm8(t8) >= 0.001  <=>  t8 >= 1
    }
    const uint32_t m = __ballot_sync(0xffffffffu, active), lane = threadIdx.x & 31u;   // blockDim.x == 32: a warp is one row segment
    if(m) {
        uint32_t base = 0;
        if(lane == 0) base = atomicAdd(listCount + parity, (uint32_t)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if(active) list[base + __popc(m & ((1u << lane) - 1u))] = (uint32_t)i;
    }
}

// One blur sub-pass of rough_blur.h:23-40 at pixel (x, y) of `in`, as the value the shader STORES (binary16); (ox, oy) = (1,0) for
// the H pass, (0,1) for the V pass.  A pixel whose transition weight is below 0.001 is not written by the shader, and every rough
// image holds the rough_prepare value there, so `in` itself is what a later pass reads.
__device__ __forceinline__ uint2 blurPixel(const uint2* __restrict__ in, const signed char* __restrict__ trans, int x, int y, int W, int H, int ox, int oy) {
    const size_t i = (size_t)y * W + x;
    const uint2 raw = __ldg(in + i);
    const float t = fromSnorm8(trans[i]);
    if(t < 0.001f) return raw;
    const F4 r = unpackHalf4(raw);
    const F4 r1 = loadOob0(in, x + ox, y + oy, W, H), r2 = loadOob0(in, x - ox, y - oy, W, H);
    const float s = 1.0f - t - t;
    return packHalf4(r.x * s + r1.x * t + r2.x * t, r.y * s + r1.y * t + r2.y * t, r.z * s + r1.z * t + r2.z * t, r.w);
}

// Blur over the list of active pixels.  FUSED: one H pass and the V pass that follows it in ONE launch: out = V(H(in)), where the three
// H values a V pixel reads (rows y - 1, y, y + 1) are recomputed on the fly and rounded to binary16 as the H pass would have stored
// them -- bit-identical to the two launches, half the launches and no intermediate image.  !FUSED: a single sub-pass (ox, oy).
template <bool FUSED>
__global__ void __launch_bounds__(256) k_rough_blur_list(const uint2* __restrict__ in, uint2* __restrict__ out, const signed char* __restrict__ trans,
                                                         const uint32_t* __restrict__ list, const uint32_t* __restrict__ count, int W, int H, int ox, int oy) {
    const uint32_t n = *count;
    for(uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const uint32_t i = list[k];
        const int y = (int)(i / (uint32_t)W), x = (int)(i - (uint32_t)y * (uint32_t)W);
        if(!FUSED) { out[i] = blurPixel(in, trans, x, y, W, H, ox, oy); continue; }
        const float t = fromSnorm8(trans[i]);   // >= 0.001: the pixel is on the list
        const F4 r = unpackHalf4(blurPixel(in, trans, x, y, W, H, 1, 0));
        const F4 r1 = y + 1 < H ? unpackHalf4(blurPixel(in, trans, x, y + 1, W, H, 1, 0)) : F4{0, 0, 0, 0};
        const F4 r2 = y > 0 ? unpackHalf4(blurPixel(in, trans, x, y - 1, W, H, 1, 0)) : F4{0, 0, 0, 0};
        const float s = 1.0f - t - t;
        out[i] = packHalf4(r.x * s + r1.x * t + r2.x * t, r.y * s + r1.y * t + r2.y * t, r.z * s + r1.z * t + r2.z * t, r.w);
    }
}

// postprocess.comp:28-50
__global__ void k_postprocess(uint2* base, const uint2* roughA_, uint2* final_, uint2* normal, uint2* rough, signed char* trans, uint2* roughA,
                              uint2* roughB, int W, int H, float4 fade, int showAlpha) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if(x >= W || y >= H) return;
    const size_t i = (size_t)y * W + x;
    const F4 b = unpackHalf4(base[i]);
    const F4 r = unpackHalf4(roughA_[i]);
    float cx = clampf(b.x * (1.0f - b.w) + r.x * b.w, 0.0f, 1.0f), cy = clampf(b.y * (1.0f - b.w) + r.y * b.w, 0.0f, 1.0f),
          cz = clampf(b.z * (1.0f - b.w) + r.z * b.w, 0.0f, 1.0f);
    // GLM-style clamp keeps NaN; GLSL leaves it undefined.  fminf/fmaxf drop NaN to the bound: 0.
    const float luma = cx * 0.299f + cy * 0.587f + cz * 0.114f;
    cx = cx * (1.0f - fade.w) + fade.x * fade.w; cy = cy * (1.0f - fade.w) + fade.y * fade.w; cz = cz * (1.0f - fade.w) + fade.z * fade.w;
    uint2 o = packHalf4(cx, cy, cz, luma);
    if(showAlpha) {  // debug path, postprocess.comp:41-49
        auto aaaa = [](uint2 v) { const uint32_t a = v.y >> 16; return make_uint2(a | (a << 16), a | (a << 16)); };
        o = aaaa(o);
        normal[i] = aaaa(normal[i]); rough[i] = aaaa(rough[i]); base[i] = aaaa(base[i]);
        trans[i] = 127;  // single-channel image: .a reads 1.0
        roughA[i] = aaaa(roughA[i]); roughB[i] = aaaa(roughB[i]);
    }
    final_[i] = o;
}

// fxaa.h:617-1134 for FXAA_PC, FXAA_GLSL_130, preset 39, FXAA_GREEN_AS_LUMA; fxaa.comp:31-35 constants.
// (gx, gy) are coordinates inside the local rectangle `tex` (LW x LH); the rcpFrame / texCoord arithmetic uses the
// FULL frame size (fw, fh) and the global pixel position so the result is bit-identical to a single-GPU frame.
__device__ F4 fxaaPixel(const uint2* __restrict__ tex, int gx, int gy, int LW, int LH, int ox, int oy, int fw, int fh) {
    const float rcpX = 1.0f / (float)fw, rcpY = 1.0f / (float)fh;
    const float P[12] = {1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.5f, 2.0f, 2.0f, 2.0f, 2.0f, 4.0f, 8.0f};
    // global texture coordinate of the pixel centre; local sampling subtracts the rectangle origin in texel space
    float posMx = ((float)(gx + ox) + 0.5f) / (float)fw, posMy = ((float)(gy + oy) + 0.5f) / (float)fh;
    // textureLod / textureLodOffset(tex, p, 0, o): linear filter with float weights, clamp-to-edge of the FULL frame, texel offset
    // added after the floor; then into the local rectangle's coordinates
    auto sG = [&](float px, float py, int dx, int dy) {
        const float u = px * (float)fw - 0.5f, v = py * (float)fh - 0.5f;
        const float fu = floorf(u), fv = floorf(v);
        const float ax = u - fu, ay = v - fv;
        const int X0 = min(max((int)fu + dx, 0), fw - 1) - ox, X1 = min(max((int)fu + dx + 1, 0), fw - 1) - ox;
        const int Y0 = min(max((int)fv + dy, 0), fh - 1) - oy, Y1 = min(max((int)fv + dy + 1, 0), fh - 1) - oy;
        const float c00 = texelG(tex, X0, Y0, LW, LH), c10 = texelG(tex, X1, Y0, LW, LH), c01 = texelG(tex, X0, Y1, LW, LH), c11 = texelG(tex, X1, Y1, LW, LH);
        const float top = c00 * (1.0f - ax) + c10 * ax, bot = c01 * (1.0f - ax) + c11 * ax;
        return top * (1.0f - ay) + bot * ay;
    };
    auto s4 = [&](float px, float py) {
        const float u = px * (float)fw - 0.5f, v = py * (float)fh - 0.5f;
        const float fu = floorf(u), fv = floorf(v);
        const float ax = u - fu, ay = v - fv;
        const int X0 = min(max((int)fu, 0), fw - 1) - ox, X1 = min(max((int)fu + 1, 0), fw - 1) - ox;
        const int Y0 = min(max((int)fv, 0), fh - 1) - oy, Y1 = min(max((int)fv + 1, 0), fh - 1) - oy;
        const F4 c00 = texel(tex, X0, Y0, LW, LH), c10 = texel(tex, X1, Y0, LW, LH), c01 = texel(tex, X0, Y1, LW, LH), c11 = texel(tex, X1, Y1, LW, LH);
        F4 o;
        o.x = (c00.x * (1.0f - ax) + c10.x * ax) * (1.0f - ay) + (c01.x * (1.0f - ax) + c11.x * ax) * ay;
        o.y = (c00.y * (1.0f - ax) + c10.y * ax) * (1.0f - ay) + (c01.y * (1.0f - ax) + c11.y * ax) * ay;
        o.z = (c00.z * (1.0f - ax) + c10.z * ax) * (1.0f - ay) + (c01.z * (1.0f - ax) + c11.z * ax) * ay;
        o.w = (c00.w * (1.0f - ax) + c10.w * ax) * (1.0f - ay) + (c01.w * (1.0f - ax) + c11.w * ax) * ay;
        return o;
    };

    // fxaa.h:804-814 (non-gather path): the taps at the pixel centre are sampler reads too
    const F4 rgbyM = s4(posMx, posMy);
    const float lumaM = rgbyM.y;
    float lumaS = sG(posMx, posMy, 0, 1), lumaE = sG(posMx, posMy, 1, 0), lumaN = sG(posMx, posMy, 0, -1), lumaW = sG(posMx, posMy, -1, 0);
    const float maxSM = fmaxf(lumaS, lumaM), minSM = fminf(lumaS, lumaM);
    const float maxESM = fmaxf(lumaE, maxSM), minESM = fminf(lumaE, minSM);
    const float maxWN = fmaxf(lumaN, lumaW), minWN = fminf(lumaN, lumaW);
    const float rangeMax = fmaxf(maxWN, maxESM), rangeMin = fminf(minWN, minESM);
    const float rangeMaxScaled = rangeMax * 0.063f;
    const float range = rangeMax - rangeMin;
    const float rangeMaxClamped = fmaxf(0.0312f, rangeMaxScaled);
    if(range < rangeMaxClamped) return rgbyM;

    const float lumaNW = sG(posMx, posMy, -1, -1), lumaSE = sG(posMx, posMy, 1, 1), lumaNE = sG(posMx, posMy, 1, -1), lumaSW = sG(posMx, posMy, -1, 1);
    const float lumaNS = lumaN + lumaS, lumaWE = lumaW + lumaE;
    const float subpixRcpRange = 1.0f / range;
    const float subpixNSWE = lumaNS + lumaWE;
    const float edgeHorz1 = (-2.0f * lumaM) + lumaNS, edgeVert1 = (-2.0f * lumaM) + lumaWE;
    const float lumaNESE = lumaNE + lumaSE, lumaNWNE = lumaNW + lumaNE;
    const float edgeHorz2 = (-2.0f * lumaE) + lumaNESE, edgeVert2 = (-2.0f * lumaN) + lumaNWNE;
    const float lumaNWSW = lumaNW + lumaSW, lumaSWSE = lumaSW + lumaSE;
    const float edgeHorz4 = (fabsf(edgeHorz1) * 2.0f) + fabsf(edgeHorz2), edgeVert4 = (fabsf(edgeVert1) * 2.0f) + fabsf(edgeVert2);
    const float edgeHorz3 = (-2.0f * lumaW) + lumaNWSW, edgeVert3 = (-2.0f * lumaS) + lumaSWSE;
    const float edgeHorz = fabsf(edgeHorz3) + edgeHorz4, edgeVert = fabsf(edgeVert3) + edgeVert4;
    const float subpixNWSWNESE = lumaNWSW + lumaNESE;
    float lengthSign = rcpX;
    const bool horzSpan = edgeHorz >= edgeVert;
    const float subpixA = subpixNSWE * 2.0f + subpixNWSWNESE;
    if(!horzSpan) lumaN = lumaW;
    if(!horzSpan) lumaS = lumaE;
    if(horzSpan) lengthSign = rcpY;
    const float subpixB = (subpixA * (1.0f / 12.0f)) - lumaM;
    const float gradientN = lumaN - lumaM, gradientS = lumaS - lumaM;
    float lumaNN = lumaN + lumaM;
    const float lumaSS = lumaS + lumaM;
    const bool pairN = fabsf(gradientN) >= fabsf(gradientS);
    const float gradient = fmaxf(fabsf(gradientN), fabsf(gradientS));
    if(pairN) lengthSign = -lengthSign;
    const float subpixC = clampf(fabsf(subpixB) * subpixRcpRange, 0.0f, 1.0f);
    float posBx = posMx, posBy = posMy;
    const float offNPx = (!horzSpan) ? 0.0f : rcpX, offNPy = (horzSpan) ? 0.0f : rcpY;
    if(!horzSpan) posBx += lengthSign * 0.5f;
    if(horzSpan) posBy += lengthSign * 0.5f;
    float posNx = posBx - offNPx * P[0], posNy = posBy - offNPy * P[0];
    float posPx = posBx + offNPx * P[0], posPy = posBy + offNPy * P[0];
    const float subpixD = ((-2.0f) * subpixC) + 3.0f;
    float lumaEndN = sG(posNx, posNy, 0, 0);
    const float subpixE = subpixC * subpixC;
    float lumaEndP = sG(posPx, posPy, 0, 0);
    if(!pairN) lumaNN = lumaSS;
    const float gradientScaled = gradient * 1.0f / 4.0f;
    const float lumaMM = lumaM - lumaNN * 0.5f;
    const float subpixF = subpixD * subpixE;
    const bool lumaMLTZero = lumaMM < 0.0f;
    lumaEndN -= lumaNN * 0.5f;
    lumaEndP -= lumaNN * 0.5f;
    bool doneN = fabsf(lumaEndN) >= gradientScaled, doneP = fabsf(lumaEndP) >= gradientScaled;
    if(!doneN) { posNx -= offNPx * P[1]; posNy -= offNPy * P[1]; }
    bool doneNP = (!doneN) || (!doneP);
    if(!doneP) { posPx += offNPx * P[1]; posPy += offNPy * P[1]; }
#pragma unroll 1
    for(int k = 2; k < 12 && doneNP; ++k) {
        if(!doneN) lumaEndN = sG(posNx, posNy, 0, 0);
        if(!doneP) lumaEndP = sG(posPx, posPy, 0, 0);
        if(!doneN) lumaEndN = lumaEndN - lumaNN * 0.5f;
        if(!doneP) lumaEndP = lumaEndP - lumaNN * 0.5f;
        doneN = fabsf(lumaEndN) >= gradientScaled;
        doneP = fabsf(lumaEndP) >= gradientScaled;
        if(!doneN) { posNx -= offNPx * P[k]; posNy -= offNPy * P[k]; }
        doneNP = (!doneN) || (!doneP);
        if(!doneP) { posPx += offNPx * P[k]; posPy += offNPy * P[k]; }
    }
    float dstN = posMx - posNx, dstP = posPx - posMx;
    if(!horzSpan) dstN = posMy - posNy;
    if(!horzSpan) dstP = posPy - posMy;
    const bool goodSpanN = (lumaEndN < 0.0f) != lumaMLTZero;
    const float spanLength = (dstP + dstN);
    const bool goodSpanP = (lumaEndP < 0.0f) != lumaMLTZero;
    const float spanLengthRcp = 1.0f / spanLength;
    const bool directionN = dstN < dstP;
    const float dst = fminf(dstN, dstP);
    const bool goodSpan = directionN ? goodSpanN : goodSpanP;
    const float subpixG = subpixF * subpixF;
    const float pixelOffset = (dst * (-spanLengthRcp)) + 0.5f;
    const float subpixH = subpixG * 1.0f;
    const float pixelOffsetGood = goodSpan ? pixelOffset : 0.0f;
    const float pixelOffsetSubpix = fmaxf(pixelOffsetGood, subpixH);
    if(!horzSpan) posMx += pixelOffsetSubpix * lengthSign;
    if(horzSpan) posMy += pixelOffsetSubpix * lengthSign;
    F4 s = s4(posMx, posMy);
    s.w = lumaM;
    return s;
}

__device__ __forceinline__ float srgbOetf(float x) { return x <= 0.0031308f ? x * 12.92f : 1.055f * powf(x, 1.0f / 2.4f) - 0.055f; }
__device__ __forceinline__ uint32_t toUnorm8(float x) {
    x = (x == x) ? clampf(x, 0.0f, 1.0f) : 0.0f;
    return (uint32_t)(x * 255.0f + 0.5f);
}
__device__ __forceinline__ uint32_t toRgba8(F4 c, bool srgb) {
    if(srgb) { c.x = srgbOetf(clampf(c.x, 0.0f, 1.0f)); c.y = srgbOetf(clampf(c.y, 0.0f, 1.0f)); c.z = srgbOetf(clampf(c.z, 0.0f, 1.0f)); }
    return toUnorm8(c.x) | (toUnorm8(c.y) << 8) | (toUnorm8(c.z) << 16) | (toUnorm8(c.w) << 24);
}

// FXAA (optional) + 8-bit convert + the tile gather: the interior of this context's region is also stored straight
// into the gather target (a full-frame RGBA8 image that may live in a peer GPU's memory, reached over NVLink).
__global__ void k_fxaa_blit(const PostParams p) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if(x >= p.rw || y >= p.rh) return;
    const size_t i = (size_t)y * p.rw + x;
    F4 c;
    if(p.flags & RG_FXAA) {
        c = fxaaPixel(p.final_, x, y, p.rw, p.rh, p.rx0, p.ry0, p.W, p.H);
        p.fxaaOut[i] = packHalf4(c.x, c.y, c.z, c.w);
        // what the blit reads is the stored (binary16) value
        c = unpackHalf4(packHalf4(c.x, c.y, c.z, c.w));
    } else {
        c = unpackHalf4(p.final_[i]);
    }
    const uint32_t px = toRgba8(c, (p.flags & RG_SRGB8) != 0);
    const int gx = x + p.rx0, gy = y + p.ry0;
    if(gx >= p.ix0 && gx < p.ix1 && gy >= p.iy0 && gy < p.iy1) {
        p.rgba8[(size_t)(gy - p.iy0) * (p.ix1 - p.ix0) + (gx - p.ix0)] = px;
        if(p.gather) p.gather[(size_t)gy * p.W + gx] = px;
    }
}

inline dim3 grid2d(int w, int h, dim3 b) { return dim3((w + b.x - 1) / b.x, (h + b.y - 1) / b.y); }

}  // namespace

void launchRoughPrepare(const PostParams& p, cudaStream_t st) {
    const dim3 b(32, 8);
    k_rough_prepare<<<grid2d(p.rw, p.rh, b), b, 0, st>>>(p.rough, p.normal, p.rw, p.rh, p.trans, p.roughA, p.roughB, p.blurList, p.blurCount, p.blurParity);
}
// The reference's 10 x (H: A -> B, V: B -> A) as 9 fused launches (B -> A, A -> B, ..., B -> A: both images hold the rough_prepare
// value wherever a pass does not write) + the last H and V on their own, so that roughA (the result) AND roughB (the last H pass)
// end up exactly as the 20 dispatches leave them.  Grid: one wave of the SMs, grid-stride over the list.
int launchRoughBlur(const PostParams& p, int numSms, cudaStream_t st) {
    const uint32_t* cnt = p.blurCount + p.blurParity;
    const int g = numSms * 4;
    const uint2 *in = p.roughB; uint2* out = p.roughA;
    for(int i = 0; i < 9; ++i) {
        k_rough_blur_list<true><<<g, 256, 0, st>>>(in, out, p.trans, p.blurList, cnt, p.rw, p.rh, 0, 0);
        const uint2* t = in; in = out; out = const_cast<uint2*>(t);
    }
    k_rough_blur_list<false><<<g, 256, 0, st>>>(p.roughA, p.roughB, p.trans, p.blurList, cnt, p.rw, p.rh, 1, 0);
    k_rough_blur_list<false><<<g, 256, 0, st>>>(p.roughB, p.roughA, p.trans, p.blurList, cnt, p.rw, p.rh, 0, 1);
    return 11;
}
void launchPostprocess(const PostParams& p, cudaStream_t st) {
    const dim3 b(32, 8);
    k_postprocess<<<grid2d(p.rw, p.rh, b), b, 0, st>>>(p.base, p.roughA, p.final_, p.normal, p.rough, p.trans, p.roughA, p.roughB, p.rw, p.rh, p.fade, p.showAlpha);
}
void launchFxaaBlit(const PostParams& p, cudaStream_t st) {
    const dim3 b(32, 8);
    k_fxaa_blit<<<grid2d(p.rw, p.rh, b), b, 0, st>>>(p);
}

}  // namespace rg
