// TEST INFRASTRUCTURE -- CPU oracle (see orc_math.h header).
//
// orc_post.cpp: restatement of the five compute post passes and the 8-bit blit, in the
// order and with the barriers of raygun/render/raytracer.cpp:106-144:
//   rough_prepare  resources/shaders/rough_prepare.comp:28-54
//   rough_blur_h/v resources/shaders/rough_blur.h:23-40, rough_blur_h.comp:28-32, rough_blur_v.comp:28-32
//   postprocess    resources/shaders/postprocess.comp:28-50
//   fxaa           resources/shaders/fxaa.comp:29-36, fxaa.h:617-1134 (preset 39, green-as-luma)
//   blit           raygun/render/render_system.cpp:130-144
// Storage semantics: every imageStore rounds to binary16; the
// transition image is R8_SNORM (raytracer.cpp:187); out-of-bounds imageLoad returns 0 and
// out-of-bounds imageStore is dropped, so the 16x16 work-group padding (raytracer.cpp:108-109)
// reduces to "loop over in-bounds pixels".  The FXAA sampler is linear / clamp-to-edge
// (raygun/compute/compute_system.cpp:76-84) evaluated with float weights; the nine taps at the pixel centre
// (textureLodOffset at posM, fxaa.h:804-845, non-gather path) go through the same bilinear arithmetic.
#include <omp.h>

#include <cmath>
#include <vector>

#include "orc_render.h"

namespace orc {

namespace {

struct Img {
    const half4* p; int W, H;
    vec4 load(int x, int y) const { return (x < 0 || y < 0 || x >= W || y >= H) ? vec4(0.0f) : unpack_half4(p[(size_t)y * W + x]); }
    vec4 texel(int x, int y) const {  // clamp-to-edge
        x = x < 0 ? 0 : (x >= W ? W - 1 : x); y = y < 0 ? 0 : (y >= H ? H - 1 : y);
        return unpack_half4(p[(size_t)y * W + x]);
    }
    // textureLod / textureLodOffset(sampler2D, p, 0, o) with linear filter and clamp-to-edge: float weights, the texel
    // offset is added after the floor (as llvmpipe's float path does for rgba16f)
    vec4 sample(float px, float py, int ox = 0, int oy = 0) const {
        const float u = px * (float)W - 0.5f, v = py * (float)H - 0.5f;
        const float fu = std::floor(u), fv = std::floor(v);
        const float ax = u - fu, ay = v - fv;
        const int x0 = (int)fu + ox, y0 = (int)fv + oy;
        const vec4 c00 = texel(x0, y0), c10 = texel(x0 + 1, y0), c01 = texel(x0, y0 + 1), c11 = texel(x0 + 1, y0 + 1);
        const vec4 top = c00 * (1.0f - ax) + c10 * ax, bot = c01 * (1.0f - ax) + c11 * ax;
        return top * (1.0f - ay) + bot * ay;
    }
};

inline int8_t toSnorm8(float x) {  // R8_SNORM store
    if(!(x == x)) return 0;
    x = x < -1.0f ? -1.0f : (x > 1.0f ? 1.0f : x);
    return (int8_t)std::lrintf(x * 127.0f);
}
inline float fromSnorm8(int8_t v) { const float f = (float)v / 127.0f; return f < -1.0f ? -1.0f : f; }
inline int8_t toUnorm8AsByte(float x) {  // alternative reading of the rgba8 declaration (compute.h:39)
    if(!(x == x)) return 0;
    x = x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x);
    return (int8_t)(uint8_t)std::lrintf(x * 255.0f);
}

// fxaa.h:617-1134 with FXAA_PC, FXAA_GLSL_130, FXAA_QUALITY_PRESET 39, FXAA_GREEN_AS_LUMA 1, FXAA_DISCARD 0
vec4 fxaaPixel(const Img& tex, int gx, int gy) {
    const float rcpX = 1.0f / (float)tex.W, rcpY = 1.0f / (float)tex.H;                     // fxaa.comp:31
    float posMx = ((float)gx + 0.5f) / (float)tex.W, posMy = ((float)gy + 0.5f) / (float)tex.H;  // fxaa.comp:32
    const float subpix = 1.0f, edgeThreshold = 0.063f, edgeThresholdMin = 0.0312f;          // fxaa.comp:34
    static const float P[12] = {1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.5f, 2.0f, 2.0f, 2.0f, 2.0f, 4.0f, 8.0f};  // fxaa.h:487-500

    // fxaa.h:804-814 (FXAA_GATHER4_ALPHA == 0): FxaaTexTop / FxaaTexOff are sampler reads AT posM; (gx + 0.5) / W * W - 0.5 is
    // not always exactly gx in binary32, so these go through the bilinear arithmetic like every other tap
    const vec4 rgbyM = tex.sample(posMx, posMy);
    const float lumaM = rgbyM.y;
    float lumaS = tex.sample(posMx, posMy, 0, 1).y, lumaE = tex.sample(posMx, posMy, 1, 0).y;
    float lumaN = tex.sample(posMx, posMy, 0, -1).y, lumaW = tex.sample(posMx, posMy, -1, 0).y;

    const float maxSM = fmax_glsl(lumaS, lumaM), minSM = fmin_glsl(lumaS, lumaM);
    const float maxESM = fmax_glsl(lumaE, maxSM), minESM = fmin_glsl(lumaE, minSM);
    const float maxWN = fmax_glsl(lumaN, lumaW), minWN = fmin_glsl(lumaN, lumaW);
    const float rangeMax = fmax_glsl(maxWN, maxESM), rangeMin = fmin_glsl(minWN, minESM);
    const float rangeMaxScaled = rangeMax * edgeThreshold;
    const float range = rangeMax - rangeMin;
    const float rangeMaxClamped = fmax_glsl(edgeThresholdMin, rangeMaxScaled);
    if(range < rangeMaxClamped) return rgbyM;

    const float lumaNW = tex.sample(posMx, posMy, -1, -1).y, lumaSE = tex.sample(posMx, posMy, 1, 1).y;   // fxaa.h:837-845
    const float lumaNE = tex.sample(posMx, posMy, 1, -1).y, lumaSW = tex.sample(posMx, posMy, -1, 1).y;

    const float lumaNS = lumaN + lumaS, lumaWE = lumaW + lumaE;
    const float subpixRcpRange = 1.0f / range;
    const float subpixNSWE = lumaNS + lumaWE;
    const float edgeHorz1 = (-2.0f * lumaM) + lumaNS, edgeVert1 = (-2.0f * lumaM) + lumaWE;

    const float lumaNESE = lumaNE + lumaSE, lumaNWNE = lumaNW + lumaNE;
    const float edgeHorz2 = (-2.0f * lumaE) + lumaNESE, edgeVert2 = (-2.0f * lumaN) + lumaNWNE;

    const float lumaNWSW = lumaNW + lumaSW, lumaSWSE = lumaSW + lumaSE;
    const float edgeHorz4 = (std::fabs(edgeHorz1) * 2.0f) + std::fabs(edgeHorz2);
    const float edgeVert4 = (std::fabs(edgeVert1) * 2.0f) + std::fabs(edgeVert2);
    const float edgeHorz3 = (-2.0f * lumaW) + lumaNWSW, edgeVert3 = (-2.0f * lumaS) + lumaSWSE;
    const float edgeHorz = std::fabs(edgeHorz3) + edgeHorz4, edgeVert = std::fabs(edgeVert3) + edgeVert4;

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
    const bool pairN = std::fabs(gradientN) >= std::fabs(gradientS);
    const float gradient = fmax_glsl(std::fabs(gradientN), std::fabs(gradientS));
    if(pairN) lengthSign = -lengthSign;
    const float subpixC = clampf(std::fabs(subpixB) * subpixRcpRange, 0.0f, 1.0f);

    float posBx = posMx, posBy = posMy;
    const float offNPx = (!horzSpan) ? 0.0f : rcpX, offNPy = (horzSpan) ? 0.0f : rcpY;
    if(!horzSpan) posBx += lengthSign * 0.5f;
    if(horzSpan) posBy += lengthSign * 0.5f;

    float posNx = posBx - offNPx * P[0], posNy = posBy - offNPy * P[0];
    float posPx = posBx + offNPx * P[0], posPy = posBy + offNPy * P[0];
    const float subpixD = ((-2.0f) * subpixC) + 3.0f;
    float lumaEndN = tex.sample(posNx, posNy).y;
    const float subpixE = subpixC * subpixC;
    float lumaEndP = tex.sample(posPx, posPy).y;

    if(!pairN) lumaNN = lumaSS;
    const float gradientScaled = gradient * 1.0f / 4.0f;
    const float lumaMM = lumaM - lumaNN * 0.5f;
    const float subpixF = subpixD * subpixE;
    const bool lumaMLTZero = lumaMM < 0.0f;

    lumaEndN -= lumaNN * 0.5f;
    lumaEndP -= lumaNN * 0.5f;
    bool doneN = std::fabs(lumaEndN) >= gradientScaled, doneP = std::fabs(lumaEndP) >= gradientScaled;
    if(!doneN) { posNx -= offNPx * P[1]; posNy -= offNPy * P[1]; }
    bool doneNP = (!doneN) || (!doneP);
    if(!doneP) { posPx += offNPx * P[1]; posPy += offNPy * P[1]; }

    for(int k = 2; k < 12 && doneNP; ++k) {  // the nested `if(doneNP)` blocks, fxaa.h:963-1103
        if(!doneN) lumaEndN = tex.sample(posNx, posNy).y;
        if(!doneP) lumaEndP = tex.sample(posPx, posPy).y;
        if(!doneN) lumaEndN = lumaEndN - lumaNN * 0.5f;
        if(!doneP) lumaEndP = lumaEndP - lumaNN * 0.5f;
        doneN = std::fabs(lumaEndN) >= gradientScaled;
        doneP = std::fabs(lumaEndP) >= gradientScaled;
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
    const float dst = fmin_glsl(dstN, dstP);
    const bool goodSpan = directionN ? goodSpanN : goodSpanP;
    const float subpixG = subpixF * subpixF;
    const float pixelOffset = (dst * (-spanLengthRcp)) + 0.5f;
    const float subpixH = subpixG * subpix;

    const float pixelOffsetGood = goodSpan ? pixelOffset : 0.0f;
    const float pixelOffsetSubpix = fmax_glsl(pixelOffsetGood, subpixH);
    if(!horzSpan) posMx += pixelOffsetSubpix * lengthSign;
    if(horzSpan) posMy += pixelOffsetSubpix * lengthSign;
    const vec4 s = tex.sample(posMx, posMy);
    return vec4(s.x, s.y, s.z, lumaM);
}

inline float srgbOetf(float x) { return x <= 0.0031308f ? x * 12.92f : 1.055f * std::pow(x, 1.0f / 2.4f) - 0.055f; }
inline uint8_t toUnorm8(float x) {
    if(!(x == x)) return 0;
    x = x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x);
    return (uint8_t)(x * 255.0f + 0.5f);
}

}  // namespace

void postChain(const Ubo& ubo, Frame& f, uint32_t flags, int threads) {
    const int W = (int)f.W, H = (int)f.H;
    const size_t N = (size_t)W * H;
    if(threads <= 0) threads = omp_get_max_threads();
    const bool unormT = (flags & ORC_TRANSITIONS_UNORM) != 0;
    auto loadT = [&](int x, int y) -> float {
        if(x < 0 || y < 0 || x >= W || y >= H) return 0.0f;
        const int8_t v = f.transitions[(size_t)y * W + x];
        return unormT ? (float)(uint8_t)v / 255.0f : fromSnorm8(v);
    };

    // ---- rough_prepare.comp:28-54
    {
        const Img rough{f.rough, W, H}, normal{f.normal, W, H};
#pragma omp parallel for num_threads(threads) schedule(static)
        for(int y = 0; y < H; ++y)
            for(int x = 0; x < W; ++x) {
                const vec4 r = rough.load(x, y);
                const vec4 ru = rough.load(x, y + 1), rd = rough.load(x, y - 1), rl = rough.load(x + 1, y), rr = rough.load(x - 1, y);
                const vec4 n = normal.load(x, y);
                const vec4 nu = normal.load(x, y + 1), nd = normal.load(x, y - 1), nl = normal.load(x + 1, y), nr = normal.load(x - 1, y);
                const float uf = fmin_glsl(ru.w, r.w) * clampf((1.f - distance(nu, n) * 10.0f), 0.0f, 1.0f);
                const float df = fmin_glsl(rd.w, r.w) * clampf((1.f - distance(nd, n) * 10.0f), 0.0f, 1.0f);
                const float lf = fmin_glsl(rl.w, r.w) * clampf((1.f - distance(nl, n) * 10.0f), 0.0f, 1.0f);
                const float rf = fmin_glsl(rr.w, r.w) * clampf((1.f - distance(nr, n) * 10.0f), 0.0f, 1.0f);
                const float trans = fmin_glsl(fmin_glsl(fmin_glsl(uf, df), lf), rf);
                const size_t i = (size_t)y * W + x;
                f.transitions[i] = unormT ? toUnorm8AsByte(trans) : toSnorm8(trans);
                f.roughA[i] = f.rough[i];
                f.roughB[i] = f.rough[i];
            }
    }

    // ---- 10 x (rough_blur_h: A -> B, offsets (+-1,0); rough_blur_v: B -> A, offsets (0,+-1)), raytracer.cpp:116-121
    auto blur = [&](const half4* inP, half4* outP, int ox, int oy) {
        const Img in{inP, W, H};
#pragma omp parallel for num_threads(threads) schedule(static)
        for(int y = 0; y < H; ++y)
            for(int x = 0; x < W; ++x) {
                const float transition = loadT(x, y);
                if(transition < 0.001f) continue;
                const vec4 r = in.load(x, y);
                const vec4 r1 = in.load(x + ox, y + oy), r2 = in.load(x - ox, y - oy);
                const vec3 col = r.xyz() * (1.f - transition - transition) + r1.xyz() * transition + r2.xyz() * transition;
                outP[(size_t)y * W + x] = pack_half4(vec4(col, r.w));
            }
    };
    for(int i = 0; i < 10; ++i) {
        blur(f.roughA, f.roughB, 1, 0);
        blur(f.roughB, f.roughA, 0, 1);
    }

    // ---- postprocess.comp:28-50
    {
        const vec3 fadeRgb(ubo.fadeColor[0], ubo.fadeColor[1], ubo.fadeColor[2]);
        const float fadeA = ubo.fadeColor[3];
#pragma omp parallel for num_threads(threads) schedule(static)
        for(int y = 0; y < H; ++y)
            for(int x = 0; x < W; ++x) {
                const size_t i = (size_t)y * W + x;
                const vec4 base = unpack_half4(f.base[i]);
                const vec4 rough = unpack_half4(f.roughA[i]);
                vec3 col = clamp3(mix(base.xyz(), rough.xyz(), base.w), 0.0f, 1.0f);
                const float luma = dot(col, vec3(0.299f, 0.587f, 0.114f));
                col = mix(col, fadeRgb, fadeA);
                f.final_[i] = pack_half4(vec4(col, luma));
                if(ubo.showAlpha) {  // debug path, postprocess.comp:41-49
                    auto aaaa = [](half4& h) { h.x = h.y = h.z = h.w; };
                    aaaa(f.final_[i]); aaaa(f.normal[i]); aaaa(f.rough[i]); aaaa(f.base[i]);
                    // single-channel image: imageLoad(...).aaaa == 1.0
                    f.transitions[i] = unormT ? (int8_t)(uint8_t)255 : (int8_t)127;
                    aaaa(f.roughA[i]); aaaa(f.roughB[i]);
                }
            }
    }

    // ---- fxaa.comp:29-36 writes baseImage, then the host swaps base <-> final (raytracer.cpp:136-140)
    if(flags & ORC_FXAA) {
        std::vector<half4> out(N);
        const Img tex{f.final_, W, H};
#pragma omp parallel for num_threads(threads) schedule(dynamic, 4)
        for(int y = 0; y < H; ++y)
            for(int x = 0; x < W; ++x) out[(size_t)y * W + x] = pack_half4(fxaaPixel(tex, x, y));
        for(size_t i = 0; i < N; ++i) { f.base[i] = f.final_[i]; f.final_[i] = out[i]; }
    }

    // ---- blit rgba16f -> 8-bit, nearest, same extent (render_system.cpp:130-144)
    const bool srgb = (flags & ORC_SRGB8) != 0;
#pragma omp parallel for num_threads(threads) schedule(static)
    for(int y = 0; y < H; ++y)
        for(int x = 0; x < W; ++x) {
            const size_t i = (size_t)y * W + x;
            const vec4 c = unpack_half4(f.final_[i]);
            f.rgba8[4 * i + 0] = toUnorm8(srgb ? srgbOetf(clampf(c.x, 0.0f, 1.0f)) : c.x);
            f.rgba8[4 * i + 1] = toUnorm8(srgb ? srgbOetf(clampf(c.y, 0.0f, 1.0f)) : c.y);
            f.rgba8[4 * i + 2] = toUnorm8(srgb ? srgbOetf(clampf(c.z, 0.0f, 1.0f)) : c.z);
            f.rgba8[4 * i + 3] = toUnorm8(c.w);
        }
}

}  // namespace orc
