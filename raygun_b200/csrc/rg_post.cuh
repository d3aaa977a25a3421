// rg_post.cuh -- parameters of the post-pass kernels (rough_prepare, blur, postprocess, fxaa + 8-bit blit + gather).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace rg {

struct PostParams {
    uint2 *base, *normal, *rough, *final_, *roughA, *roughB, *fxaaOut;  // rgba16f, pitch rw
    signed char* trans;   // R8_SNORM, pitch rw
    uint32_t* blurList;   // rw * rh entries: pixels the blur passes write (built by rough_prepare)
    uint32_t* blurCount;  // [2]: this frame's list length at [blurParity], the other one is cleared for the next frame
    int blurParity;
    uint32_t* rgba8;      // interior region only, pitch ix1 - ix0
    uint32_t* gather;     // optional full-frame RGBA8 target (may be peer memory), pitch W
    int W, H;             // full frame
    int rx0, ry0, rw, rh; // rendered rectangle (region + halo, clipped to the frame)
    int ix0, iy0, ix1, iy1;  // interior region (what this context owns)
    float4 fade;
    int showAlpha;
    uint32_t flags;
};

void launchRoughPrepare(const PostParams& p, cudaStream_t st);
int launchRoughBlur(const PostParams& p, int numSms, cudaStream_t st);     // 10 x (H, V); returns the number of launches
void launchPostprocess(const PostParams& p, cudaStream_t st);
void launchFxaaBlit(const PostParams& p, cudaStream_t st);

}  // namespace rg
