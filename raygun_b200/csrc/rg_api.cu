// rg_api.cu -- the C ABI of librgb200.so (include/rgb200.h): context, uploads, per-frame command stream.
// Mirrors raygun::render::Raytracer (raygun/render/raytracer.{hpp,cpp}) for the one path this library replaces.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rgb200.h"
#include "rg_build.cuh"
#include "rg_post.cuh"
#include "rg_scene.cuh"
#include "rg_trace.cuh"

using namespace rg;

static_assert(sizeof(rg_vertex) == 32 && sizeof(rg_material) == 64 && sizeof(rg_ubo) == 192 && sizeof(rg_instance) == 64 && sizeof(rg_mesh_range) == 16,
              "POD layouts must match the reference's .def files");
static_assert(sizeof(rg_entity) == 64 && sizeof(rg_sphere_body) == 32 && sizeof(rg_timings) % 8 == 0, "extension records of include/rgb200.h");

namespace {

constexpr int kHalo = 40;  // post-chain dependency radius: 1 (prepare) + 10 (blur) + 28 (FXAA search + bilinear) = 39 -> 40

struct MeshBlas {
    rg_mesh_range range{};
    LbvhScratch scratch;
    uint32_t nodeOffset = 0, nNodes = 0, triOffset = 0, nTris = 0;
    bool built = false;
};

enum { EV_AS0, EV_AS1, EV_RT0, EV_RTONLY1, EV_ROUGH0, EV_ROUGH1, EV_POST1, EV_GATHER1, EV_USER0, EV_USER1, EV_TRACE1, EV_PROBE0, EV_PROBE1, EV_N };

}  // namespace

struct rg_ctx {
    int device = 0, numSms = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t side = nullptr;       // the tile order of the NEXT frame is computed here, beside the post chain
    cudaEvent_t evTraceDone = nullptr, evOrderDone = nullptr; bool orderPending = false;
    std::string err;
    uint64_t launches = 0;

    uint32_t W = 0, H = 0;
    int ix0 = 0, iy0 = 0, ix1 = 0, iy1 = 0;   // interior region
    int rx0 = 0, ry0 = 0, rw = 0, rh = 0;     // rendered rectangle (region + halo)
    uint2 *base = nullptr, *normal = nullptr, *rough = nullptr, *final_ = nullptr, *roughA = nullptr, *roughB = nullptr;
    signed char* trans = nullptr;
    uint32_t *blurList = nullptr, *blurCount = nullptr; int blurParity = 0;   // pixels the blur passes write (rg_post.cu)
    uint32_t *rgba8 = nullptr, *idInst = nullptr, *idPrim = nullptr;
    uint2* fxaaOut = nullptr;          // FXAA target (the reference writes baseImage and swaps; pointers stay put here so peers can keep them)
    bool lastFxaa = false;
    // partitioned multi-GPU mode: who traces what, where finished pixels go, and the two cross-GPU barriers per frame
    uint32_t rank = 0, world = 1, frameId = 0;
    TraceParams::Target peers[kMaxPeers]{};
    uint32_t* peerArriveTrace[kMaxPeers]{}; uint32_t* peerArrivePost[kMaxPeers]{};
    void* peerIpcOpened[kMaxPeers][5]{};
    bool peerAttached[kMaxPeers]{};
    uint32_t *arriveTrace = nullptr, *arrivePost = nullptr, *dSyncErr = nullptr;   // [kMaxPeers] each, in this GPU's memory
    uint32_t **dPeerTraceFlags = nullptr, **dPeerPostFlags = nullptr;             // device copies of the peers' flag-array pointers
    // trace scheduling state (see TraceParams): per-sample scratch + per-tile cost history
    float4* sampleScratch = nullptr; uint32_t* sampleDone = nullptr; uint32_t* tileOrder = nullptr; uint32_t* tileCost = nullptr;
    uint32_t schedSlots = 0, schedSamples = 0; uint64_t schedKey = 0; bool haveTileHistory = false;
    uint32_t* gatherOwn = nullptr;     // full-frame buffer owned by this context (GPU 0 role)
    uint32_t* gatherTarget = nullptr;  // where the final kernel stores the region (may be peer memory)

    void* dVertices = nullptr; uint32_t nVertices = 0;
    uint32_t* dIndices = nullptr; uint32_t nIndices = 0;
    void* dMaterials = nullptr; uint32_t nMaterials = 0;
    std::vector<MeshBlas> meshes;
    Node8* blasNodes = nullptr; Tri* tris = nullptr; uint32_t nodeCap = 0, triCap = 0, nodesUsed = 0, trisUsed = 0;
    float* dMeshBoxes = nullptr; float4* dMeshSpheres = nullptr;

    rg_instance* dInstRaw = nullptr; InstTrav* dInstTrav = nullptr; InstShade* dInstShade = nullptr; uint32_t* dMeshRoots = nullptr;
    uint32_t instCap = 0, nInst = 0;
    LbvhScratch tlasScratch;
    Node8* tlasNodes = nullptr; InstTrav* tlasLeaves = nullptr;
    rg_instance* hInstPinned = nullptr; uint32_t hInstCap = 0;

    float* dUbo = nullptr; rg_ubo hUbo{}; rg_ubo* hUboPinned = nullptr; bool uboOnDevice = false;
    unsigned long long* dCounters = nullptr; float4* ctxPool = nullptr;
    rg_entity* dEntities = nullptr; rg_instance* dEntTmp = nullptr; uint32_t* dEntEmit = nullptr; uint32_t* dEntCount = nullptr; uint32_t entCap = 0;
    // trace scheduler (rg_trace.cu: k_trace_lanes / k_trace_pool).  RG_SCHED_AUTO times both on consecutive frames and keeps the
    // faster one; the comparison is repeated every kSchedReprobe frames so a changing scene can change the choice.
    int schedMode = RG_SCHED_AUTO, schedChosen = RG_SCHED_LANES, schedProbe = -1, schedLast = RG_SCHED_LANES;
    float schedMs[2] = {-1.0f, -1.0f};
    uint32_t schedFrames = 0, schedInterval = 256;
    cudaEvent_t ev[EV_N]{};
    bool haveFrame = false, haveAs = false, blasBuilt = false;
    uint32_t lastFlags = 0;
    void* flushBuf = nullptr;
    float lastRaysMs = 0.0f;
    bool debugPostOnly = false;
};

namespace {

int fail(rg_ctx* c, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if(c) c->err = buf;
    return 1;
}
#define CK(call) do { cudaError_t e_ = (call); if(e_ != cudaSuccess) return fail(ctx, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while(0)
#define USE_DEVICE() CK(cudaSetDevice(ctx->device))

void freeImages(rg_ctx* c) {
    cudaFree(c->base); cudaFree(c->normal); cudaFree(c->rough); cudaFree(c->final_); cudaFree(c->roughA); cudaFree(c->roughB); cudaFree(c->trans);
    cudaFree(c->fxaaOut); c->fxaaOut = nullptr;
    cudaFree(c->rgba8); cudaFree(c->idInst); cudaFree(c->idPrim);
    cudaFree(c->blurList); cudaFree(c->blurCount); c->blurList = c->blurCount = nullptr;
    c->base = c->normal = c->rough = c->final_ = c->roughA = c->roughB = nullptr; c->trans = nullptr; c->rgba8 = c->idInst = c->idPrim = nullptr;
}

int allocImages(rg_ctx* ctx) {
    freeImages(ctx);
    ctx->rx0 = ctx->ix0 - kHalo < 0 ? 0 : ctx->ix0 - kHalo;
    ctx->ry0 = ctx->iy0 - kHalo < 0 ? 0 : ctx->iy0 - kHalo;
    const int rx1 = ctx->ix1 + kHalo > (int)ctx->W ? (int)ctx->W : ctx->ix1 + kHalo;
    const int ry1 = ctx->iy1 + kHalo > (int)ctx->H ? (int)ctx->H : ctx->iy1 + kHalo;
    ctx->rw = rx1 - ctx->rx0; ctx->rh = ry1 - ctx->ry0;
    const size_t n = (size_t)ctx->rw * ctx->rh;
    uint2** imgs[7] = {&ctx->base, &ctx->normal, &ctx->rough, &ctx->final_, &ctx->roughA, &ctx->roughB, &ctx->fxaaOut};
    for(auto p: imgs) { CK(cudaMalloc(p, n * sizeof(uint2))); CK(cudaMemsetAsync(*p, 0, n * sizeof(uint2), ctx->stream)); }
    CK(cudaMalloc(&ctx->trans, n)); CK(cudaMemsetAsync(ctx->trans, 0, n, ctx->stream));
    CK(cudaMalloc(&ctx->blurList, n * 4)); CK(cudaMalloc(&ctx->blurCount, 8)); CK(cudaMemsetAsync(ctx->blurCount, 0, 8, ctx->stream)); ctx->blurParity = 0;
    CK(cudaMalloc(&ctx->idInst, n * 4)); CK(cudaMalloc(&ctx->idPrim, n * 4));
    CK(cudaMemsetAsync(ctx->idInst, 0xff, n * 4, ctx->stream)); CK(cudaMemsetAsync(ctx->idPrim, 0xff, n * 4, ctx->stream));
    const size_t ni = (size_t)(ctx->ix1 - ctx->ix0) * (ctx->iy1 - ctx->iy0);
    CK(cudaMalloc(&ctx->rgba8, ni * 4)); CK(cudaMemsetAsync(ctx->rgba8, 0, ni * 4, ctx->stream));
    ctx->haveFrame = false; ctx->lastFxaa = false;
    for(auto& a: ctx->peerAttached) a = false;   // peers must re-attach after a re-allocation
    return 0;
}

// Cross-GPU barrier flags: rank r publishes "frame f reached" by storing f into slot r of every peer's flag array (peer store,
// system-scope fence first so the G-buffer / frame stores of the preceding kernel are visible), and waits on its own array.
__global__ void k_signal(uint32_t* const* peerFlags, uint32_t nPeers, uint32_t myRank, uint32_t frame) {
    __threadfence_system();
    if(threadIdx.x < nPeers && peerFlags[threadIdx.x]) {
        volatile uint32_t* f = peerFlags[threadIdx.x] + myRank;
        *f = frame;
    }
    __threadfence_system();
}
__global__ void k_wait(const uint32_t* myFlags, uint32_t nPeers, uint32_t frame, uint32_t* err) {
    if(threadIdx.x < nPeers) {
        const volatile uint32_t* f = myFlags + threadIdx.x;
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while((int)(*f - frame) < 0) {
            __nanosleep(200);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if(t1 - t0 > 4000000000ull) { atomicExch(err, 1u + threadIdx.x); break; }   // 4 s: a peer is gone; never hang the GPU
        }
    }
    __threadfence_system();
}

int ensureInstanceCapacity(rg_ctx* ctx, uint32_t n) {
    if(n <= ctx->instCap) return 0;
    cudaFree(ctx->dInstRaw); cudaFree(ctx->dInstTrav); cudaFree(ctx->dInstShade); cudaFree(ctx->tlasNodes); cudaFree(ctx->tlasLeaves);
    const uint32_t cap = n + n / 2 + 16;
    CK(cudaMalloc(&ctx->dInstRaw, sizeof(rg_instance) * (size_t)cap));
    CK(cudaMalloc(&ctx->dInstTrav, sizeof(InstTrav) * (size_t)cap));
    CK(cudaMalloc(&ctx->dInstShade, sizeof(InstShade) * (size_t)cap));
    CK(cudaMalloc(&ctx->tlasNodes, sizeof(Node8) * (size_t)cap));
    CK(cudaMalloc(&ctx->tlasLeaves, sizeof(InstTrav) * (size_t)cap * kLeafStride));   // kLeafStride elements per TLAS leaf child
    ctx->instCap = cap;
    ctx->tlasScratch.reserve(cap);
    return 0;
}

int buildTlasFromRaw(rg_ctx* ctx, uint32_t n) {
    ctx->nInst = n;
    CK(cudaEventRecord(ctx->ev[EV_AS0], ctx->stream));
    if(n) {
        const uint64_t before = ctx->tlasScratch.launches;
        buildTlas(ctx->tlasScratch, ctx->dInstRaw, n, ctx->dMeshRoots, (uint32_t)ctx->meshes.size(), ctx->dInstTrav, ctx->dInstShade, ctx->dMeshBoxes,
                  ctx->dMeshSpheres, ctx->tlasNodes, ctx->tlasLeaves, ctx->stream);
        ctx->launches += ctx->tlasScratch.launches - before;
    }
    CK(cudaEventRecord(ctx->ev[EV_AS1], ctx->stream));
    CK(cudaGetLastError());
    ctx->haveAs = true;
    return 0;
}

// (Re)allocate the scheduling buffers when the trace share or the sample count changed; returns non-zero on failure.
int ensureSchedule(rg_ctx* ctx) {
    const bool part = ctx->world > 1;
    const uint32_t dw = part ? ctx->W : (uint32_t)ctx->rw, dh = part ? ctx->H : (uint32_t)ctx->rh;
    const uint32_t slots = traceShareTiles(dw, dh, ctx->rank, ctx->world);
    const uint64_t key = ((uint64_t)dw << 40) ^ ((uint64_t)dh << 20) ^ ((uint64_t)ctx->rank << 8) ^ ctx->world ^ ((uint64_t)ctx->rx0 << 50) ^ ((uint64_t)ctx->ry0 << 12);
    if(slots != ctx->schedSlots || key != ctx->schedKey) {
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaStreamSynchronize(ctx->side)); ctx->orderPending = false;
        cudaFree(ctx->sampleDone); cudaFree(ctx->tileOrder); cudaFree(ctx->tileCost);
        ctx->sampleDone = ctx->tileOrder = ctx->tileCost = nullptr;
        CK(cudaMalloc(&ctx->sampleDone, 4 * (size_t)(slots ? slots : 1) * 32));
        CK(cudaMalloc(&ctx->tileOrder, 4 * (size_t)(slots ? slots : 1)));
        CK(cudaMalloc(&ctx->tileCost, 4 * (size_t)(slots ? slots : 1)));
        CK(cudaMemsetAsync(ctx->sampleDone, 0, 4 * (size_t)(slots ? slots : 1) * 32, ctx->stream));
        CK(cudaMemsetAsync(ctx->tileCost, 0, 4 * (size_t)(slots ? slots : 1), ctx->stream));
        ctx->schedSlots = slots; ctx->schedKey = key; ctx->haveTileHistory = false; ctx->schedSamples = 0;
        cudaFree(ctx->sampleScratch); ctx->sampleScratch = nullptr;
    }
    return 0;
}

// The per-sample scratch (48 B x pixels x numSamples) exists only while a kernel that parks samples runs: the pool kernel, or the
// lanes kernel on a small share of the frame (see lanesSequential).
int ensureSampleScratch(rg_ctx* ctx) {
    const uint32_t S = (uint32_t)ctx->hUbo.num_samples, slots = ctx->schedSlots;
    if(S > 1 && S > ctx->schedSamples) {
        CK(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->sampleScratch); ctx->sampleScratch = nullptr;
        CK(cudaMalloc(&ctx->sampleScratch, sizeof(float4) * 3 * (size_t)(slots ? slots : 1) * 32 * S));
        ctx->schedSamples = S;
    }
    return 0;
}

// Lanes kernel: all numSamples samples of a pixel in one lane, one after the other (sums in registers / local memory, nothing
// parked) when the rank's share has enough tiles to balance whole pixels over the persistent grid; on a small share one work
// item per sample keeps the tail short (a sample's ray tree is sequential).  RGB200_LANES_SEQ=0|1 overrides (developer knob).
bool lanesSequential(rg_ctx* ctx) {
    static const int forced = [] { const char* e = getenv("RGB200_LANES_SEQ"); return e ? atoi(e) : -1; }();
    if(forced >= 0) return forced != 0;
    return false;   // measured (B200, C2): 4.80 ms against 4.43 ms -- see DESIGN.md 4.1
}

constexpr uint32_t kSchedReprobe = 256;   // two probe frames (one per kernel) every 256 frames; every 4 096 when the loser took more than
                                          // 1.5x the winner's time (C2: the pool kernel needs 12 ms against 4 ms -- 1 % of the run at 256)

// Which trace kernel runs this frame.  AUTO: once the heavy-first tile order exists (second frame on), one frame is timed with
// each scheduler (CUDA events around the kernel; the host waits for that one frame's kernel when it needs the number) and the
// faster one is kept for the next kSchedReprobe frames.  Both produce bit-identical images, so switching is invisible.
bool chooseScheduler(rg_ctx* c, uint32_t flags) {
    if(c->schedMode != RG_SCHED_AUTO) { c->schedProbe = -1; c->schedLast = c->schedMode; return c->schedMode == RG_SCHED_POOL; }
    if(c->schedProbe >= 0) {   // harvest the probe frame
        float ms = -1.0f;
        if(cudaEventSynchronize(c->ev[EV_PROBE1]) == cudaSuccess && cudaEventElapsedTime(&ms, c->ev[EV_PROBE0], c->ev[EV_PROBE1]) == cudaSuccess) c->schedMs[c->schedProbe] = ms;
        else c->schedMs[c->schedProbe] = 1e30f;
        c->schedProbe = -1;
        if(c->schedMs[0] >= 0.0f && c->schedMs[1] >= 0.0f) {
            c->schedChosen = c->schedMs[1] < c->schedMs[0] ? RG_SCHED_POOL : RG_SCHED_LANES; c->schedFrames = 0;
            const float lo = fminf(c->schedMs[0], c->schedMs[1]), hi = fmaxf(c->schedMs[0], c->schedMs[1]);
            c->schedInterval = hi > 1.5f * lo ? 16u * kSchedReprobe : kSchedReprobe;
        }
    }
    int mode = c->schedChosen;
    if(c->haveTileHistory && !(flags & RG_COUNT_TRAVERSAL)) {   // the instrumented kernel is slower: never a probe frame
        if(c->schedMs[0] < 0.0f) mode = c->schedProbe = RG_SCHED_LANES;
        else if(c->schedMs[1] < 0.0f) mode = c->schedProbe = RG_SCHED_POOL;
        else if(++c->schedFrames >= c->schedInterval) { c->schedMs[0] = c->schedMs[1] = -1.0f; mode = c->schedProbe = RG_SCHED_LANES; }
    }
    c->schedLast = mode;
    return mode == RG_SCHED_POOL;
}

void fillTraceParams(rg_ctx* c, TraceParams& p, uint32_t flags) {
    p.tlasNodes = c->tlasNodes; p.tlasLeaves = c->tlasLeaves; p.blasNodes = c->blasNodes; p.tris = c->tris; p.instShade = c->dInstShade;
    p.vertices = (const float4*)c->dVertices; p.indices = c->dIndices; p.materials = (const float4*)c->dMaterials; p.ubo = c->dUbo;
    p.nInst = c->nInst; p.W = c->W; p.H = c->H;
    p.rank = c->rank; p.world = c->world;
    const TraceParams::Target self{c->base, c->normal, c->rough, c->rx0, c->ry0, c->rw, c->rh};
    p.sx0 = c->rx0; p.sy0 = c->ry0; p.sw = c->rw; p.sh = c->rh;
    if(c->world > 1) {   // partitioned: the whole frame is the domain, every rank's rectangle is a store target
        p.dx0 = 0; p.dy0 = 0; p.dw = c->W; p.dh = c->H;
        for(uint32_t q = 0; q < c->world; ++q) p.targets[q] = (q == c->rank) ? self : c->peers[q];
        p.nTargets = c->world; p.self = c->rank;
    } else {             // single GPU, or overdraw mode: trace exactly the own rectangle
        p.dx0 = (uint32_t)c->rx0; p.dy0 = (uint32_t)c->ry0; p.dw = (uint32_t)c->rw; p.dh = (uint32_t)c->rh;
        p.targets[0] = self; p.nTargets = 1; p.self = 0;
    }
    {
        auto magic = [](uint32_t d) { return d <= 1u ? 0xffffffffu : (uint32_t)(0x100000000ull / d); };
        p.magicS = magic((uint32_t)c->hUbo.num_samples); p.magicTilesX = magic((p.dw + 7u) / 8u);
    }
    p.idInst = (flags & RG_DEBUG_IDS) ? c->idInst : nullptr; p.idPrim = (flags & RG_DEBUG_IDS) ? c->idPrim : nullptr;
    p.workCounter = reinterpret_cast<uint32_t*>(c->dCounters + 16); p.counters = c->dCounters; p.flags = flags; p.ctxPool = c->ctxPool;
    p.sampleScratch = c->sampleScratch; p.sampleDone = c->sampleDone;
    p.tileOrder = c->haveTileHistory ? c->tileOrder : nullptr; p.tileCost = c->tileCost;
    // miss.rmiss:63-66 (oracle/orc_shade.cpp skyMix): scatter = 1 - clamp(pow(4 - lightDir.y, 1/15), .8, 1);
    // scatterColor = mix(vec3(1), vec3(1, .3, 0) * 1.5, scatter) -- per frame, not per ray
    {
        float scatter = std::pow(4.0f - c->hUbo.light_dir[1], 1.0f / 15.0f);
        scatter = scatter < 0.8f ? 0.8f : (scatter > 1.0f ? 1.0f : scatter);
        scatter = 1.0f - scatter;
        const float tone[3] = {1.0f * 1.5f, 0.3f * 1.5f, 0.0f * 1.5f};
        for(int k = 0; k < 3; ++k) p.scatterColor[k] = 1.0f * (1.0f - scatter) + tone[k] * scatter;
    }
}

void fillPostParams(rg_ctx* c, PostParams& p, uint32_t flags) {
    p.base = c->base; p.normal = c->normal; p.rough = c->rough; p.final_ = c->final_; p.roughA = c->roughA; p.roughB = c->roughB;
    p.fxaaOut = c->fxaaOut;  // fxaa.comp:35 writes baseImage and the host swaps base <-> final (raytracer.cpp:138-140); here the FXAA
                             // result has its own buffer and rg_read_image maps the selectors, so image pointers never move
    p.blurList = c->blurList; p.blurCount = c->blurCount; p.blurParity = c->blurParity;
    p.trans = c->trans; p.rgba8 = c->rgba8; p.gather = (flags & RG_NO_GATHER) ? nullptr : c->gatherTarget;
    p.W = (int)c->W; p.H = (int)c->H; p.rx0 = c->rx0; p.ry0 = c->ry0; p.rw = c->rw; p.rh = c->rh;
    p.ix0 = c->ix0; p.iy0 = c->iy0; p.ix1 = c->ix1; p.iy1 = c->iy1;
    p.fade = make_float4(c->hUbo.fade_color[0], c->hUbo.fade_color[1], c->hUbo.fade_color[2], c->hUbo.fade_color[3]);
    p.showAlpha = (c->hUbo.show_alpha & 0xffu) != 0;
    p.flags = flags;
}

int runPost(rg_ctx* ctx, uint32_t flags) {
    ctx->blurParity ^= 1;
    PostParams pp; fillPostParams(ctx, pp, flags);
    CK(cudaEventRecord(ctx->ev[EV_ROUGH0], ctx->stream));
    launchRoughPrepare(pp, ctx->stream);
    const int blurLaunches = launchRoughBlur(pp, ctx->numSms, ctx->stream);
    CK(cudaEventRecord(ctx->ev[EV_ROUGH1], ctx->stream));
    launchPostprocess(pp, ctx->stream);
    launchFxaaBlit(pp, ctx->stream);
    ctx->launches += 1 + blurLaunches + 1 + 1;
    ctx->lastFxaa = (flags & RG_FXAA) != 0;
    CK(cudaEventRecord(ctx->ev[EV_POST1], ctx->stream));
    CK(cudaEventRecord(ctx->ev[EV_GATHER1], ctx->stream));
    CK(cudaGetLastError());
    return 0;
}

int readRegion(rg_ctx* ctx, const void* dImg, size_t bpp, void* dst) {
    const size_t w = (size_t)(ctx->ix1 - ctx->ix0), h = (size_t)(ctx->iy1 - ctx->iy0);
    const char* src = (const char*)dImg + ((size_t)(ctx->iy0 - ctx->ry0) * ctx->rw + (size_t)(ctx->ix0 - ctx->rx0)) * bpp;
    CK(cudaMemcpy2DAsync(dst, w * bpp, src, (size_t)ctx->rw * bpp, w * bpp, h, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

}  // namespace

extern "C" {

int rg_create(rg_ctx** out, int cuda_device, uint32_t width, uint32_t height) {
    if(!out) return 1;
    *out = nullptr;
    int nDev = 0;
    if(cudaGetDeviceCount(&nDev) != cudaSuccess || nDev <= 0 || cuda_device < 0 || cuda_device >= nDev) {
        fprintf(stderr, "rgb200: no CUDA device %d available (there is no CPU fallback)\n", cuda_device);
        return 2;
    }
    if(width == 0 || height == 0) return 3;
    rg_ctx* ctx = new rg_ctx();
    ctx->device = cuda_device;
    if(cudaSetDevice(cuda_device) != cudaSuccess) { delete ctx; return 4; }
    cudaDeviceProp prop{};
    cudaGetDeviceProperties(&prop, cuda_device);
    ctx->numSms = prop.multiProcessorCount;
    if(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return 5; }
    if(cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking) != cudaSuccess) { cudaStreamDestroy(ctx->stream); delete ctx; return 5; }
    if(const char* e = getenv("RGB200_TRACE_SCHED"))   // developer override for A/B timing: lanes | pool | auto
        ctx->schedMode = !strcmp(e, "lanes") ? RG_SCHED_LANES : (!strcmp(e, "pool") ? RG_SCHED_POOL : RG_SCHED_AUTO);
    // every allocation is checked: a half-initialised context must not reach the caller
    bool ok = true;
    auto good = [&](cudaError_t e) { ok = ok && e == cudaSuccess; };
    for(auto& e: ctx->ev) good(cudaEventCreate(&e));
    good(cudaEventCreateWithFlags(&ctx->evTraceDone, cudaEventDisableTiming)); good(cudaEventCreateWithFlags(&ctx->evOrderDone, cudaEventDisableTiming));
    good(cudaMalloc(&ctx->dUbo, 192)); good(cudaMalloc(&ctx->dCounters, 17 * 8 + 17 * 4));   // 16 counters + the trace kernels' work counter (one memset per frame) + the tile sort's scratch
    good(cudaMalloc(&ctx->ctxPool, tracePoolBytes(ctx->numSms)));
    good(cudaMallocHost(&ctx->hUboPinned, sizeof(rg_ubo)));
    good(cudaMalloc(&ctx->arriveTrace, 4 * kMaxPeers)); good(cudaMalloc(&ctx->arrivePost, 4 * kMaxPeers)); good(cudaMalloc(&ctx->dSyncErr, 4));
    good(cudaMalloc(&ctx->dPeerTraceFlags, sizeof(void*) * kMaxPeers)); good(cudaMalloc(&ctx->dPeerPostFlags, sizeof(void*) * kMaxPeers));
    if(ok) {
        good(cudaMemset(ctx->dUbo, 0, 192)); good(cudaMemset(ctx->dCounters, 0, 17 * 8 + 17 * 4));
        good(cudaMemset(ctx->arriveTrace, 0, 4 * kMaxPeers)); good(cudaMemset(ctx->arrivePost, 0, 4 * kMaxPeers)); good(cudaMemset(ctx->dSyncErr, 0, 4));
        good(cudaMemset(ctx->dPeerTraceFlags, 0, sizeof(void*) * kMaxPeers)); good(cudaMemset(ctx->dPeerPostFlags, 0, sizeof(void*) * kMaxPeers));
    }
    if(!ok) {
        fprintf(stderr, "rgb200: rg_create: device allocation failed: %s\n", cudaGetErrorString(cudaGetLastError()));
        rg_destroy(ctx);
        return 7;
    }
    ctx->W = width; ctx->H = height; ctx->ix0 = 0; ctx->iy0 = 0; ctx->ix1 = (int)width; ctx->iy1 = (int)height;
    if(allocImages(ctx)) { fprintf(stderr, "rgb200: %s\n", ctx->err.c_str()); rg_destroy(ctx); return 6; }
    *out = ctx;
    return 0;
}

void rg_destroy(rg_ctx* ctx) {
    if(!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if(ctx->side) cudaStreamSynchronize(ctx->side);
    freeImages(ctx);
    for(auto& pr: ctx->peerIpcOpened) for(void* ptr: pr) if(ptr) cudaIpcCloseMemHandle(ptr);
    cudaFree(ctx->arriveTrace); cudaFree(ctx->arrivePost); cudaFree(ctx->dSyncErr); cudaFree(ctx->dPeerTraceFlags); cudaFree(ctx->dPeerPostFlags);
    cudaFree(ctx->sampleScratch); cudaFree(ctx->sampleDone); cudaFree(ctx->tileOrder); cudaFree(ctx->tileCost);
    cudaFree(ctx->gatherOwn); cudaFree(ctx->flushBuf);
    cudaFree(ctx->dVertices); cudaFree(ctx->dIndices); cudaFree(ctx->dMaterials); cudaFree(ctx->blasNodes); cudaFree(ctx->tris); cudaFree(ctx->dMeshBoxes); cudaFree(ctx->dMeshSpheres);
    cudaFree(ctx->dInstRaw); cudaFree(ctx->dInstTrav); cudaFree(ctx->dInstShade); cudaFree(ctx->dMeshRoots); cudaFree(ctx->tlasNodes); cudaFree(ctx->tlasLeaves);
    cudaFree(ctx->dUbo); cudaFree(ctx->dCounters); cudaFree(ctx->ctxPool);
    cudaFree(ctx->dEntities); cudaFree(ctx->dEntTmp); cudaFree(ctx->dEntEmit); cudaFree(ctx->dEntCount);
    cudaFreeHost(ctx->hInstPinned); cudaFreeHost(ctx->hUboPinned);
    for(auto& m: ctx->meshes) m.scratch.release();
    ctx->tlasScratch.release();
    for(auto& e: ctx->ev) cudaEventDestroy(e);
    if(ctx->evTraceDone) cudaEventDestroy(ctx->evTraceDone);
    if(ctx->evOrderDone) cudaEventDestroy(ctx->evOrderDone);
    if(ctx->side) cudaStreamDestroy(ctx->side);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* rg_last_error(const rg_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
uint64_t rg_launch_count(const rg_ctx* ctx) { return ctx ? ctx->launches : 0; }

int rg_resize(rg_ctx* ctx, uint32_t width, uint32_t height) {
    if(!ctx || !width || !height) return fail(ctx, "rg_resize: bad size");
    USE_DEVICE();
    CK(cudaStreamSynchronize(ctx->stream));
    if(ctx->world > 1) return fail(ctx, "rg_resize: the context is in partitioned mode; peers hold pointers into its images -- call rg_peer_detach_all and rg_set_partition(ctx, 0, 1) on every rank first");
    ctx->W = width; ctx->H = height; ctx->ix0 = 0; ctx->iy0 = 0; ctx->ix1 = (int)width; ctx->iy1 = (int)height;
    // the gather buffer (own or a peer's) has the old frame's stride: the caller sets a target for the new size again
    cudaFree(ctx->gatherOwn); ctx->gatherOwn = nullptr; ctx->gatherTarget = nullptr;
    return allocImages(ctx);
}

int rg_set_region(rg_ctx* ctx, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1) {
    if(!ctx) return 1;
    if(x0 >= x1 || y0 >= y1 || x1 > ctx->W || y1 > ctx->H) return fail(ctx, "rg_set_region: bad rectangle");
    // The images are re-allocated: ranks that attached this one (rg_peer_attach) would keep storing G-buffer pixels into freed memory.
    if(ctx->world > 1) return fail(ctx, "rg_set_region: the context is in partitioned mode; peers hold pointers into its images -- call rg_peer_detach_all and rg_set_partition(ctx, 0, 1) on every rank first");
    USE_DEVICE();
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->ix0 = (int)x0; ctx->iy0 = (int)y0; ctx->ix1 = (int)x1; ctx->iy1 = (int)y1;
    return allocImages(ctx);
}

int rg_upload_geometry(rg_ctx* ctx, const rg_vertex* vertices, uint32_t n_vertices, const uint32_t* indices, uint32_t n_indices, const rg_mesh_range* meshes,
                       uint32_t n_meshes) {
    if(!ctx) return 1;
    if((n_vertices && !vertices) || (n_indices && !indices) || (n_meshes && !meshes)) return fail(ctx, "rg_upload_geometry: null input");
    for(uint32_t m = 0; m < n_meshes; ++m) {
        const rg_mesh_range& r = meshes[m];
        if((uint64_t)r.vtx_off + r.vtx_cnt > n_vertices || (uint64_t)r.idx_off + r.idx_cnt > n_indices || r.idx_cnt % 3)
            return fail(ctx, "rg_upload_geometry: mesh %u out of range", m);
        for(uint32_t k = 0; k < r.idx_cnt; ++k)
            if(indices[r.idx_off + k] >= r.vtx_cnt) return fail(ctx, "rg_upload_geometry: mesh %u index %u out of range", m, k);
    }
    USE_DEVICE();
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->dVertices); cudaFree(ctx->dIndices); ctx->dVertices = nullptr; ctx->dIndices = nullptr;
    CK(cudaMalloc(&ctx->dVertices, (size_t)(n_vertices ? n_vertices : 1) * 32));
    CK(cudaMalloc(&ctx->dIndices, (size_t)(n_indices ? n_indices : 1) * 4));
    CK(cudaMemcpyAsync(ctx->dVertices, vertices, (size_t)n_vertices * 32, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->dIndices, indices, (size_t)n_indices * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->nVertices = n_vertices; ctx->nIndices = n_indices;
    for(auto& m: ctx->meshes) m.scratch.release();
    ctx->meshes.clear(); ctx->meshes.resize(n_meshes);
    for(uint32_t m = 0; m < n_meshes; ++m) ctx->meshes[m].range = meshes[m];
    ctx->nodesUsed = ctx->trisUsed = 0;
    ctx->haveAs = false; ctx->blasBuilt = false;
    cudaFree(ctx->dMeshBoxes); cudaFree(ctx->dMeshRoots); cudaFree(ctx->dMeshSpheres); ctx->dMeshBoxes = nullptr; ctx->dMeshRoots = nullptr; ctx->dMeshSpheres = nullptr;
    CK(cudaMalloc(&ctx->dMeshBoxes, sizeof(float) * 6 * ((size_t)n_meshes + 1)));
    CK(cudaMalloc(&ctx->dMeshSpheres, sizeof(float4) * ((size_t)n_meshes + 1)));
    CK(cudaMalloc(&ctx->dMeshRoots, sizeof(uint32_t) * ((size_t)n_meshes + 1)));
    return 0;
}

int rg_upload_materials(rg_ctx* ctx, const rg_material* materials, uint32_t n_materials) {
    if(!ctx) return 1;
    if(n_materials && !materials) return fail(ctx, "rg_upload_materials: null input");
    USE_DEVICE();
    if(n_materials > ctx->nMaterials || !ctx->dMaterials) {
        CK(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->dMaterials); ctx->dMaterials = nullptr;
        CK(cudaMalloc(&ctx->dMaterials, (size_t)(n_materials ? n_materials : 1) * 64));
    }
    CK(cudaMemcpyAsync(ctx->dMaterials, materials, (size_t)n_materials * 64, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->nMaterials = n_materials;
    return 0;
}

int rg_build_blas(rg_ctx* ctx) {
    if(!ctx) return 1;
    USE_DEVICE();
    uint64_t totalTris = 0;
    for(auto& m: ctx->meshes) totalTris += m.range.idx_cnt / 3;
    // every leaf child of a wide node reserves kLeafStride elements of the primitive array (rg_types.cuh); at worst one triangle per leaf
    if(totalTris * kLeafStride + 1 >= 0x7fffffffull) return fail(ctx, "rg_build_blas: %llu triangles exceed the primitive index range", (unsigned long long)totalTris);
    const uint32_t needNodes = (uint32_t)(totalTris + ctx->meshes.size() + 1), needTris = (uint32_t)(kLeafStride * totalTris + 1);
    bool rebuildAll = false;
    if(needNodes > ctx->nodeCap || needTris > ctx->triCap) {
        CK(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->blasNodes); cudaFree(ctx->tris); ctx->blasNodes = nullptr; ctx->tris = nullptr;
        CK(cudaMalloc(&ctx->blasNodes, sizeof(Node8) * (size_t)needNodes));
        CK(cudaMalloc(&ctx->tris, sizeof(Tri) * (size_t)needTris));
        ctx->nodeCap = needNodes; ctx->triCap = needTris; ctx->nodesUsed = ctx->trisUsed = 0;
        rebuildAll = true;
    }
    std::vector<uint32_t> roots(ctx->meshes.size() + 1, kInvalid);
    for(size_t m = 0; m < ctx->meshes.size(); ++m) {
        MeshBlas& mb = ctx->meshes[m];
        if(rebuildAll) mb.built = false;
        if(!mb.built) {  // only missing BLASes are built (raytracer.cpp:65-69)
            mb.nodeOffset = ctx->nodesUsed; mb.triOffset = ctx->trisUsed;
            TriSource src{ctx->dVertices, ctx->dIndices, mb.range.vtx_off, mb.range.idx_off, mb.range.idx_cnt / 3};
            const uint64_t before = mb.scratch.launches;
            buildBlas(mb.scratch, src, ctx->blasNodes, mb.nodeOffset, ctx->tris, mb.triOffset, &mb.nNodes, &mb.nTris, ctx->dMeshBoxes + 6 * m, ctx->dMeshSpheres + m, ctx->stream);
            ctx->launches += mb.scratch.launches - before;
            ctx->nodesUsed += mb.nNodes; ctx->trisUsed += mb.nTris;
            mb.built = true;
        }
        roots[m] = mb.nTris ? mb.nodeOffset : kInvalid;
    }
    CK(cudaMemcpyAsync(ctx->dMeshRoots, roots.data(), sizeof(uint32_t) * roots.size(), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    ctx->blasBuilt = true;
    return 0;
}

static int refitBlasCommon(rg_ctx* ctx, uint32_t mesh, const rg_vertex* new_vertices, cudaMemcpyKind kind, const char* who) {
    if(!ctx) return 1;
    if(mesh >= ctx->meshes.size() || !ctx->meshes[mesh].built) return fail(ctx, "%s: mesh %u has no BLAS", who, mesh);
    if(!new_vertices) return fail(ctx, "%s: null input", who);
    USE_DEVICE();
    MeshBlas& mb = ctx->meshes[mesh];
    CK(cudaMemcpyAsync((char*)ctx->dVertices + (size_t)mb.range.vtx_off * 32, new_vertices, (size_t)mb.range.vtx_cnt * 32, kind, ctx->stream));
    TriSource src{ctx->dVertices, ctx->dIndices, mb.range.vtx_off, mb.range.idx_off, mb.range.idx_cnt / 3};
    const uint64_t before = mb.scratch.launches;
    CK(cudaEventRecord(ctx->ev[EV_AS0], ctx->stream));
    refitBlas(mb.scratch, src, ctx->blasNodes, mb.nodeOffset, mb.nNodes, ctx->tris, mb.triOffset, ctx->dMeshBoxes + 6 * mesh, ctx->dMeshSpheres + mesh, ctx->stream);
    CK(cudaEventRecord(ctx->ev[EV_AS1], ctx->stream));
    ctx->launches += mb.scratch.launches - before;
    CK(cudaGetLastError());
    return 0;
}

int rg_refit_blas(rg_ctx* ctx, uint32_t mesh, const rg_vertex* new_vertices) {
    return refitBlasCommon(ctx, mesh, new_vertices, cudaMemcpyHostToDevice, "rg_refit_blas");
}

// The animated vertices already live in device memory (a skinning / physics kernel wrote them): no host round trip, the copy into the
// library's vertex buffer (which the closest-hit shading reads) is device to device on the library's stream.
int rg_refit_blas_device(rg_ctx* ctx, uint32_t mesh, const rg_vertex* d_new_vertices) {
    return refitBlasCommon(ctx, mesh, d_new_vertices, cudaMemcpyDeviceToDevice, "rg_refit_blas_device");
}

int rg_set_instances(rg_ctx* ctx, const rg_instance* instances, uint32_t n_instances) {
    if(!ctx) return 1;
    if(n_instances && !instances) return fail(ctx, "rg_set_instances: null input");
    if(!ctx->blasBuilt) return fail(ctx, "rg_set_instances: call rg_build_blas first");
    for(uint32_t i = 0; i < n_instances; ++i)
        if(instances[i].mesh >= ctx->meshes.size()) return fail(ctx, "rg_set_instances: instance %u references mesh %u", i, instances[i].mesh);
    USE_DEVICE();
    if(ensureInstanceCapacity(ctx, n_instances ? n_instances : 1)) return 1;
    if(n_instances > ctx->hInstCap) {
        CK(cudaStreamSynchronize(ctx->stream));
        cudaFreeHost(ctx->hInstPinned); ctx->hInstPinned = nullptr;
        CK(cudaMallocHost(&ctx->hInstPinned, sizeof(rg_instance) * (size_t)(n_instances + n_instances / 2 + 16)));
        ctx->hInstCap = n_instances + n_instances / 2 + 16;
    }
    if(n_instances) {
        CK(cudaStreamSynchronize(ctx->stream));  // the pinned staging buffer may still be in flight
        memcpy(ctx->hInstPinned, instances, sizeof(rg_instance) * (size_t)n_instances);
        CK(cudaMemcpyAsync(ctx->dInstRaw, ctx->hInstPinned, sizeof(rg_instance) * (size_t)n_instances, cudaMemcpyHostToDevice, ctx->stream));
    }
    return buildTlasFromRaw(ctx, n_instances);
}

int rg_set_instances_device(rg_ctx* ctx, const rg_instance* d_instances, uint32_t n_instances) {
    if(!ctx) return 1;
    if(!ctx->blasBuilt) return fail(ctx, "rg_set_instances_device: call rg_build_blas first");
    USE_DEVICE();
    if(ensureInstanceCapacity(ctx, n_instances ? n_instances : 1)) return 1;
    if(n_instances) CK(cudaMemcpyAsync(ctx->dInstRaw, d_instances, sizeof(rg_instance) * (size_t)n_instances, cudaMemcpyDeviceToDevice, ctx->stream));
    return buildTlasFromRaw(ctx, n_instances);
}

static int setEntities(rg_ctx* ctx, const rg_entity* src, bool onHost, uint32_t n, uint32_t* nOut) {
    if(!ctx) return 1;
    if(n && !src) return fail(ctx, "rg_set_entities: null input");
    if(!ctx->blasBuilt) return fail(ctx, "rg_set_entities: call rg_build_blas first");
    USE_DEVICE();
    if(onHost) {   // validate the order and the nesting depth (the device walk keeps kMaxEntityDepth ancestors)
        std::vector<uint16_t> depth(n);
        for(uint32_t i = 0; i < n; ++i) {
            if(src[i].parent >= (int32_t)i) return fail(ctx, "rg_set_entities: entity %u has parent %d; entities must be in DFS pre-order (parents first)", i, src[i].parent);
            depth[i] = src[i].parent < 0 ? 1 : (uint16_t)(depth[(size_t)src[i].parent] + 1);
            if(depth[i] > kMaxEntityDepth) return fail(ctx, "rg_set_entities: entity %u is nested %u deep; the device walk supports %d levels", i, (unsigned)depth[i], kMaxEntityDepth);
        }
    }
    if(n > ctx->entCap) {
        CK(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->dEntities); cudaFree(ctx->dEntTmp); cudaFree(ctx->dEntEmit);
        ctx->entCap = n + n / 2 + 16;
        CK(cudaMalloc(&ctx->dEntities, sizeof(rg_entity) * (size_t)ctx->entCap));
        CK(cudaMalloc(&ctx->dEntTmp, sizeof(rg_instance) * (size_t)ctx->entCap));
        CK(cudaMalloc(&ctx->dEntEmit, 4 * (size_t)ctx->entCap));
    }
    if(!ctx->dEntCount) CK(cudaMalloc(&ctx->dEntCount, 4));
    if(ensureInstanceCapacity(ctx, n ? n : 1)) return 1;   // at most one instance per entity
    if(n) CK(cudaMemcpyAsync(ctx->dEntities, src, sizeof(rg_entity) * (size_t)n, onHost ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, ctx->stream));
    launchEntityInstances(ctx->dEntities, n, ctx->dEntTmp, ctx->dEntEmit, ctx->dInstRaw, ctx->dEntCount, ctx->stream);
    ctx->launches += n ? 2 : 0;
    uint32_t count = 0;   // the builder sizes its launches on the host: one 4-byte read-back per frame
    CK(cudaMemcpyAsync(&count, ctx->dEntCount, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if(nOut) *nOut = count;
    return buildTlasFromRaw(ctx, count);
}
int rg_set_entities(rg_ctx* ctx, const rg_entity* entities, uint32_t n_entities, uint32_t* n_instances_out) {
    return setEntities(ctx, entities, true, n_entities, n_instances_out);
}
int rg_set_entities_device(rg_ctx* ctx, const rg_entity* d_entities, uint32_t n_entities, uint32_t* n_instances_out) {
    return setEntities(ctx, d_entities, false, n_entities, n_instances_out);
}
int rg_physics_step_spheres(rg_ctx* ctx, rg_entity* d_entities, rg_sphere_body* d_bodies, uint32_t n, float dt, float floor_y) {
    if(!ctx) return 1;
    if(n && (!d_entities || !d_bodies)) return fail(ctx, "rg_physics_step_spheres: null input");
    USE_DEVICE();
    launchStepSpheres(d_entities, d_bodies, n, dt, floor_y, ctx->stream);
    ctx->launches += n ? 1 : 0;
    CK(cudaGetLastError());
    return 0;
}
int rg_debug_read_instances(rg_ctx* ctx, rg_instance* out, uint32_t capacity, uint32_t* n_out) {
    if(!ctx || !n_out) return 1;
    USE_DEVICE();
    CK(cudaStreamSynchronize(ctx->stream));
    *n_out = ctx->nInst;
    const uint32_t n = ctx->nInst < capacity ? ctx->nInst : capacity;
    if(n && out) CK(cudaMemcpy(out, ctx->dInstRaw, sizeof(rg_instance) * (size_t)n, cudaMemcpyDeviceToHost));
    return 0;
}

int rg_set_ubo(rg_ctx* ctx, const rg_ubo* ubo) {
    if(!ctx || !ubo) return fail(ctx, "rg_set_ubo: null input");
    if(ubo->num_samples < 1 || ubo->num_samples > 64) return fail(ctx, "rg_set_ubo: numSamples %d out of range 1..64", ubo->num_samples);
    if(ubo->max_recursions < 0 || ubo->max_recursions > kMaxRecursions) return fail(ctx, "rg_set_ubo: maxRecursions %d out of range 0..%d", ubo->max_recursions, kMaxRecursions);
    USE_DEVICE();
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->hUbo = *ubo;
    *ctx->hUboPinned = *ubo;
    CK(cudaMemcpyAsync(ctx->dUbo, ctx->hUboPinned, 192, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

int rg_set_ubo_device(rg_ctx* ctx, const rg_ubo* d_ubo) {
    if(!ctx || !d_ubo) return fail(ctx, "rg_set_ubo_device: null input");
    USE_DEVICE();
    CK(cudaMemcpyAsync(ctx->dUbo, d_ubo, 192, cudaMemcpyDeviceToDevice, ctx->stream));
    // fade / showAlpha are kernel parameters of the post passes: fetch the tail of the UBO
    CK(cudaMemcpyAsync(ctx->hUboPinned, d_ubo, 192, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const rg_ubo& u = *ctx->hUboPinned;
    if(u.num_samples < 1 || u.num_samples > 64 || u.max_recursions < 0 || u.max_recursions > kMaxRecursions) {   // same bars as rg_set_ubo
        const int ns = u.num_samples, mr = u.max_recursions;
        *ctx->hUboPinned = ctx->hUbo;   // put the previous (valid) block back: the kernels divide by numSamples
        CK(cudaMemcpyAsync(ctx->dUbo, ctx->hUboPinned, 192, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        return fail(ctx, "rg_set_ubo_device: numSamples %d (1..64) or maxRecursions %d (0..%d) out of range", ns, mr, kMaxRecursions);
    }
    ctx->hUbo = u;
    return 0;
}

int rg_render(rg_ctx* ctx, uint32_t flags) {
    if(!ctx) return 1;
    if(!ctx->haveAs) return fail(ctx, "rg_render: no acceleration structure (rg_build_blas + rg_set_instances first)");
    USE_DEVICE();
    const bool partitioned = ctx->world > 1;
    if(partitioned) {
        for(uint32_t q = 0; q < ctx->world; ++q)
            if(q != ctx->rank && !ctx->peerAttached[q]) return fail(ctx, "rg_render: partitioned mode but peer %u is not attached (rg_peer_attach)", q);
        ctx->frameId++;
        // nobody may still be reading last frame's G-buffer when new pixels start to land in it
        k_wait<<<1, 32, 0, ctx->stream>>>(ctx->arrivePost, ctx->world, ctx->frameId - 1, ctx->dSyncErr);
        ctx->launches++;
    }
    if(ensureSchedule(ctx)) return 1;
    // the tile sort of the previous frame (side stream) reads the ray counters and writes the order this frame uses
    if(ctx->orderPending) { CK(cudaStreamWaitEvent(ctx->stream, ctx->evOrderDone, 0)); ctx->orderPending = false; }
    CK(cudaMemsetAsync(ctx->dCounters, 0, 17 * 8, ctx->stream));   // ray counters + work counter
    CK(cudaEventRecord(ctx->ev[EV_RT0], ctx->stream));
    TraceParams tp; fillTraceParams(ctx, tp, flags);
    const bool pool = chooseScheduler(ctx, flags);
    const bool seq = !pool && lanesSequential(ctx);
    if(!seq) { if(ensureSampleScratch(ctx)) return 1; tp.sampleScratch = ctx->sampleScratch; }
    if(ctx->schedProbe >= 0) CK(cudaEventRecord(ctx->ev[EV_PROBE0], ctx->stream));
    launchTrace(tp, ctx->numSms, pool, seq, ctx->stream);
    if(ctx->schedProbe >= 0) CK(cudaEventRecord(ctx->ev[EV_PROBE1], ctx->stream));
    ctx->launches++;
    CK(cudaEventRecord(ctx->ev[EV_TRACE1], ctx->stream));
    {   // heavy-first tile order for the NEXT frame from this frame's per-tile ray counts: a one-block kernel, run beside the post chain
        CK(cudaEventRecord(ctx->evTraceDone, ctx->stream));
        CK(cudaStreamWaitEvent(ctx->side, ctx->evTraceDone, 0));
        launchOrderTiles(ctx->tileCost, ctx->schedSlots, ctx->tileOrder, ctx->dCounters, reinterpret_cast<uint32_t*>(ctx->dCounters + 17), ctx->numSms, ctx->side);
        CK(cudaEventRecord(ctx->evOrderDone, ctx->side));
        ctx->launches += 2;
        ctx->orderPending = true; ctx->haveTileHistory = true;
    }
    if(partitioned) {   // every rank's share of this rectangle has landed once all ranks signalled
        k_signal<<<1, 32, 0, ctx->stream>>>(ctx->dPeerTraceFlags, ctx->world, ctx->rank, ctx->frameId);
        k_wait<<<1, 32, 0, ctx->stream>>>(ctx->arriveTrace, ctx->world, ctx->frameId, ctx->dSyncErr);
        ctx->launches += 2;
    }
    CK(cudaEventRecord(ctx->ev[EV_RTONLY1], ctx->stream));
    ctx->debugPostOnly = false;
    if(runPost(ctx, flags)) return 1;
    if(partitioned) {
        k_signal<<<1, 32, 0, ctx->stream>>>(ctx->dPeerPostFlags, ctx->world, ctx->rank, ctx->frameId);
        ctx->launches++;
    }
    ctx->haveFrame = true; ctx->lastFlags = flags;
    return 0;
}

int rg_sync(rg_ctx* ctx) {
    if(!ctx) return 1;
    USE_DEVICE();
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    return 0;
}

int rg_read_rgba8(rg_ctx* ctx, void* dst) {
    if(!ctx || !dst) return fail(ctx, "rg_read_rgba8: null");
    USE_DEVICE();
    const size_t n = (size_t)(ctx->ix1 - ctx->ix0) * (ctx->iy1 - ctx->iy0) * 4;
    CK(cudaMemcpyAsync(dst, ctx->rgba8, n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int rg_read_image(rg_ctx* ctx, int which, void* dst) {
    if(!ctx || !dst) return fail(ctx, "rg_read_image: null");
    USE_DEVICE();
    switch(which) {
    // after an FXAA frame the reference has swapped the images: "final" is the FXAA output, "base" the postprocess output
    case RG_IMG_FINAL: return readRegion(ctx, ctx->lastFxaa ? ctx->fxaaOut : ctx->final_, 8, dst);
    case RG_IMG_BASE: return readRegion(ctx, ctx->lastFxaa ? ctx->final_ : ctx->base, 8, dst);
    case RG_IMG_NORMAL: return readRegion(ctx, ctx->normal, 8, dst);
    case RG_IMG_ROUGH: return readRegion(ctx, ctx->rough, 8, dst);
    case RG_IMG_TRANSITIONS: return readRegion(ctx, ctx->trans, 1, dst);
    case RG_IMG_ROUGH_A: return readRegion(ctx, ctx->roughA, 8, dst);
    case RG_IMG_ROUGH_B: return readRegion(ctx, ctx->roughB, 8, dst);
    default: return fail(ctx, "rg_read_image: bad selector %d", which);
    }
}

int rg_read_ids(rg_ctx* ctx, uint32_t* instance_ids, uint32_t* primitive_ids) {
    if(!ctx) return 1;
    USE_DEVICE();
    if(instance_ids && readRegion(ctx, ctx->idInst, 4, instance_ids)) return 1;
    if(primitive_ids && readRegion(ctx, ctx->idPrim, 4, primitive_ids)) return 1;
    return 0;
}

int rg_get_timings(rg_ctx* ctx, rg_timings* out) {
    if(!ctx || !out) return fail(ctx, "rg_get_timings: null");
    USE_DEVICE();
    CK(cudaStreamSynchronize(ctx->stream));
    memset(out, 0, sizeof(*out));
    if(ctx->haveAs) cudaEventElapsedTime(&out->as_build_ms, ctx->ev[EV_AS0], ctx->ev[EV_AS1]);
    if(ctx->haveFrame) {
        cudaEventElapsedTime(&out->rt_total_ms, ctx->ev[EV_RT0], ctx->ev[EV_POST1]);
        cudaEventElapsedTime(&out->rt_only_ms, ctx->ev[EV_RT0], ctx->ev[EV_RTONLY1]);
        cudaEventElapsedTime(&out->rough_ms, ctx->ev[EV_ROUGH0], ctx->ev[EV_ROUGH1]);
        cudaEventElapsedTime(&out->postproc_ms, ctx->ev[EV_RTONLY1], ctx->ev[EV_POST1]);
        cudaEventElapsedTime(&out->gather_ms, ctx->ev[EV_POST1], ctx->ev[EV_GATHER1]);
        if(!ctx->debugPostOnly) cudaEventElapsedTime(&out->trace_kernel_ms, ctx->ev[EV_RT0], ctx->ev[EV_TRACE1]);
        out->trace_scheduler = (uint32_t)ctx->schedLast;
        unsigned long long c[16];
        CK(cudaMemcpy(c, ctx->dCounters, sizeof c, cudaMemcpyDeviceToHost));
        out->generic_hits = c[8];
        out->rays_primary = c[0]; out->rays_shadow = c[1]; out->rays_reflect = c[2]; out->rays_refract = c[3]; out->sky_lookups = c[4];
        out->nodes_visited = c[5]; out->tris_tested = c[6]; out->instances_entered = c[7];
        if(c[13] && getenv("RGB200_DEBUG_TAIL")) {   // written by a -DRG_DEBUG_TAIL build of the lanes kernel only
            const double span = (double)(c[10] - ~c[12]), busy = (double)c[11] / (double)c[13];
            fprintf(stderr, "[rgb200] trace warps %llu: first start -> last end %.3f ms, mean warp lifetime %.3f ms (%.1f %% of the span)\n", c[13], span * 1e-6, busy * 1e-6,
                    100.0 * busy / span);
        }
    }
    cudaGetLastError();
    return 0;
}

int rg_framebuffer_device_ptr(rg_ctx* ctx, void** d_ptr) {
    if(!ctx || !d_ptr) return 1;
    *d_ptr = ctx->rgba8;
    return 0;
}

// A pointer handed over inside ONE process may live on another GPU: enable peer access to its device (CUDA IPC mappings come with it).
static int ensurePeerAccess(rg_ctx* ctx, const void* p, const char* who) {
    if(!p) return 0;
    cudaPointerAttributes a{};
    if(cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return 0; }
    if(a.type != cudaMemoryTypeDevice || a.device == ctx->device) return 0;
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, ctx->device, a.device));
    if(!can) return fail(ctx, "%s: device %d cannot access memory of device %d", who, ctx->device, a.device);
    const cudaError_t e = cudaDeviceEnablePeerAccess(a.device, 0);
    if(e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return 0; }
    if(e != cudaSuccess) return fail(ctx, "%s: cudaDeviceEnablePeerAccess(%d) failed: %s", who, a.device, cudaGetErrorString(e));
    return 0;
}

int rg_set_gather_target(rg_ctx* ctx, void* d_target_rgba8) {
    if(!ctx) return 1;
    USE_DEVICE();
    if(ensurePeerAccess(ctx, d_target_rgba8, "rg_set_gather_target")) return 1;
    ctx->gatherTarget = (uint32_t*)d_target_rgba8;
    return 0;
}

int rg_host_frame_register(rg_ctx* ctx, void* host_ptr, size_t bytes, void** d_ptr) {
    if(!ctx || !host_ptr || !bytes || !d_ptr) return fail(ctx, "rg_host_frame_register: null");
    USE_DEVICE();
    const cudaError_t e = cudaHostRegister(host_ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
    if(e != cudaSuccess) { cudaGetLastError(); return fail(ctx, "rg_host_frame_register: cudaHostRegister failed: %s", cudaGetErrorString(e)); }
    void* d = nullptr;
    if(cudaHostGetDevicePointer(&d, host_ptr, 0) != cudaSuccess || !d) {
        cudaGetLastError(); cudaHostUnregister(host_ptr);
        return fail(ctx, "rg_host_frame_register: the registered memory has no device address");
    }
    *d_ptr = d;
    return 0;
}

int rg_host_frame_unregister(rg_ctx* ctx, void* host_ptr) {
    if(!ctx || !host_ptr) return 1;
    USE_DEVICE();
    void* d = nullptr;
    if(cudaHostGetDevicePointer(&d, host_ptr, 0) == cudaSuccess && ctx->gatherTarget == d) ctx->gatherTarget = nullptr;
    cudaGetLastError();
    CK(cudaStreamSynchronize(ctx->stream));   // no kernel of this context still stores into it
    CK(cudaHostUnregister(host_ptr));
    return 0;
}

int rg_gather_buffer_export(rg_ctx* ctx, void* handle64, void** d_ptr) {
    if(!ctx) return 1;
    USE_DEVICE();
    if(!ctx->gatherOwn) {
        CK(cudaMalloc(&ctx->gatherOwn, (size_t)ctx->W * ctx->H * 4));
        CK(cudaMemset(ctx->gatherOwn, 0, (size_t)ctx->W * ctx->H * 4));
    }
    if(handle64) {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        cudaIpcMemHandle_t h;
        CK(cudaIpcGetMemHandle(&h, ctx->gatherOwn));
        memcpy(handle64, &h, 64);
    }
    if(d_ptr) *d_ptr = ctx->gatherOwn;
    return 0;
}

int rg_gather_buffer_open(rg_ctx* ctx, const void* handle64, void** d_ptr) {
    if(!ctx || !handle64 || !d_ptr) return fail(ctx, "rg_gather_buffer_open: null");
    USE_DEVICE();
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    CK(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

int rg_gather_buffer_close(rg_ctx* ctx, void* d_ptr) {
    if(!ctx) return 1;
    USE_DEVICE();
    if(ctx->gatherTarget == d_ptr) ctx->gatherTarget = nullptr;
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaIpcCloseMemHandle(d_ptr));
    return 0;
}

int rg_read_gathered_rgba8(rg_ctx* ctx, void* dst) {
    if(!ctx || !dst) return fail(ctx, "rg_read_gathered_rgba8: null");
    if(!ctx->gatherOwn) return fail(ctx, "rg_read_gathered_rgba8: no gather buffer on this context");
    USE_DEVICE();
    CK(cudaMemcpyAsync(dst, ctx->gatherOwn, (size_t)ctx->W * ctx->H * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int rg_set_partition(rg_ctx* ctx, uint32_t rank, uint32_t world) {
    if(!ctx) return 1;
    if(world < 1 || world > (uint32_t)kMaxPeers || rank >= world) return fail(ctx, "rg_set_partition: rank %u of %u (max %d)", rank, world, kMaxPeers);
    USE_DEVICE();
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->rank = rank; ctx->world = world; ctx->frameId = 0;
    for(auto& a: ctx->peerAttached) a = false;
    CK(cudaMemset(ctx->arriveTrace, 0, 4 * kMaxPeers)); CK(cudaMemset(ctx->arrivePost, 0, 4 * kMaxPeers)); CK(cudaMemset(ctx->dSyncErr, 0, 4));
    // own slots: a rank also signals itself, so the flag-pointer tables contain the local arrays at [rank]
    uint32_t* tr[kMaxPeers]{}; uint32_t* po[kMaxPeers]{};
    tr[rank] = ctx->arriveTrace; po[rank] = ctx->arrivePost;
    for(auto& x: ctx->peerArriveTrace) x = nullptr;
    for(auto& x: ctx->peerArrivePost) x = nullptr;
    ctx->peerArriveTrace[rank] = ctx->arriveTrace; ctx->peerArrivePost[rank] = ctx->arrivePost;
    CK(cudaMemcpy(ctx->dPeerTraceFlags, tr, sizeof tr, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->dPeerPostFlags, po, sizeof po, cudaMemcpyHostToDevice));
    return 0;
}

int rg_peer_export(rg_ctx* ctx, rg_peer_desc* out) {
    if(!ctx || !out) return 1;
    USE_DEVICE();
    CK(cudaStreamSynchronize(ctx->stream));
    memset(out, 0, sizeof(*out));
    out->base = ctx->base; out->normal = ctx->normal; out->rough = ctx->rough; out->arrive_trace = ctx->arriveTrace; out->arrive_post = ctx->arrivePost;
    out->x0 = ctx->rx0; out->y0 = ctx->ry0; out->w = ctx->rw; out->h = ctx->rh;
    void* ptrs[5] = {ctx->base, ctx->normal, ctx->rough, ctx->arriveTrace, ctx->arrivePost};
    for(int k = 0; k < 5; ++k) {
        cudaIpcMemHandle_t h;
        CK(cudaIpcGetMemHandle(&h, ptrs[k]));
        memcpy(out->ipc[k], &h, 64);
    }
    return 0;
}

int rg_peer_attach(rg_ctx* ctx, uint32_t peer_rank, const rg_peer_desc* desc, int open_ipc) {
    if(!ctx || !desc) return 1;
    if(peer_rank >= ctx->world || peer_rank == ctx->rank) return fail(ctx, "rg_peer_attach: bad peer rank %u", peer_rank);
    USE_DEVICE();
    CK(cudaStreamSynchronize(ctx->stream));
    void* ptrs[5] = {desc->base, desc->normal, desc->rough, desc->arrive_trace, desc->arrive_post};
    if(open_ipc) {
        for(int k = 0; k < 5; ++k) {
            if(ctx->peerIpcOpened[peer_rank][k]) { cudaIpcCloseMemHandle(ctx->peerIpcOpened[peer_rank][k]); ctx->peerIpcOpened[peer_rank][k] = nullptr; }
            cudaIpcMemHandle_t h;
            memcpy(&h, desc->ipc[k], 64);
            CK(cudaIpcOpenMemHandle(&ptrs[k], h, cudaIpcMemLazyEnablePeerAccess));
            ctx->peerIpcOpened[peer_rank][k] = ptrs[k];
        }
    } else {
        for(int k = 0; k < 5; ++k) if(ensurePeerAccess(ctx, ptrs[k], "rg_peer_attach")) return 1;
    }
    ctx->peers[peer_rank] = TraceParams::Target{(uint2*)ptrs[0], (uint2*)ptrs[1], (uint2*)ptrs[2], desc->x0, desc->y0, desc->w, desc->h};
    ctx->peerArriveTrace[peer_rank] = (uint32_t*)ptrs[3]; ctx->peerArrivePost[peer_rank] = (uint32_t*)ptrs[4];
    CK(cudaMemcpy(ctx->dPeerTraceFlags, ctx->peerArriveTrace, sizeof(void*) * kMaxPeers, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->dPeerPostFlags, ctx->peerArrivePost, sizeof(void*) * kMaxPeers, cudaMemcpyHostToDevice));
    ctx->peerAttached[peer_rank] = true;
    return 0;
}

int rg_peer_detach_all(rg_ctx* ctx) {
    if(!ctx) return 1;
    USE_DEVICE();
    CK(cudaStreamSynchronize(ctx->stream));
    for(uint32_t q = 0; q < (uint32_t)kMaxPeers; ++q) {
        for(int k = 0; k < 5; ++k) if(ctx->peerIpcOpened[q][k]) { cudaIpcCloseMemHandle(ctx->peerIpcOpened[q][k]); ctx->peerIpcOpened[q][k] = nullptr; }
        ctx->peerAttached[q] = false;
    }
    return 0;
}

int rg_set_trace_scheduler(rg_ctx* ctx, int mode) {
    if(!ctx) return 1;
    if(mode != RG_SCHED_LANES && mode != RG_SCHED_POOL && mode != RG_SCHED_AUTO) return fail(ctx, "rg_set_trace_scheduler: mode must be RG_SCHED_LANES, RG_SCHED_POOL or RG_SCHED_AUTO");
    ctx->schedMode = mode; ctx->schedProbe = -1; ctx->schedMs[0] = ctx->schedMs[1] = -1.0f; ctx->schedFrames = 0;
    return 0;
}

int rg_sync_error(rg_ctx* ctx) {   // non-zero if a cross-GPU wait timed out (1 + rank that never arrived)
    if(!ctx) return -1;
    uint32_t e = 0;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaMemcpy(&e, ctx->dSyncErr, 4, cudaMemcpyDeviceToHost);
    return (int)e;
}

int rg_timer_begin(rg_ctx* ctx) {
    if(!ctx) return 1;
    USE_DEVICE();
    CK(cudaEventRecord(ctx->ev[EV_USER0], ctx->stream));
    return 0;
}

int rg_timer_end(rg_ctx* ctx, float* ms) {
    if(!ctx || !ms) return 1;
    USE_DEVICE();
    CK(cudaEventRecord(ctx->ev[EV_USER1], ctx->stream));
    CK(cudaEventSynchronize(ctx->ev[EV_USER1]));
    CK(cudaEventElapsedTime(ms, ctx->ev[EV_USER0], ctx->ev[EV_USER1]));
    return 0;
}

int rg_flush_l2(rg_ctx* ctx) {
    if(!ctx) return 1;
    USE_DEVICE();
    const size_t bytes = 256u << 20;  // > 126 MB L2
    if(!ctx->flushBuf) CK(cudaMalloc(&ctx->flushBuf, bytes));
    CK(cudaMemsetAsync(ctx->flushBuf, (int)(ctx->launches & 0xff), bytes, ctx->stream));
    return 0;
}

// ---- parity / introspection entry points ----------------------------------------------------------------
int rg_debug_blas_sort(rg_ctx* ctx, uint32_t mesh, uint32_t* keys_sorted, uint32_t* prim_order, uint32_t capacity) {
    if(!ctx) return 1;
    if(mesh >= ctx->meshes.size() || !ctx->meshes[mesh].built) return fail(ctx, "rg_debug_blas_sort: mesh %u has no BLAS", mesh);
    USE_DEVICE();
    MeshBlas& mb = ctx->meshes[mesh];
    const uint32_t n = mb.range.idx_cnt / 3;
    if(capacity < n) return fail(ctx, "rg_debug_blas_sort: capacity %u < %u", capacity, n);
    if(n) {
        CK(cudaMemcpy(keys_sorted, mb.scratch.keys[mb.scratch.sortedBuf], 4 * (size_t)n, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(prim_order, mb.scratch.vals[mb.scratch.sortedBuf], 4 * (size_t)n, cudaMemcpyDeviceToHost));
    }
    return 0;
}

int rg_debug_tlas_sort(rg_ctx* ctx, uint32_t* keys_sorted, uint32_t* inst_order, uint32_t capacity) {
    if(!ctx) return 1;
    USE_DEVICE();
    const uint32_t n = ctx->nInst;
    if(capacity < n) return fail(ctx, "rg_debug_tlas_sort: capacity %u < %u", capacity, n);
    CK(cudaStreamSynchronize(ctx->stream));
    if(n) {
        CK(cudaMemcpy(keys_sorted, ctx->tlasScratch.keys[ctx->tlasScratch.sortedBuf], 4 * (size_t)n, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(inst_order, ctx->tlasScratch.vals[ctx->tlasScratch.sortedBuf], 4 * (size_t)n, cudaMemcpyDeviceToHost));
    }
    return 0;
}

int rg_debug_trace_rays(rg_ctx* ctx, const float* rays8, uint32_t n, float* tuv, uint32_t* inst_prim) {
    if(!ctx) return 1;
    if(!ctx->haveAs) return fail(ctx, "rg_debug_trace_rays: no acceleration structure");
    USE_DEVICE();
    float *dR = nullptr, *dT = nullptr; uint32_t* dI = nullptr;
    CK(cudaMalloc(&dR, 32 * (size_t)(n + 1))); CK(cudaMalloc(&dT, 12 * (size_t)(n + 1))); CK(cudaMalloc(&dI, 8 * (size_t)(n + 1)));
    CK(cudaMemcpyAsync(dR, rays8, 32 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    TraceParams tp; fillTraceParams(ctx, tp, 0);
    launchTraceRays(tp, dR, n, dT, dI, ctx->stream);   // warm-up
    CK(cudaEventRecord(ctx->ev[EV_USER0], ctx->stream));
    launchTraceRays(tp, dR, n, dT, dI, ctx->stream);
    CK(cudaEventRecord(ctx->ev[EV_USER1], ctx->stream));
    ctx->launches += 2;
    CK(cudaMemcpyAsync(tuv, dT, 12 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(inst_prim, dI, 8 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(dR); cudaFree(dT); cudaFree(dI);
    CK(cudaGetLastError());
    cudaEventElapsedTime(&ctx->lastRaysMs, ctx->ev[EV_USER0], ctx->ev[EV_USER1]);
    return 0;
}

float rg_debug_last_trace_rays_ms(const rg_ctx* ctx) { return ctx ? ctx->lastRaysMs : 0.0f; }

int rg_debug_bvh_stats(rg_ctx* ctx, uint64_t* out8) {
    if(!ctx || !out8) return 1;
    out8[0] = ctx->nodesUsed; out8[1] = ctx->trisUsed;
    out8[2] = (uint64_t)ctx->nodesUsed * sizeof(Node8) + (uint64_t)ctx->trisUsed * sizeof(Tri);
    uint32_t c[4] = {0, 0, 0, 0};
    if(ctx->tlasScratch.counters) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); cudaMemcpy(c, ctx->tlasScratch.counters, sizeof c, cudaMemcpyDeviceToHost); }
    out8[3] = ctx->nInst ? c[2] : 0; out8[4] = ctx->nInst ? c[3] : 0;
    out8[5] = (uint64_t)out8[3] * sizeof(Node8) + (uint64_t)out8[4] * sizeof(InstTrav);
    out8[6] = ctx->meshes.size(); out8[7] = ctx->nInst;
    return 0;
}

// Parity helpers for the post chain: replace the G-buffer with caller data (region-sized, tightly packed) and run
// rough_prepare .. fxaa + blit on it.  Only meaningful when no region split is active.
int rg_debug_upload_gbuffer(rg_ctx* ctx, const void* base, const void* normal, const void* rough) {
    if(!ctx || !base || !normal || !rough) return fail(ctx, "rg_debug_upload_gbuffer: null");
    if(ctx->rw != (int)ctx->W || ctx->rh != (int)ctx->H) return fail(ctx, "rg_debug_upload_gbuffer: needs the full-frame region");
    USE_DEVICE();
    const size_t n = (size_t)ctx->W * ctx->H * 8;
    CK(cudaMemcpyAsync(ctx->base, base, n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->normal, normal, n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->rough, rough, n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int rg_debug_run_post(rg_ctx* ctx, uint32_t flags) {
    if(!ctx) return 1;
    USE_DEVICE();
    CK(cudaEventRecord(ctx->ev[EV_RT0], ctx->stream));
    CK(cudaEventRecord(ctx->ev[EV_RTONLY1], ctx->stream));
    ctx->debugPostOnly = true;
    if(runPost(ctx, flags)) return 1;
    ctx->haveFrame = true; ctx->lastFlags = flags;
    return 0;
}

}  // extern "C"
