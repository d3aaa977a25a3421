/* rgb200.h -- C ABI of librgb200.so: the B200-native replacement for Raygun's per-frame
 * ray-tracing path (TLAS/BLAS build -> raygen / closest-hit / miss recursion -> five compute
 * post passes -> 8-bit frame).  Plain pointers and sizes only; no CUDA / torch types.
 *
 * The reference has no plugin or FFI layer: the seam is the C++ class raygun::render::Raytracer
 * plus four POD layouts shared between C++ and GLSL.  Every entry point below names the
 * reference interface it replaces (paths relative to the reference tree).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; rg_last_error() gives the text.
 *     Nothing throws across the boundary (the reference aborts via RAYGUN_FATAL, logging.hpp:35-40).
 *   - uploads copy; host pointers are never retained (the reference keeps host-visible, persistently
 *     re-mapped buffers, gpu/gpu_buffer.cpp:60-75).
 *   - one rg_ctx is used from one host thread (the reference render path is single-threaded:
 *     render/render_system.cpp:332-361).  Work is asynchronous on the context's CUDA stream until
 *     rg_sync / rg_read_* / rg_get_timings.
 *   - there is NO CPU fallback: without a CUDA device rg_create fails.
 */
#ifndef RGB200_H
#define RGB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rg_ctx rg_ctx;

/* resources/shaders/vertex.def:3-7 == raygun/render/vertex.hpp:27-29 (32 bytes) */
typedef struct rg_vertex {
    float position[3];
    uint32_t mat_index;
    float normal[3];
    float pad1;
} rg_vertex;

/* resources/shaders/gpu_material.def:11-26 == raygun/gpu/gpu_material.hpp (64 bytes) */
typedef struct rg_material {
    float diffuse[3];
    float transparency;
    float specular[3];
    float reflectivity;
    float roughness;
    float ior;
    uint32_t effect_id;
    uint32_t ray_consumption;
    float emission;
    float pad0, pad1, pad2;
} rg_material;

/* resources/shaders/uniform_buffer_object.def:3-17 == raygun/gpu/uniform_buffer.hpp (192 bytes).
 * Matrices are column-major (GLM).  show_alpha is read as "first byte non-zero" (C++ bool). */
typedef struct rg_ubo {
    float view_inverse[16];
    float proj_inverse[16];
    float clear_color[3];
    int32_t num_samples;
    float light_dir[3];
    int32_t max_recursions;
    float time;
    uint32_t show_alpha;
    float pad0, pad1;
    float fade_color[4];
} rg_ubo;

/* BufferRef offsets of one mesh inside the shared vertex / index buffers, in ELEMENTS
 * (raygun/gpu/gpu_buffer.hpp:67-76, raygun/render/render_system.cpp:270-305). */
typedef struct rg_mesh_range {
    uint32_t vtx_off, vtx_cnt, idx_off, idx_cnt;
} rg_mesh_range;

/* One TLAS instance: VkAccelerationStructureInstanceKHR as filled by instanceFromEntity
 * (raygun/render/acceleration_structure.cpp:34-52: 3x4 row-major object->world, mask 0xff,
 * TriangleCullDisable, customIndex = position in the array) fused with its InstanceOffsetTableEntry
 * (resources/shaders/instance_offset_table.def:1-3, acceleration_structure.cpp:77-82). 64 bytes. */
typedef struct rg_instance {
    float xform[12];
    uint32_t mesh;    /* which rg_mesh_range / BLAS */
    uint32_t vtx_off; /* vertexBufferOffset   */
    uint32_t idx_off; /* indexBufferOffset    */
    uint32_t mat_off; /* materialBufferOffset */
} rg_instance;

/* One node of the scene graph for the device-side walk (rg_set_entities): what TopLevelAS::TopLevelAS reads of an Entity
 * (raygun/render/acceleration_structure.cpp:55-85, raygun/entity.hpp:32-131): the LOCAL transform as TRS (raygun/transform.hpp:
 * 108-112), visibility, and the model's BufferRef offsets.  Entities are listed in the order Entity::forEachEntity visits them
 * (DFS pre-order, entity.hpp:67-84), so a parent always precedes its children.  64 bytes. */
typedef struct rg_entity {
    float position[3];
    int32_t parent;      /* index of the parent entity, -1 for the root */
    float rotation[4];   /* glm::quat as w, x, y, z */
    float scaling[3];
    uint32_t flags;      /* RG_ENTITY_VISIBLE | RG_ENTITY_HAS_MODEL */
    uint32_t mesh, vtx_off, idx_off, mat_off;   /* as in rg_instance; ignored without RG_ENTITY_HAS_MODEL */
} rg_entity;
#define RG_ENTITY_VISIBLE 1u
#define RG_ENTITY_HAS_MODEL 2u

/* The five GPU sections of the reference profiler (raygun/profiler.def:6-10), from CUDA events,
 * plus the NVLink gather and device-side ray counters.  A "ray" is one traceRayEXT with a non-zero
 * cull mask; sky look-ups (cull mask 0, closesthit.rchit:144) are counted separately. */
typedef struct rg_timings {
    float as_build_ms;  /* ASBuild  : raytracer.cpp:78-84  */
    float rt_total_ms;  /* RTTotal  : raytracer.cpp:93,144 */
    float rt_only_ms;   /* RTOnly   : raytracer.cpp:95,104 */
    float rough_ms;     /* Rough    : raytracer.cpp:111,123 */
    float postproc_ms;  /* Postproc : raytracer.cpp:106,142 */
    float gather_ms;    /* tile gather to the target GPU (no reference counterpart) */
    uint64_t rays_primary, rays_shadow, rays_reflect, rays_refract, sky_lookups;
    uint64_t nodes_visited, tris_tested, instances_entered, generic_hits; /* only with RG_COUNT_TRAVERSAL */
    float trace_kernel_ms; /* the trace kernel alone (rt_only_ms additionally holds the cross-GPU barrier in partitioned mode) */
    uint32_t trace_scheduler; /* RG_SCHED_LANES or RG_SCHED_POOL: the trace kernel the last frame used */
} rg_timings;

/* rg_render flags */
#define RG_FXAA 1u            /* Raytracer::m_useFXAA (raytracer.cpp:136) */
#define RG_SRGB8 2u           /* 8-bit target is an SRGB format (vulkan_context.cpp:180) */
#define RG_STRICT_IEEE 4u     /* keep the 0/0 of closesthit.rchit:257 (SURVEY hazard 8); default NaN-free */
#define RG_DEBUG_IDS 8u       /* also write primary-hit instance / primitive ids */
#define RG_COUNT_TRAVERSAL 16u/* instrumented traversal counters (slower) */
#define RG_NO_GATHER 32u      /* skip the peer gather even if a target is set */

/* rg_read_image selectors; order mirrors the ImGui "Image" combo (raytracer.cpp:379-392) */
enum { RG_IMG_FINAL = 0, RG_IMG_BASE = 1, RG_IMG_NORMAL = 2, RG_IMG_ROUGH = 3, RG_IMG_TRANSITIONS = 4, RG_IMG_ROUGH_A = 5, RG_IMG_ROUGH_B = 6 };

/* Raytracer::Raytracer() (raytracer.cpp:38-52): images are sized from the window size. */
int rg_create(rg_ctx** out, int cuda_device, uint32_t width, uint32_t height);
void rg_destroy(rg_ctx* ctx);
/* RenderSystem::reload() -> new Raytracer (render_system.cpp:77-78) */
int rg_resize(rg_ctx* ctx, uint32_t width, uint32_t height);
const char* rg_last_error(const rg_ctx* ctx);

/* Multi-GPU screen split (no reference counterpart; the reference is single-device,
 * vulkan_context.cpp:165).  The context renders only pixels [x0,x1) x [y0,y1) of the full
 * width x height frame (plus an internal halo so the post chain is bit-identical to a
 * single-GPU frame).  Default region = whole frame. */
int rg_set_region(rg_ctx* ctx, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1);

/* RenderSystem::setupModelBuffers / updateModelBuffers (render_system.cpp:192-233, :270-330) */
int rg_upload_geometry(rg_ctx* ctx, const rg_vertex* vertices, uint32_t n_vertices, const uint32_t* indices, uint32_t n_indices,
                       const rg_mesh_range* meshes, uint32_t n_meshes);
int rg_upload_materials(rg_ctx* ctx, const rg_material* materials, uint32_t n_materials);

/* Raytracer::setupBottomLevelAS (raytracer.cpp:54-74; BottomLevelAS acceleration_structure.cpp:140-194):
 * builds the BLAS of every uploaded mesh; synchronous like the reference's fence wait. */
int rg_build_blas(rg_ctx* ctx);
/* Extension (BASELINE config 4; the reference never refits): new vertex records for one mesh
 * (same count), topology kept, boxes refitted bottom-up and re-quantised. */
int rg_refit_blas(rg_ctx* ctx, uint32_t mesh, const rg_vertex* new_vertices);
/* The same with the new vertex records already in DEVICE memory (written by a skinning / physics kernel of the caller on any stream
 * it has synchronised with): device-to-device copy on the library's stream, then the refit; no host round trip. */
int rg_refit_blas_device(rg_ctx* ctx, uint32_t mesh, const rg_vertex* d_new_vertices);

/* Raytracer::setupTopLevelAS (raytracer.cpp:76-85; TopLevelAS acceleration_structure.cpp:55-138):
 * called every frame; rebuilds the TLAS from scratch on the device.  The caller performs the entity
 * DFS of acceleration_structure.cpp:63-85 (raygun_b200/host does). */
int rg_set_instances(rg_ctx* ctx, const rg_instance* instances, uint32_t n_instances);

/* The same step with the scene-graph walk on the device: globalTransform = parent's global * local as TRS
 * structs (raygun/entity.cpp:187-199, transform.hpp:99-106), subtrees of invisible or zero-volume entities pruned
 * (acceleration_structure.cpp:65-67), instance = transpose(toMat4()) as 3x4 (:44-45), instance order = DFS order.  The records are
 * bit-identical to the host walk.  rg_set_entities copies from host memory; rg_set_entities_device reads device memory (for
 * simulations that live on the GPU).  Returns the number of instances through n_instances_out (may be NULL). */
int rg_set_entities(rg_ctx* ctx, const rg_entity* entities, uint32_t n_entities, uint32_t* n_instances_out);
int rg_set_entities_device(rg_ctx* ctx, const rg_entity* d_entities, uint32_t n_entities, uint32_t* n_instances_out);

/* Stand-in for PhysicsSystem::update (raygun/physics/physics_system.cpp:241-258): rigid spheres under gravity
 * (0, -9.81, 0) (physics_system.cpp:68) over the plane y = floor_y, semi-implicit Euler, no sphere-sphere contacts.  Both arrays
 * live in DEVICE memory and are updated in place: body i drives entity i (radius <= 0: no dynamic actor); the pose is written as
 * the entity's local transform, as the reference does for actors under an identity parent.  Follow with rg_set_entities_device:
 * an animated frame then needs no host data at all. */
typedef struct rg_sphere_body {
    float velocity[3];
    float radius;
    float angular_velocity[3];
    float restitution;   /* default material: 0.6 (physics_system.cpp:40) */
} rg_sphere_body;
int rg_physics_step_spheres(rg_ctx* ctx, rg_entity* d_entities, rg_sphere_body* d_bodies, uint32_t n, float dt, float floor_y);

/* RenderSystem::updateUniformBuffer + Raytracer::updateRenderTarget (render_system.cpp:246-268, raytracer.cpp:149-171) */
int rg_set_ubo(rg_ctx* ctx, const rg_ubo* ubo);

/* Raytracer::doRaytracing (raytracer.cpp:87-147) + the blit to the 8-bit target (render_system.cpp:130-144). */
int rg_render(rg_ctx* ctx, uint32_t flags);
int rg_sync(rg_ctx* ctx);

/* Scheduler of the trace kernel (no reference counterpart: traceRaysKHR leaves scheduling to the driver).  LANES: one pixel sample
 * per lane, fastest on coherent scenes.  POOL: per-warp context pools with shared-memory ray / hit queues, fastest on incoherent
 * bounces.  AUTO (default): times both on consecutive frames, keeps the faster, re-checks every 256 frames.  Images are
 * bit-identical either way. */
#define RG_SCHED_LANES 0
#define RG_SCHED_POOL 1
#define RG_SCHED_AUTO 2
int rg_set_trace_scheduler(rg_ctx* ctx, int mode);

/* Read-back.  dst is tightly packed width x height (full frame when no region is set, else the region).
 * rg_read_rgba8: 4 bytes / pixel.  rg_read_image: 8 bytes / pixel (4 x binary16), 1 byte for RG_IMG_TRANSITIONS. */
int rg_read_rgba8(rg_ctx* ctx, void* dst);
int rg_read_image(rg_ctx* ctx, int which, void* dst);
int rg_read_ids(rg_ctx* ctx, uint32_t* instance_ids, uint32_t* primitive_ids);
/* Profiler::getTimeRangeMS (profiler.cpp:50-53) for the last rendered frame. */
int rg_get_timings(rg_ctx* ctx, rg_timings* out);

/* Device-resident variants (no host copies): pointers are CUDA device pointers on the context's GPU. */
int rg_set_instances_device(rg_ctx* ctx, const rg_instance* d_instances, uint32_t n_instances);
int rg_set_ubo_device(rg_ctx* ctx, const rg_ubo* d_ubo);
/* Device address of the RGBA8 frame buffer (region-sized, row pitch = region width * 4). */
int rg_framebuffer_device_ptr(rg_ctx* ctx, void** d_ptr);

/* Partitioned multi-GPU mode (preferred over plain regions: no halo is traced twice and the load is balanced).
 * Every rank keeps its region (rg_set_region) for the post chain, but TRACES every world-th chunk of 8x4-pixel tiles of the
 * whole frame; the trace kernel stores each finished pixel's G-buffer straight into the images of every rank whose
 * region + halo contains it (own memory or peer memory over NVLink).  Two device-side flag barriers per frame (after the
 * trace, after the post chain) keep the ranks in step without host round trips; a wait gives up after 4 s (rg_sync_error).
 * rg_set_region / rg_resize re-allocate the images a peer stores into, so they FAIL while the context is partitioned (world > 1):
 * leave the mode on every rank first (rg_peer_detach_all, rg_set_partition(ctx, 0, 1)), resize, then attach again.
 * Order of calls: rg_set_region, rg_set_partition, then on every rank rg_peer_export -> exchange the descriptors ->
 * rg_peer_attach for every other rank (open_ipc = 1 across processes, 0 for contexts of one process). */
typedef struct rg_peer_desc {
    void* base; void* normal; void* rough;          /* G-buffer images of the exporting rank (device pointers in ITS address space) */
    void* arrive_trace; void* arrive_post;          /* its two flag arrays */
    int32_t x0, y0, w, h;                           /* its rectangle (region + halo) in frame coordinates */
    uint8_t ipc[5][64];                             /* cudaIpcMemHandle_t of the five allocations above */
} rg_peer_desc;
int rg_set_partition(rg_ctx* ctx, uint32_t rank, uint32_t world);
int rg_peer_export(rg_ctx* ctx, rg_peer_desc* out);
int rg_peer_attach(rg_ctx* ctx, uint32_t peer_rank, const rg_peer_desc* desc, int open_ipc);
int rg_peer_detach_all(rg_ctx* ctx);
int rg_sync_error(rg_ctx* ctx);

/* Tile gather over NVLink: the final kernel (FXAA + 8-bit convert) stores this context's region
 * straight into `d_target` (a width x height RGBA8 frame that may live on a PEER GPU: either a
 * pointer in the same process with peer access enabled, or one opened from a CUDA IPC handle).
 * Pinned host memory works as well (unified addressing): the frame then reaches the host as the
 * kernel's own stores over PCIe, without a separate copy -- valid after rg_sync (bench.py, e2e). */
int rg_set_gather_target(rg_ctx* ctx, void* d_target_rgba8);   /* rg_resize clears it: the buffer has the old frame's stride */
/* Page-lock and map `bytes` of ordinary host memory (page aligned: a POSIX shared-memory mapping, a
 * malloc'ed frame) so that it can serve as a gather target; *d_ptr is the address to pass to
 * rg_set_gather_target.  With one mapping of the same shared-memory frame per process, every GPU of
 * a node stores its band of the frame into host memory over its own PCIe link and the consumer reads
 * the assembled frame without any copy (bench.py, e2e at N > 1).  No reference counterpart (the
 * reference presents through the swapchain, render_system.cpp:130-159). */
int rg_host_frame_register(rg_ctx* ctx, void* host_ptr, size_t bytes, void** d_ptr);
int rg_host_frame_unregister(rg_ctx* ctx, void* host_ptr);
/* 64-byte cudaIpcMemHandle_t of this context's full-frame gather buffer (allocated on demand) and
 * its opening on another process' context. */
int rg_gather_buffer_export(rg_ctx* ctx, void* handle64, void** d_ptr);
int rg_gather_buffer_open(rg_ctx* ctx, const void* handle64, void** d_ptr);
int rg_gather_buffer_close(rg_ctx* ctx, void* d_ptr);
/* Read the full-frame gather buffer of this context (GPU 0) to the host. */
int rg_read_gathered_rgba8(rg_ctx* ctx, void* dst);

/* Builder introspection for the parity tests (Morton / sort output must be bit-exact):
 * sorted Morton keys and the primitive order of one mesh's BLAS, and of the current TLAS. */
int rg_debug_blas_sort(rg_ctx* ctx, uint32_t mesh, uint32_t* keys_sorted, uint32_t* prim_order, uint32_t capacity);
/* The instance records of the current TLAS in instance order (what rg_set_instances received or rg_set_entities composed). */
int rg_debug_read_instances(rg_ctx* ctx, rg_instance* out, uint32_t capacity, uint32_t* n_out);
int rg_debug_tlas_sort(rg_ctx* ctx, uint32_t* keys_sorted, uint32_t* inst_order, uint32_t capacity);
/* Closest-hit queries straight into the traversal kernel (n rays: org xyz, dir xyz, tmin, tmax = 8 floats each);
 * out: t,u,v (3 floats) + inst, prim (2 u32) per ray; inst = 0xffffffff on miss. */
int rg_debug_trace_rays(rg_ctx* ctx, const float* rays8, uint32_t n, float* tuv, uint32_t* inst_prim);
/* Device time of the last rg_debug_trace_rays kernel (second of two identical launches), milliseconds. */
float rg_debug_last_trace_rays_ms(const rg_ctx* ctx);
/* Counts: wide nodes, triangles, bytes of the BLAS set and of the current TLAS. */
int rg_debug_bvh_stats(rg_ctx* ctx, uint64_t* out8);

/* Post-chain parity helpers: replace the G-buffer (base, normal, rough: width x height x 4 binary16 each) with caller
 * data and run rough_prepare .. fxaa + blit on it.  Full-frame region only. */
int rg_debug_upload_gbuffer(rg_ctx* ctx, const void* base, const void* normal, const void* rough);
int rg_debug_run_post(rg_ctx* ctx, uint32_t flags);

/* Device-side stopwatch on the context's stream (CUDA events), and an L2 flush (writes a 256 MiB scratch buffer)
 * for benchmark hygiene.  rg_timer_end synchronises and returns the elapsed milliseconds since rg_timer_begin. */
int rg_timer_begin(rg_ctx* ctx);
int rg_timer_end(rg_ctx* ctx, float* ms);
int rg_flush_l2(rg_ctx* ctx);

/* Number of kernels this library launched on the context since creation (bench: gpu_launches). */
uint64_t rg_launch_count(const rg_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* RGB200_H */
