// rg_scene.cu -- the scene-graph walk that feeds the per-frame TLAS, on the device.
//
// Replaces the host loop of TopLevelAS::TopLevelAS + instanceFromEntity (raygun/render/acceleration_structure.cpp:34-85)
// together with Entity::globalTransform (raygun/entity.cpp:187-199) and Transform's operator* / toMat4
// (raygun/transform.hpp:38-46, 99-106) for scenes where that loop dominates the host side of a frame (BASELINE config 4:
// 10 000 animated entities):
//   * entities arrive as a flat array in DFS pre-order (the order Entity::forEachEntity visits them, entity.hpp:67-84) with the
//     index of their parent, their LOCAL position / rotation / scale, visibility and model references;
//   * a subtree is pruned when its root is invisible (acceleration_structure.cpp:65) or has zero volume (:67, transform.hpp:65);
//   * globalTransform composes TRS structs top-down, G(e) = G(parent) * L(e), never matrices;
//   * instance i (in DFS order among the entities that survive and have a model) gets transpose(G.toMat4()) as 3x4 row-major.
// Every float operation is an explicitly rounded single operation in GLM's order, so the records are BIT-IDENTICAL to what the
// host path (raygun_b200/host/raygun_host.cpp: Raytracer::gatherInstances) produces (tests/test_gpu_parity.py).
#include "rg_scene.cuh"

namespace rg {

namespace {

#define FM(a, b) __fmul_rn((a), (b))
#define FA(a, b) __fadd_rn((a), (b))
#define FS(a, b) __fsub_rn((a), (b))

struct Trs { float p[3]; float q[4] /* w x y z */; float s[3]; };

__device__ __forceinline__ void cross3(const float a[3], const float b[3], float r[3]) {   // glm::cross
    r[0] = FS(FM(a[1], b[2]), FM(b[1], a[2]));
    r[1] = FS(FM(a[2], b[0]), FM(b[2], a[0]));
    r[2] = FS(FM(a[0], b[1]), FM(b[0], a[1]));
}
// glm::rotate(quat, vec3) == q * v:  v + ((uv * q.w) + uuv) * 2
__device__ __forceinline__ void rotateQ(const float q[4], const float v[3], float r[3]) {
    const float qv[3] = {q[1], q[2], q[3]};
    float uv[3], uuv[3];
    cross3(qv, v, uv);
    cross3(qv, uv, uuv);
    for(int k = 0; k < 3; ++k) r[k] = FA(v[k], FM(FA(FM(uv[k], q[0]), uuv[k]), 2.0f));
}
// Transform operator* (transform.hpp:99-106): x = parent (global), y = local
__device__ __forceinline__ Trs compose(const Trs& x, const Trs& y) {
    Trs r;
    const float sp[3] = {FM(x.s[0], y.p[0]), FM(x.s[1], y.p[1]), FM(x.s[2], y.p[2])};
    float rp[3];
    rotateQ(x.q, sp, rp);
    for(int k = 0; k < 3; ++k) r.p[k] = FA(rp[k], x.p[k]);
    const float* p = x.q; const float* q = y.q;   // glm quat product, w x y z
    r.q[0] = FS(FS(FS(FM(p[0], q[0]), FM(p[1], q[1])), FM(p[2], q[2])), FM(p[3], q[3]));
    r.q[1] = FS(FA(FA(FM(p[0], q[1]), FM(p[1], q[0])), FM(p[2], q[3])), FM(p[3], q[2]));
    r.q[2] = FS(FA(FA(FM(p[0], q[2]), FM(p[2], q[0])), FM(p[3], q[1])), FM(p[1], q[3]));
    r.q[3] = FS(FA(FA(FM(p[0], q[3]), FM(p[3], q[0])), FM(p[1], q[2])), FM(p[2], q[1]));
    for(int k = 0; k < 3; ++k) r.s[k] = FM(x.s[k], y.s[k]);
    return r;
}
// column-major 4x4 product in GLM's order: r[c][row] = ((a[0][row]*b[c][0] + a[1][row]*b[c][1]) + a[2][row]*b[c][2]) + a[3][row]*b[c][3]
__device__ __forceinline__ void mul44(const float a[4][4], const float b[4][4], float r[4][4]) {
    for(int c = 0; c < 4; ++c)
        for(int row = 0; row < 4; ++row)
            r[c][row] = FA(FA(FA(FM(a[0][row], b[c][0]), FM(a[1][row], b[c][1])), FM(a[2][row], b[c][2])), FM(a[3][row], b[c][3]));
}
// Transform::toMat4 = translate(p) * mat4_cast(q) * scale(s) (transform.hpp:38-46), then transposed into 3x4 row-major
__device__ __forceinline__ void toInstanceXform(const Trs& t, float out[12]) {
    float T[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {t.p[0], t.p[1], t.p[2], 1}};
    float S[4][4] = {{t.s[0], 0, 0, 0}, {0, t.s[1], 0, 0}, {0, 0, t.s[2], 0}, {0, 0, 0, 1}};
    float R[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    const float w = t.q[0], x = t.q[1], y = t.q[2], z = t.q[3];
    const float qxx = FM(x, x), qyy = FM(y, y), qzz = FM(z, z), qxz = FM(x, z), qxy = FM(x, y), qyz = FM(y, z), qwx = FM(w, x), qwy = FM(w, y), qwz = FM(w, z);
    R[0][0] = FS(1.0f, FM(2.0f, FA(qyy, qzz))); R[0][1] = FM(2.0f, FA(qxy, qwz)); R[0][2] = FM(2.0f, FS(qxz, qwy));
    R[1][0] = FM(2.0f, FS(qxy, qwz)); R[1][1] = FS(1.0f, FM(2.0f, FA(qxx, qzz))); R[1][2] = FM(2.0f, FA(qyz, qwx));
    R[2][0] = FM(2.0f, FA(qxz, qwy)); R[2][1] = FM(2.0f, FS(qyz, qwx)); R[2][2] = FS(1.0f, FM(2.0f, FA(qxx, qyy)));
    float TR[4][4], M[4][4];
    mul44(T, R, TR);
    mul44(TR, S, M);
    // transpose(M) as rows: row r of the 3x4 = (M[0][r], M[1][r], M[2][r], M[3][r])
    for(int r = 0; r < 3; ++r)
        for(int c = 0; c < 4; ++c) out[r * 4 + c] = M[c][r];
}

__device__ __forceinline__ Trs localOf(const rg_entity& e) {
    Trs t;
    for(int k = 0; k < 3; ++k) { t.p[k] = e.position[k]; t.s[k] = e.scaling[k]; }
    for(int k = 0; k < 4; ++k) t.q[k] = e.rotation[k];
    return t;
}

// One thread per entity: walk up to the root (pruning test on the way), then fold the local transforms root-first.
__global__ void k_entity_instances(const rg_entity* __restrict__ ents, uint32_t n, rg_instance* __restrict__ tmp, uint32_t* __restrict__ emit) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    uint32_t chain[kMaxEntityDepth];
    int depth = 0;
    bool alive = true;
    uint32_t cur = i;
    while(true) {
        const rg_entity& e = ents[cur];
        // Entity::isVisible (acceleration_structure.cpp:65); Transform::isZeroVolume on the LOCAL transform (:67, transform.hpp:65)
        if(!(e.flags & RG_ENTITY_VISIBLE) || FM(FM(e.scaling[0], e.scaling[1]), e.scaling[2]) == 0.0f) { alive = false; break; }
        if(depth == kMaxEntityDepth) { alive = false; break; }   // deeper than the library supports: dropped (rg_set_entities rejects such input from host memory)
        chain[depth++] = cur;
        if(e.parent < 0 || (uint32_t)e.parent >= cur) break;     // the root (parents precede their children in DFS pre-order)
        cur = (uint32_t)e.parent;
    }
    const rg_entity& self = ents[i];
    const bool out = alive && (self.flags & RG_ENTITY_HAS_MODEL);
    emit[i] = out ? 1u : 0u;
    if(!out) return;
    Trs g = localOf(ents[chain[depth - 1]]);
    for(int k = depth - 2; k >= 0; --k) g = compose(g, localOf(ents[chain[k]]));   // G(e) = G(parent) * L(e), entity.cpp:196-199
    rg_instance in;
    toInstanceXform(g, in.xform);
    in.mesh = self.mesh; in.vtx_off = self.vtx_off; in.idx_off = self.idx_off; in.mat_off = self.mat_off;
    tmp[i] = in;
}

// Stable compaction (instance order = DFS order of the surviving entities): one block scans the emit flags chunk by chunk with a
// running carry and moves the 64-byte records.
__global__ void __launch_bounds__(1024) k_entity_compact(const rg_instance* __restrict__ tmp, const uint32_t* __restrict__ emit, uint32_t n,
                                                         rg_instance* __restrict__ out, uint32_t* __restrict__ count) {
    __shared__ uint32_t sWarp[32];
    __shared__ uint32_t sCarry;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if(tid == 0) sCarry = 0;
    __syncthreads();
    for(uint32_t base = 0; base < n; base += 1024u) {
        const uint32_t i = base + tid;
        const uint32_t f = i < n ? emit[i] : 0u;
        const uint32_t ballot = __ballot_sync(0xffffffffu, f != 0u);
        const uint32_t inWarp = __popc(ballot & ((1u << lane) - 1u));
        if(lane == 0) sWarp[warp] = __popc(ballot);
        __syncthreads();
        uint32_t before = 0;
        for(uint32_t w = 0; w < warp; ++w) before += sWarp[w];
        const uint32_t carry = sCarry;
        if(f) {
            const uint4* src = reinterpret_cast<const uint4*>(tmp + i);
            uint4* dst = reinterpret_cast<uint4*>(out + (carry + before + inWarp));
            dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
        }
        __syncthreads();
        if(tid == 1023) sCarry = carry + before + inWarp + (f ? 1u : 0u);
        __syncthreads();
    }
    if(tid == 0) *count = sCarry;
}

// Stand-in for PhysicsSystem::update (raygun/physics/physics_system.cpp:241-258; PhysX itself is out of scope): the contract of that
// function is "after the step, write each dynamic actor's pose into its entity's transform".  Rigid spheres under gravity
// (physics_system.cpp:68: (0, -9.81, 0)) over a floor plane, semi-implicit Euler, restitution per body (default material 0.6,
// physics_system.cpp:40); orientation integrated from the angular velocity.  No sphere-sphere contacts.  Bodies must hang directly
// under an identity-transform parent (the pose is written as the LOCAL transform).  Explicitly rounded operations in a fixed order:
// the oracle's numpy restatement (oracle/oracle.py: step_spheres) reproduces the state bit for bit.
__global__ void k_step_spheres(rg_entity* __restrict__ ents, rg_sphere_body* __restrict__ bodies, uint32_t n, float dt, float floorY) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    rg_sphere_body b = bodies[i];
    if(!(b.radius > 0.0f)) return;   // no dynamic actor
    rg_entity e = ents[i];
    b.velocity[1] = FA(b.velocity[1], FM(-9.81f, dt));
    for(int k = 0; k < 3; ++k) e.position[k] = FA(e.position[k], FM(b.velocity[k], dt));
    const float rest = FA(floorY, b.radius);
    if(e.position[1] < rest && b.velocity[1] < 0.0f) {
        e.position[1] = FA(rest, FM(FS(rest, e.position[1]), b.restitution));   // the part of the step below the floor comes back up
        b.velocity[1] = FM(-b.velocity[1], b.restitution);
    }
    // q += dt/2 * (0, w) * q, then normalise
    const float h = FM(0.5f, dt);
    const float wx = FM(b.angular_velocity[0], h), wy = FM(b.angular_velocity[1], h), wz = FM(b.angular_velocity[2], h);
    const float qw = e.rotation[0], qx = e.rotation[1], qy = e.rotation[2], qz = e.rotation[3];
    float nw = FA(qw, FS(FS(FM(-wx, qx), FM(wy, qy)), FM(wz, qz)));
    float nx = FA(qx, FS(FA(FM(wx, qw), FM(wy, qz)), FM(wz, qy)));
    float ny = FA(qy, FA(FS(FM(wy, qw), FM(wx, qz)), FM(wz, qx)));
    float nz = FA(qz, FS(FA(FM(wz, qw), FM(wx, qy)), FM(wy, qx)));
    const float inv = __fdiv_rn(1.0f, __fsqrt_rn(FA(FA(FM(nw, nw), FM(nx, nx)), FA(FM(ny, ny), FM(nz, nz)))));
    e.rotation[0] = FM(nw, inv); e.rotation[1] = FM(nx, inv); e.rotation[2] = FM(ny, inv); e.rotation[3] = FM(nz, inv);
    ents[i] = e;
    bodies[i] = b;
}

}  // namespace

void launchStepSpheres(rg_entity* dEntities, rg_sphere_body* dBodies, uint32_t n, float dt, float floorY, cudaStream_t st) {
    if(n) k_step_spheres<<<(n + 127) / 128, 128, 0, st>>>(dEntities, dBodies, n, dt, floorY);
}

void launchEntityInstances(const rg_entity* dEntities, uint32_t n, rg_instance* dTmp, uint32_t* dEmit, rg_instance* dOut, uint32_t* dCount, cudaStream_t st) {
    if(n == 0) { cudaMemsetAsync(dCount, 0, 4, st); return; }
    k_entity_instances<<<(n + 127) / 128, 128, 0, st>>>(dEntities, n, dTmp, dEmit);
    k_entity_compact<<<1, 1024, 0, st>>>(dTmp, dEmit, n, dOut, dCount);
}

}  // namespace rg
