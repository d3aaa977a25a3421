// TEST INFRASTRUCTURE -- CPU oracle (see orc_math.h / orc_scene.h headers).
#include "orc_scene.h"

#include <algorithm>
#include <cfloat>
#include <cmath>

namespace orc {

namespace {

struct Box {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    void grow(const float* p) { for(int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], p[a]); hi[a] = std::max(hi[a], p[a]); } }
    void grow(const Box& b) { grow(b.lo); grow(b.hi); }
};

// Median split on the longest centroid axis; leaves hold <= 4 items.  Only an accelerator.
void buildRec(Bvh2& bvh, const std::vector<Box>& boxes, uint32_t first, uint32_t count, uint32_t nodeIdx) {
    Box b, cb;
    for(uint32_t i = first; i < first + count; ++i) {
        const Box& ib = boxes[bvh.items[i]];
        b.grow(ib);
        const float c[3] = {0.5f * (ib.lo[0] + ib.hi[0]), 0.5f * (ib.lo[1] + ib.hi[1]), 0.5f * (ib.lo[2] + ib.hi[2])};
        cb.grow(c);
    }
    // padded: neither the slab arithmetic nor the triangle test is exact.  The padding of EVERY axis follows the largest coordinate of
    // the box: the rounding noise of the watertight test scales with the operands of all three axes, so a flat box (a floor at y = 0)
    // must not be padded by its own (zero) extent only -- a ray leaving such a plane with a tiny un-normalised direction is reported
    // as a hit by the loop over all triangles (the definition) at a t that is pure noise, and the accelerator has to visit the leaf
    // (found by tests/test_gpu_traversal.py::test_skewed_and_grazing_rays)
    float m = 0.0f;
    for(int a = 0; a < 3; ++a) m = std::max(m, std::max(std::max(std::fabs(b.lo[a]), std::fabs(b.hi[a])), b.hi[a] - b.lo[a]));
    for(int a = 0; a < 3; ++a) {
        const float e = 2e-6f * m + 1e-30f;
        bvh.nodes[nodeIdx].lo[a] = b.lo[a] - e; bvh.nodes[nodeIdx].hi[a] = b.hi[a] + e;
    }
    if(count <= 4) { bvh.nodes[nodeIdx].left = first; bvh.nodes[nodeIdx].count = count; return; }
    int axis = 0;
    for(int a = 1; a < 3; ++a) if(cb.hi[a] - cb.lo[a] > cb.hi[axis] - cb.lo[axis]) axis = a;
    const uint32_t mid = count / 2;
    std::nth_element(bvh.items.begin() + first, bvh.items.begin() + first + mid, bvh.items.begin() + first + count,
                     [&](uint32_t x, uint32_t y) {
                         const float cx = boxes[x].lo[axis] + boxes[x].hi[axis], cy = boxes[y].lo[axis] + boxes[y].hi[axis];
                         return cx < cy || (cx == cy && x < y);
                     });
    const uint32_t l = (uint32_t)bvh.nodes.size();
    bvh.nodes.emplace_back(); bvh.nodes.emplace_back();
    bvh.nodes[nodeIdx].left = l; bvh.nodes[nodeIdx].count = 0;
    buildRec(bvh, boxes, first, mid, l);
    buildRec(bvh, boxes, first + mid, count - mid, l + 1);
}

void buildBvh(Bvh2& bvh, const std::vector<Box>& boxes) {
    bvh.nodes.clear(); bvh.items.resize(boxes.size());
    for(uint32_t i = 0; i < boxes.size(); ++i) bvh.items[i] = i;
    bvh.nodes.reserve(boxes.size() * 2 + 1);
    bvh.nodes.emplace_back();
    if(boxes.empty()) { bvh.nodes[0].left = 0; bvh.nodes[0].count = 0; for(int a = 0; a < 3; ++a) { bvh.nodes[0].lo[a] = 1; bvh.nodes[0].hi[a] = -1; } return; }
    buildRec(bvh, boxes, 0, (uint32_t)boxes.size(), 0);
}

void invert3x4(const float* m, float* inv) {
    const double a = m[0], b = m[1], c = m[2], d = m[4], e = m[5], f = m[6], g = m[8], h = m[9], i = m[10];
    const double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
    const double det = a * A + b * B + c * C;
    const double r = 1.0 / det;
    const double R[9] = {A * r, -(b * i - c * h) * r, (b * f - c * e) * r,
                         B * r, (a * i - c * g) * r, -(a * f - c * d) * r,
                         C * r, -(a * h - b * g) * r, (a * e - b * d) * r};
    const double tx = m[3], ty = m[7], tz = m[11];
    for(int rr = 0; rr < 3; ++rr) {
        inv[rr * 4 + 0] = (float)R[rr * 3 + 0]; inv[rr * 4 + 1] = (float)R[rr * 3 + 1]; inv[rr * 4 + 2] = (float)R[rr * 3 + 2];
        inv[rr * 4 + 3] = (float)(-(R[rr * 3 + 0] * tx + R[rr * 3 + 1] * ty + R[rr * 3 + 2] * tz));
    }
}

struct RayPre { int kx, ky, kz; float Sx, Sy, Sz; };

inline RayPre precompute(vec3 dir) {
    RayPre r;
    const float ax = std::fabs(dir.x), ay = std::fabs(dir.y), az = std::fabs(dir.z);
    r.kz = (ax > ay) ? ((ax > az) ? 0 : 2) : ((ay > az) ? 1 : 2);
    r.kx = (r.kz + 1) % 3; r.ky = (r.kx + 1) % 3;
    if(dir[r.kz] < 0.0f) std::swap(r.kx, r.ky);
    r.Sx = dir[r.kx] / dir[r.kz]; r.Sy = dir[r.ky] / dir[r.kz]; r.Sz = 1.0f / dir[r.kz];
    return r;
}

// Woop/Benthin/Wald watertight test, no culling.  Returns t,u(=weight of v1),v(=weight of v2).
inline bool intersectTri(const RayPre& rp, vec3 org, const float* p0, const float* p1, const float* p2, float tmin, float tmax,
                         float& tOut, float& uOut, float& vOut) {
    const vec3 A(p0[0] - org.x, p0[1] - org.y, p0[2] - org.z);
    const vec3 B(p1[0] - org.x, p1[1] - org.y, p1[2] - org.z);
    const vec3 C(p2[0] - org.x, p2[1] - org.y, p2[2] - org.z);
    const float Ax = A[rp.kx] - rp.Sx * A[rp.kz], Ay = A[rp.ky] - rp.Sy * A[rp.kz];
    const float Bx = B[rp.kx] - rp.Sx * B[rp.kz], By = B[rp.ky] - rp.Sy * B[rp.kz];
    const float Cx = C[rp.kx] - rp.Sx * C[rp.kz], Cy = C[rp.ky] - rp.Sy * C[rp.kz];
    float U = Cx * By - Cy * Bx, V = Ax * Cy - Ay * Cx, W = Bx * Ay - By * Ax;
    if(U == 0.0f || V == 0.0f || W == 0.0f) {
        const double CxBy = (double)Cx * (double)By, CyBx = (double)Cy * (double)Bx; U = (float)(CxBy - CyBx);
        const double AxCy = (double)Ax * (double)Cy, AyCx = (double)Ay * (double)Cx; V = (float)(AxCy - AyCx);
        const double BxAy = (double)Bx * (double)Ay, ByAx = (double)By * (double)Ax; W = (float)(BxAy - ByAx);
    }
    if((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    const float det = U + V + W;
    if(det == 0.0f) return false;
    const float Az = rp.Sz * A[rp.kz], Bz = rp.Sz * B[rp.kz], Cz = rp.Sz * C[rp.kz];
    const float T = U * Az + V * Bz + W * Cz;
    const float rcpDet = 1.0f / det;
    const float t = T * rcpDet;
    if(!(t > tmin && t < tmax)) return false;
    tOut = t; uOut = V * rcpDet; vOut = W * rcpDet;
    return true;
}

inline bool slab(const float* lo, const float* hi, vec3 org, vec3 idir, float tmin, float tmax) {
    float t0 = tmin, t1 = tmax;
    for(int a = 0; a < 3; ++a) {
        float n = (lo[a] - org[a]) * idir[a], f = (hi[a] - org[a]) * idir[a];
        if(n > f) std::swap(n, f);
        // conservative (cf. Ize 2013): never cull a box the exact triangle test would enter
        n -= std::fabs(n) * 2e-6f; f += std::fabs(f) * 2e-6f;
        // NaN (0*inf) means the ray lies in the slab plane: keep the interval
        if(n == n) t0 = n > t0 ? n : t0;
        if(f == f) t1 = f < t1 ? f : t1;
    }
    return t0 <= t1;
}

inline bool better(float t, uint32_t inst, uint32_t prim, const Hit& h) {
    return t < h.t || (t == h.t && (inst < h.inst || (inst == h.inst && prim < h.prim)));
}

}  // namespace

void Scene::build() {
    blas.assign(meshes.size(), Bvh2());
    std::vector<Box> meshBox(meshes.size());
    for(size_t m = 0; m < meshes.size(); ++m) {
        const MeshRange& r = meshes[m];
        const uint32_t nTri = r.idxCnt / 3;
        std::vector<Box> boxes(nTri);
        for(uint32_t p = 0; p < nTri; ++p)
            for(int k = 0; k < 3; ++k) {
                const Vertex& v = vertices[r.vtxOff + indices[r.idxOff + 3 * p + k]];
                boxes[p].grow(&v.px);
            }
        for(auto& b: boxes) meshBox[m].grow(b);
        buildBvh(blas[m], boxes);
    }
    std::vector<Box> ib(instances.size());
    for(size_t i = 0; i < instances.size(); ++i) {
        Instance& in = instances[i];
        invert3x4(in.m, in.inv);
        const Box& mb = meshBox[in.mesh];
        if(meshes[in.mesh].idxCnt == 0) { const float z[3] = {in.m[3], in.m[7], in.m[11]}; ib[i].grow(z); continue; }
        for(int c = 0; c < 8; ++c) {
            const float x = (c & 1) ? mb.hi[0] : mb.lo[0], y = (c & 2) ? mb.hi[1] : mb.lo[1], z = (c & 4) ? mb.hi[2] : mb.lo[2];
            const float w[3] = {in.m[0] * x + in.m[1] * y + in.m[2] * z + in.m[3], in.m[4] * x + in.m[5] * y + in.m[6] * z + in.m[7],
                                in.m[8] * x + in.m[9] * y + in.m[10] * z + in.m[11]};
            ib[i].grow(w);
        }
        // pad: the float transform above is not exact
        for(int a = 0; a < 3; ++a) { const float e = 1e-5f * std::max(std::fabs(ib[i].lo[a]), std::fabs(ib[i].hi[a])) + 1e-30f; ib[i].lo[a] -= e; ib[i].hi[a] += e; }
    }
    buildBvh(tlas, ib);
}

bool Scene::closestHit(vec3 org, vec3 dir, float tmin, float tmax, Hit& hit, bool brute) const {
    hit.t = tmax; hit.inst = 0xffffffffu; hit.prim = 0xffffffffu; hit.u = hit.v = 0;
    if(dir.x == 0.0f && dir.y == 0.0f && dir.z == 0.0f) return false;  // zero direction (hazard 7): miss
    if(!(dir.x == dir.x && dir.y == dir.y && dir.z == dir.z) || !(org.x == org.x && org.y == org.y && org.z == org.z)) return false;

    auto testInstance = [&](uint32_t ii) {
        const Instance& in = instances[ii];
        const MeshRange& mr = meshes[in.mesh];
        const float* w = in.inv;
        const vec3 o(w[0] * org.x + w[1] * org.y + w[2] * org.z + w[3], w[4] * org.x + w[5] * org.y + w[6] * org.z + w[7],
                     w[8] * org.x + w[9] * org.y + w[10] * org.z + w[11]);
        const vec3 d(w[0] * dir.x + w[1] * dir.y + w[2] * dir.z, w[4] * dir.x + w[5] * dir.y + w[6] * dir.z,
                     w[8] * dir.x + w[9] * dir.y + w[10] * dir.z);
        if(d.x == 0.0f && d.y == 0.0f && d.z == 0.0f) return;
        const RayPre rp = precompute(d);
        auto testTri = [&](uint32_t p) {
            const uint32_t* ix = &indices[mr.idxOff + 3 * p];
            const Vertex &v0 = vertices[mr.vtxOff + ix[0]], &v1 = vertices[mr.vtxOff + ix[1]], &v2 = vertices[mr.vtxOff + ix[2]];
            float t, u, v;
            // the interval's upper end stays the ORIGINAL tmax; ties are resolved by `better`
            if(intersectTri(rp, o, &v0.px, &v1.px, &v2.px, tmin, tmax, t, u, v) && better(t, ii, p, hit)) {
                hit.t = t; hit.u = u; hit.v = v; hit.inst = ii; hit.prim = p;
            }
        };
        if(brute) { for(uint32_t p = 0; p < mr.idxCnt / 3; ++p) testTri(p); return; }
        const Bvh2& bvh = blas[in.mesh];
        if(bvh.items.empty()) return;
        const vec3 id(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
        uint32_t stack[128]; int sp = 0; stack[sp++] = 0;
        while(sp) {
            const Bvh2::Node& n = bvh.nodes[stack[--sp]];
            if(!slab(n.lo, n.hi, o, id, tmin, hit.t)) continue;
            if(n.count) { for(uint32_t k = 0; k < n.count; ++k) testTri(bvh.items[n.left + k]); }
            else { stack[sp++] = n.left; stack[sp++] = n.left + 1; }
        }
    };

    if(brute) { for(uint32_t i = 0; i < instances.size(); ++i) testInstance(i); }
    else if(!tlas.items.empty()) {
        const vec3 id(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z);
        uint32_t stack[128]; int sp = 0; stack[sp++] = 0;
        while(sp) {
            const Bvh2::Node& n = tlas.nodes[stack[--sp]];
            if(!slab(n.lo, n.hi, org, id, tmin, hit.t)) continue;
            if(n.count) { for(uint32_t k = 0; k < n.count; ++k) testInstance(tlas.items[n.left + k]); }
            else { stack[sp++] = n.left; stack[sp++] = n.left + 1; }
        }
    }
    return hit.inst != 0xffffffffu;
}

}  // namespace orc
