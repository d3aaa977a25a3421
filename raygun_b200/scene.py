"""Host-side scene data for the B200 ray-tracing path (Python mirror).

Holds exactly what the reference hands the GPU per scene / per frame:
the packed vertex / index / material buffers of RenderSystem::setupModelBuffers
(raygun/render/render_system.cpp:192-223, :270-330), the instance list + offset table of
TopLevelAS (raygun/render/acceleration_structure.cpp:34-85) and the UniformBufferObject
(resources/shaders/uniform_buffer_object.def:3-17, raygun/render/render_system.cpp:235-268).

Also builds the benchmark scenes of BASELINE.md section 2 (C1/C2/C5 example scene from the committed
snapshot, C3 sphere grid, C4 animated instances).  Pure numpy; no device code here.
"""
from __future__ import annotations

import dataclasses
import math
import os

import numpy as np

_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

F32 = np.float32


# ----------------------------------------------------------------------------- transforms
def quat_mul(a, b):
    """Hamilton product, (w, x, y, z); glm::operator*(quat, quat)."""
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz, aw * bz + az * bw + ax * by - ay * bx], F32)


def quat_to_mat3(q):
    """glm::mat3_cast."""
    w, x, y, z = (F32(v) for v in q)
    one, two = F32(1), F32(2)
    return np.array([[one - two * (y * y + z * z), two * (x * y - w * z), two * (x * z + w * y)],
                     [two * (x * y + w * z), one - two * (x * x + z * z), two * (y * z - w * x)],
                     [two * (x * z - w * y), two * (y * z + w * x), one - two * (x * x + y * y)]], F32)


def quat_rotate(q, v):
    """glm::rotate(quat, vec3) == q * v."""
    w = F32(q[0]); qv = np.asarray(q[1:], F32); v = np.asarray(v, F32)
    uv = np.cross(qv, v).astype(F32)
    uuv = np.cross(qv, uv).astype(F32)
    return (v + ((uv * w) + uuv) * F32(2)).astype(F32)


def quat_angle_axis(angle, axis):
    axis = np.asarray(axis, F32)
    s = F32(math.sin(angle * 0.5))
    return np.array([F32(math.cos(angle * 0.5)), axis[0] * s, axis[1] * s, axis[2] * s], F32)


def quat_from_mat3(m):
    """glm::quat_cast (rows of m are matrix rows)."""
    m = np.asarray(m, np.float64)
    fx = m[0, 0] - m[1, 1] - m[2, 2]
    fy = m[1, 1] - m[0, 0] - m[2, 2]
    fz = m[2, 2] - m[0, 0] - m[1, 1]
    fw = m[0, 0] + m[1, 1] + m[2, 2]
    big, idx = fw, 0
    for i, f in enumerate((fx, fy, fz), 1):
        if f > big:
            big, idx = f, i
    bv = math.sqrt(big + 1.0) * 0.5
    mult = 0.25 / bv
    # glm indexes m[col][row]; here m[row, col]
    if idx == 0:
        q = (bv, (m[2, 1] - m[1, 2]) * mult, (m[0, 2] - m[2, 0]) * mult, (m[1, 0] - m[0, 1]) * mult)
    elif idx == 1:
        q = ((m[2, 1] - m[1, 2]) * mult, bv, (m[1, 0] + m[0, 1]) * mult, (m[0, 2] + m[2, 0]) * mult)
    elif idx == 2:
        q = ((m[0, 2] - m[2, 0]) * mult, (m[1, 0] + m[0, 1]) * mult, bv, (m[2, 1] + m[1, 2]) * mult)
    else:
        q = ((m[1, 0] - m[0, 1]) * mult, (m[0, 2] + m[2, 0]) * mult, (m[2, 1] + m[1, 2]) * mult, bv)
    return np.array(q, F32)


def quat_look_at(direction, up=(0.0, 1.0, 0.0)):
    """glm::quatLookAtRH (raygun/transform.hpp:82-86)."""
    d = np.asarray(direction, np.float64)
    up = np.asarray(up, np.float64)
    c2 = -d
    c0 = np.cross(up, c2); c0 /= np.linalg.norm(c0)
    c1 = np.cross(c2, c0)
    return quat_from_mat3(np.stack([c0, c1, c2], axis=1))


@dataclasses.dataclass
class Transform:
    """TRS transform; composition and matrix form as raygun/transform.hpp:38-46, :99-106."""
    position: np.ndarray = dataclasses.field(default_factory=lambda: np.zeros(3, F32))
    rotation: np.ndarray = dataclasses.field(default_factory=lambda: np.array([1, 0, 0, 0], F32))  # w x y z
    scaling: np.ndarray = dataclasses.field(default_factory=lambda: np.ones(3, F32))

    def __matmul__(self, y: "Transform") -> "Transform":
        return Transform(quat_rotate(self.rotation, self.scaling * y.position) + self.position,
                         quat_mul(self.rotation, y.rotation), (self.scaling * y.scaling).astype(F32))

    def to_3x4(self) -> np.ndarray:
        """Row-major 3x4 object->world = rows 0..2 of T*R*S (acceleration_structure.cpp:44-45)."""
        r = quat_to_mat3(self.rotation) * np.asarray(self.scaling, F32)[None, :]
        return np.concatenate([r, np.asarray(self.position, F32)[:, None]], axis=1).astype(F32)

    def to_mat4_colmajor(self) -> np.ndarray:
        m = np.eye(4, dtype=F32)
        m[:3, :] = self.to_3x4()
        return m.T.reshape(16).copy()  # column-major flat

    def is_zero_volume(self) -> bool:
        s = self.scaling
        return float(s[0] * s[1] * s[2]) == 0.0

    def look_at(self, target):
        d = np.asarray(target, F32) - self.position
        self.rotation = quat_look_at(d / F32(np.linalg.norm(d)))


def proj_inverse(width: int, height: int, fov_deg=45.0, near=0.1, far=100.0) -> np.ndarray:
    """inverse(perspectiveRH_ZO(fov, w/h, near, far) with [1][1] *= -1), column-major flat
    (raygun/camera.cpp:34-47, raygun/camera.hpp:36, GLM_FORCE_DEPTH_ZERO_TO_ONE pch.hpp:82)."""
    aspect = F32(width) / F32(height)
    t = F32(math.tan(F32(math.radians(fov_deg)) / F32(2)))
    p00 = F32(1) / (aspect * t)
    p11 = -(F32(1) / t)
    p22 = F32(far) / (F32(near) - F32(far))
    p32 = -(F32(far) * F32(near)) / (F32(far) - F32(near))  # glm Result[3][2]
    inv = np.zeros((4, 4), F32)  # inv[col][row]
    inv[0][0] = F32(1) / p00
    inv[1][1] = F32(1) / p11
    inv[2][3] = F32(1) / p32
    inv[3][2] = F32(-1)
    inv[3][3] = p22 / p32
    return inv.reshape(16).copy()


# ----------------------------------------------------------------------------- POD layouts
def make_material(diffuse=(1.0, 0.0, 1.0), transparency=0.0, specular=(1.0, 1.0, 1.0), reflectivity=0.0, roughness=0.0,
                  ior=1.0, effectId=0, rayConsumption=1, emission=0.0) -> np.ndarray:
    """64-byte gpu::Material (resources/shaders/gpu_material.def:11-26) as a (16,) uint32 record."""
    m = np.zeros(16, np.uint32)
    f = m.view(F32)
    f[0:3] = diffuse; f[3] = transparency; f[4:7] = specular; f[7] = reflectivity
    f[8] = roughness; f[9] = ior; m[10] = effectId; m[11] = rayConsumption; f[12] = emission
    return m


def make_ubo(view_inverse, proj_inv, num_samples=1, max_recursions=5, light_dir=None, fade=(0, 0, 0, 0), show_alpha=False,
             time=0.0) -> np.ndarray:
    """192-byte UniformBufferObject (uniform_buffer_object.def:3-17); defaults render_system.cpp:235-244."""
    u = np.zeros(48, np.uint32)
    f = u.view(F32)
    f[0:16] = view_inverse; f[16:32] = proj_inv
    f[32:35] = (0.2, 0.2, 0.2); u[35] = np.uint32(num_samples)
    if light_dir is None:
        l = np.array([.4, -.6, -.8], F32)
        light_dir = l * (F32(1) / np.sqrt(np.dot(l, l), dtype=F32))
    f[36:39] = light_dir; u[39] = np.uint32(max_recursions)
    f[40] = time; u[41] = 1 if show_alpha else 0
    f[44:48] = fade
    return u


@dataclasses.dataclass
class SceneData:
    vertices: np.ndarray     # (N, 8) uint32 : 32-byte Vertex records (vertex.def:3-7)
    indices: np.ndarray      # (M,) uint32, mesh-local
    meshes: np.ndarray       # (K, 4) uint32 : vtx_off, vtx_cnt, idx_off, idx_cnt (elements)
    materials: np.ndarray    # (L, 16) uint32 : 64-byte gpu::Material records
    inst_xform: np.ndarray   # (I, 12) float32 row-major 3x4 object->world
    inst_meta: np.ndarray    # (I, 4) uint32 : mesh, vtx_off, idx_off, mat_off
    name: str = "scene"

    @property
    def n_triangles_instanced(self) -> int:
        return int(sum(self.meshes[m, 3] // 3 for m in self.inst_meta[:, 0]))

    def positions(self) -> np.ndarray:
        return self.vertices.view(F32)[:, 0:3]


def _bits(a):
    return np.asarray(a, np.uint32).view(F32)


def load_example_scene() -> tuple[SceneData, dict]:
    """The reference's example scene (example/example_scene.cpp:9-32) from the committed snapshot
    tests/golden/example_scene.npz (made by tools/make_example_scene.py from the reference's assets)."""
    z = np.load(os.path.join(_GOLDEN, "example_scene.npz"))
    sd = SceneData(vertices=z["vertices"].copy(), indices=z["indices"].copy(), meshes=z["meshes"].copy(),
                   materials=z["materials"].view(np.uint32).reshape(-1, 16).copy(), inst_xform=_bits(z["instance_xform"]).reshape(-1, 12).copy(),
                   inst_meta=z["instance_mesh_voff_ioff_moff"].copy(), name="example")
    cam = dict(view_inverse=_bits(z["view_inverse"]).copy(), light_dir=_bits(z["light_dir"]).copy())
    return sd, cam


def load_text_scene() -> tuple[SceneData, np.ndarray]:
    """Ray-traced UI text (raygun/ui/text.cpp) over a ground quad: the committed snapshot tests/golden/text_scene.npz, made by
    tools/make_text_scene.py with the C++ host shim's loadFont / TextGenerator from the reference's NotoSans.obj.  One instance per
    glyph.  Returns the scene and the 48-word UBO prefix (view / projection inverse for 640x360) the shim's Camera produced."""
    z = np.load(os.path.join(_GOLDEN, "text_scene.npz"))
    inst = z["instances"]
    sd = SceneData(vertices=z["vertices"].copy(), indices=z["indices"].copy(), meshes=z["meshes"].copy(), materials=z["materials"].copy(),
                   inst_xform=inst[:, :12].copy().view(F32), inst_meta=inst[:, 12:].copy(), name="text")
    return sd, z["ubo"].copy()


def example_ubo(width, height, num_samples=1, max_recursions=5, **kw) -> np.ndarray:
    _, cam = load_example_scene()
    return make_ubo(cam["view_inverse"], proj_inverse(width, height), num_samples, max_recursions, cam["light_dir"], **kw)


def example_camera_transform() -> Transform:
    """Camera pose of C1/C2/C5: position ball + (5,10,10), lookAt(ball) (example_scene.hpp:16, example_scene.cpp:58-62)."""
    t = Transform(position=np.array([8, 10, 7], F32))
    t.look_at(np.array([3, 0, -3], F32))
    return t


# ----------------------------------------------------------------------------- synthetic scenes
def _ball_mesh(sd: SceneData):
    """Vertices / indices of the 1 280-triangle ball mesh (4th mesh of the example snapshot)."""
    vo, vc, io, ic = (int(v) for v in sd.meshes[3])
    return sd.vertices[vo:vo + vc].copy(), sd.indices[io:io + ic].copy()


GLASS = dict(diffuse=(1, 1, 1), specular=(1, 1, 1), reflectivity=0.9, transparency=0.98, roughness=0.0, ior=1.5)
MIRROR = dict(diffuse=(.9, .9, .9), specular=(1, 1, 1), reflectivity=0.9, transparency=0.0, ior=1.0)
FLOOR = dict(diffuse=(1, 1, 1), specular=(1, 1, 1), reflectivity=0.3, transparency=0.0, roughness=0.0, ior=1.0, effectId=1)


def sphere_grid_scene(n=28, flattened=False) -> tuple[SceneData, np.ndarray]:
    """n x n spheres (ball mesh, radius 1) at (2.5 i, 1, 2.5 j), checkerboard mirror / glass,
    floor quad 80 x 80.  flattened=False: n*n instances of one BLAS (variant A); True: one pre-transformed mesh (variant B).
    Returns (scene, view_inverse)."""
    ex, _ = load_example_scene()
    bv, bi = _ball_mesh(ex)
    nb = len(bv)
    c = 2.5 * (n - 1) / 2.0
    floor = np.zeros((4, 8), np.uint32)
    ff = floor.view(F32)
    for k, (x, z) in enumerate(((c - 40, c - 40), (c + 40, c - 40), (c + 40, c + 40), (c - 40, c + 40))):
        ff[k, 0:3] = (x, 0.0, z); ff[k, 4:7] = (0, 1, 0)
    floor_idx = np.array([0, 2, 1, 0, 3, 2], np.uint32)
    mats = np.stack([make_material(**MIRROR), make_material(**GLASS), make_material(**FLOOR)])
    if not flattened:
        # two ball meshes sharing geometry but with matIndex 0 / 1 would double memory; instead one mesh and
        # per-instance material offset (instance_offset_table.def: materialBufferOffset) selects mirror/glass.
        bv[:, 3] = 0
        vertices = np.concatenate([bv, floor])
        indices = np.concatenate([bi, floor_idx])
        meshes = np.array([(0, nb, 0, len(bi)), (nb, 4, len(bi), 6)], np.uint32)
        xf, meta = [], []
        for i in range(n):
            for j in range(n):
                xf.append(Transform(position=np.array([2.5 * i, 1.0, 2.5 * j], F32)).to_3x4().reshape(12))
                meta.append((0, 0, 0, (i + j) & 1))
        xf.append(Transform().to_3x4().reshape(12)); meta.append((1, nb, len(bi), 2))
        sd = SceneData(vertices, indices, meshes, mats, np.stack(xf).astype(F32), np.array(meta, np.uint32), name=f"spheres{n}x{n}_inst")
    else:
        vs, ids = [], []
        for i in range(n):
            for j in range(n):
                v = bv.copy()
                p = v.view(F32)
                p[:, 0] += F32(2.5 * i); p[:, 1] += F32(1.0); p[:, 2] += F32(2.5 * j)
                v[:, 3] = (i + j) & 1
                ids.append(bi + np.uint32(len(vs) * nb)); vs.append(v)
        fl = floor.copy(); fl[:, 3] = 2
        ids.append(floor_idx + np.uint32(len(vs) * nb)); vs.append(fl)
        vertices = np.concatenate(vs); indices = np.concatenate(ids)
        meshes = np.array([(0, len(vertices), 0, len(indices))], np.uint32)
        sd = SceneData(vertices, indices, meshes, mats, Transform().to_3x4().reshape(1, 12), np.array([(0, 0, 0, 0)], np.uint32),
                       name=f"spheres{n}x{n}_flat")
    cam = Transform(position=np.array([35, 18, -20], F32))
    cam.look_at(np.array([c, 1.0, c], F32))
    return sd, cam.to_mat4_colmajor()


def pcg32(seed: int, stream: int, count: int) -> np.ndarray:
    """PCG32 (XSH-RR) reference generator; `count` 32-bit outputs."""
    mask = (1 << 64) - 1
    inc = ((stream << 1) | 1) & mask
    state = 0
    state = (state * 6364136223846793005 + inc) & mask
    state = (state + seed) & mask
    state = (state * 6364136223846793005 + inc) & mask
    out = np.empty(count, np.uint32)
    for i in range(count):
        old = state
        state = (old * 6364136223846793005 + inc) & mask
        xs = (((old >> 18) ^ old) >> 27) & 0xffffffff
        rot = old >> 59
        out[i] = ((xs >> rot) | (xs << ((-rot) & 31))) & 0xffffffff
    return out


class AnimatedBalls:
    """n x n instances of the ball BLAS, spacing 2.5, closed-form bounce
    y_k(t) = 1 + 3 |sin(2 pi (0.5 + u_k) t + 2 pi v_k)|, rotation about Y by t (1 + u_k); u_k, v_k from PCG32(0x5EED, k)."""

    def __init__(self, n=100):
        ex, _ = load_example_scene()
        bv, bi = _ball_mesh(ex)
        bv[:, 3] = 0
        self.n = n
        nb = len(bv)
        c = 2.5 * (n - 1) / 2.0
        floor = np.zeros((4, 8), np.uint32)
        ff = floor.view(F32)
        e = c + 10
        for k, (x, z) in enumerate(((c - e, c - e), (c + e, c - e), (c + e, c + e), (c - e, c + e))):
            ff[k, 0:3] = (x, 0.0, z); ff[k, 4:7] = (0, 1, 0)
        self.vertices = np.concatenate([bv, floor])
        self.indices = np.concatenate([bi, np.array([0, 2, 1, 0, 3, 2], np.uint32)])
        self.meshes = np.array([(0, nb, 0, len(bi)), (nb, 4, len(bi), 6)], np.uint32)
        self.materials = np.stack([make_material(**MIRROR), make_material(**GLASS), make_material(**FLOOR)])
        uv = np.stack([pcg32(0x5EED, k, 2) for k in range(n * n)]).astype(np.float64) / 4294967296.0
        self.u, self.v = uv[:, 0], uv[:, 1]
        ii, jj = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
        self.gx, self.gz = (2.5 * ii).reshape(-1), (2.5 * jj).reshape(-1)
        self.meta = np.concatenate([np.stack([np.zeros(n * n), np.zeros(n * n), np.zeros(n * n), (ii + jj).reshape(-1) & 1], axis=1),
                                    [[1, nb, len(bi), 2]]]).astype(np.uint32)
        cam = Transform(position=np.array([c + 60, 45, -40], F32))
        cam.look_at(np.array([c, 1.0, c], F32))
        self.view_inverse = cam.to_mat4_colmajor()

    def instances(self, t: float) -> np.ndarray:
        y = 1.0 + 3.0 * np.abs(np.sin(2 * np.pi * (0.5 + self.u) * t + 2 * np.pi * self.v))
        ang = t * (1.0 + self.u)
        cs, sn = np.cos(ang), np.sin(ang)
        xf = np.zeros((self.n * self.n + 1, 12), F32)
        xf[:-1, 0] = cs; xf[:-1, 2] = sn; xf[:-1, 3] = self.gx
        xf[:-1, 5] = 1; xf[:-1, 7] = y
        xf[:-1, 8] = -sn; xf[:-1, 10] = cs; xf[:-1, 11] = self.gz
        xf[-1, 0] = xf[-1, 5] = xf[-1, 10] = 1
        return xf

    def entities(self, t: float) -> np.ndarray:
        """The same frame as a scene graph for rg_set_entities: a root, the balls as its children (local TRS: bounce height, rotation
        about Y as a quaternion) and the floor.  The device composes the transforms; the host only fills 10 floats per ball."""
        from . import ENTITY_DTYPE, RG_ENTITY_VISIBLE, RG_ENTITY_HAS_MODEL
        nb = self.n * self.n
        e = np.zeros(nb + 2, ENTITY_DTYPE)
        e["parent"][0] = -1; e["parent"][1:] = 0
        e["rotation"][:, 0] = 1.0
        e["scaling"][:] = 1.0
        e["flags"][0] = RG_ENTITY_VISIBLE
        e["flags"][1:] = RG_ENTITY_VISIBLE | RG_ENTITY_HAS_MODEL
        y = 1.0 + 3.0 * np.abs(np.sin(2 * np.pi * (0.5 + self.u) * t + 2 * np.pi * self.v))
        half = 0.5 * t * (1.0 + self.u)
        e["position"][1:-1, 0] = self.gx; e["position"][1:-1, 1] = y; e["position"][1:-1, 2] = self.gz
        e["rotation"][1:-1, 0] = np.cos(half); e["rotation"][1:-1, 2] = np.sin(half)
        for k, name in enumerate(("mesh", "vtx_off", "idx_off", "mat_off")):
            e[name][1:] = self.meta[:, k]
        return e

    def scene(self, t: float) -> SceneData:
        return SceneData(self.vertices, self.indices, self.meshes, self.materials, self.instances(t), self.meta, name=f"balls{self.n}x{self.n}")
