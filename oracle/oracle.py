"""TEST INFRASTRUCTURE -- ctypes wrapper around the CPU oracle (oracle/_build/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module.  The product package (raygun_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

FXAA, SRGB8, STRICT_IEEE, BRUTE_FORCE, TRANSITIONS_UNORM = 1, 2, 4, 8, 16
RAY_KINDS = ("primary", "shadow", "reflect", "refract", "skylookup", "zero_dir")


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".h")) or f == "Makefile"]
    if force or not os.path.exists(_LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_scene_create.restype = C.c_void_p
        _lib.orc_trace.restype = C.c_double
        _lib.orc_post.restype = C.c_double
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class OracleScene:
    def __init__(self, sd):
        self.sd = sd
        self._keep = [np.ascontiguousarray(a) for a in (sd.vertices, sd.indices, sd.meshes, sd.materials, sd.inst_xform, sd.inst_meta)]
        v, i, m, mat, xf, meta = self._keep
        self.h = C.c_void_p(lib().orc_scene_create(_p(v), C.c_uint32(len(v)), _p(i), C.c_uint32(len(i)), _p(m), C.c_uint32(len(m)),
                                                   _p(mat), C.c_uint32(len(mat)), _p(xf), _p(meta), C.c_uint32(len(xf))))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_scene_destroy(self.h)
            self.h = None

    def set_instances(self, xform, meta):
        xform = np.ascontiguousarray(xform, np.float32); meta = np.ascontiguousarray(meta, np.uint32)
        lib().orc_scene_set_instances(self.h, _p(xform), _p(meta), C.c_uint32(len(xform)))

    def trace(self, ubo, W, H, flags=0, threads=0, rows=None):
        """raygen + recursion.  Returns dict(base, normal, rough (H,W,4) uint16 half bits; inst, prim (H,W) uint32; t; counters; seconds)."""
        ubo = np.ascontiguousarray(ubo, np.uint32)
        out = {k: np.zeros((H, W, 4), np.uint16) for k in ("base", "normal", "rough")}
        out["inst"] = np.full((H, W), 0xffffffff, np.uint32); out["prim"] = np.full((H, W), 0xffffffff, np.uint32)
        out["t"] = np.zeros((H, W), np.float32)
        cnt = np.zeros(6, np.uint64)
        y0, y1 = rows if rows else (0, H)
        sec = lib().orc_trace(self.h, _p(ubo), C.c_uint32(W), C.c_uint32(H), C.c_uint32(flags), C.c_int(threads), C.c_uint32(y0), C.c_uint32(y1),
                              _p(out["base"]), _p(out["normal"]), _p(out["rough"]), _p(out["inst"]), _p(out["prim"]), _p(out["t"]), _p(cnt))
        out["counters"] = dict(zip(RAY_KINDS, (int(c) for c in cnt)))
        out["seconds"] = sec
        return out

    def closest_hit(self, org, direction, tmin, tmax, brute=False):
        org = np.ascontiguousarray(org, np.float32); direction = np.ascontiguousarray(direction, np.float32)
        tuv = np.zeros(3, np.float32); ip = np.zeros(2, np.uint32)
        hit = lib().orc_closest_hit(self.h, _p(org), _p(direction), C.c_float(tmin), C.c_float(tmax), C.c_int(1 if brute else 0), _p(tuv), _p(ip))
        return bool(hit), float(tuv[0]), float(tuv[1]), float(tuv[2]), int(ip[0]), int(ip[1])

    def render(self, ubo, W, H, flags=FXAA, threads=0):
        """Whole path: trace + post chain.  Returns the 7 images + rgba8 + ids (all numpy)."""
        out = self.trace(ubo, W, H, flags, threads)
        gb = {k: out[k].copy() for k in ("base", "normal", "rough")}  # G-buffer as raygen wrote it
        post = post_chain(ubo, out["base"], out["normal"], out["rough"], flags, threads)
        out.update(post)
        out["gbuffer"] = gb
        out["seconds_trace"] = out["seconds"]
        out["seconds"] = out["seconds_trace"] + post["seconds_post"]
        return out


def post_chain(ubo, base, normal, rough, flags=FXAA, threads=0):
    """Post passes on (H,W,4) uint16 images; base/normal/rough are modified in place as the reference does."""
    H, W = base.shape[:2]
    ubo = np.ascontiguousarray(ubo, np.uint32)
    res = dict(final=np.zeros((H, W, 4), np.uint16), roughA=np.zeros((H, W, 4), np.uint16), roughB=np.zeros((H, W, 4), np.uint16),
               transitions=np.zeros((H, W), np.int8), rgba8=np.zeros((H, W, 4), np.uint8))
    sec = lib().orc_post(_p(ubo), C.c_uint32(W), C.c_uint32(H), C.c_uint32(flags), C.c_int(threads), _p(base), _p(normal), _p(rough),
                         _p(res["final"]), _p(res["roughA"]), _p(res["roughB"]), _p(res["transitions"]), _p(res["rgba8"]))
    res["seconds_post"] = sec
    res["base"], res["normal"], res["rough"] = base, normal, rough
    return res


def morton_triangles(sd, mesh):
    vo, vc, io, ic = (int(v) for v in sd.meshes[mesh])
    n = ic // 3
    idx = np.ascontiguousarray(sd.indices[io:io + ic])
    codes = np.zeros(n, np.uint32); order = np.zeros(n, np.uint32); box = np.zeros(6, np.float32)
    v = np.ascontiguousarray(sd.vertices)
    lib().orc_morton_triangles(_p(v), C.c_uint32(vo), _p(idx), C.c_uint32(n), _p(codes), _p(order), _p(box))
    return codes, order, box


def morton_boxes(boxes):
    boxes = np.ascontiguousarray(boxes, np.float32)
    n = len(boxes)
    codes = np.zeros(n, np.uint32); order = np.zeros(n, np.uint32)
    lib().orc_morton_boxes(_p(boxes), C.c_uint32(n), _p(codes), _p(order))
    return codes, order


def instance_world_box(xform12, mesh_box6):
    xf = np.ascontiguousarray(xform12, np.float32); mb = np.ascontiguousarray(mesh_box6, np.float32)
    out = np.zeros(6, np.float32)
    lib().orc_instance_world_box(_p(xf), _p(mb), _p(out))
    return out


def f32_to_f16(a):
    a = np.ascontiguousarray(a, np.float32); out = np.zeros(a.shape, np.uint16)
    lib().orc_f32_to_f16(_p(a), _p(out), C.c_uint64(a.size)); return out


def f16_to_f32(a):
    a = np.ascontiguousarray(a, np.uint16); out = np.zeros(a.shape, np.float32)
    lib().orc_f16_to_f32(_p(a), _p(out), C.c_uint64(a.size)); return out


def max_threads():
    return int(lib().orc_max_threads())


def step_spheres(entities, bodies, dt, floor_y=0.0):
    """CPU restatement (numpy binary32, one rounded operation at a time, same order) of raygun_b200/csrc/rg_scene.cu:k_step_spheres --
    the stand-in for PhysicsSystem::update (raygun/physics/physics_system.cpp:241-258; gravity :68, default restitution :40).
    entities: ENTITY_DTYPE array, bodies: SPHERE_BODY_DTYPE array; returns updated copies."""
    f = np.float32
    e, b = entities.copy(), bodies.copy()
    dyn = b["radius"] > 0
    dt = f(dt)
    vel = b["velocity"].copy(); pos = e["position"].copy()
    vel[:, 1] = vel[:, 1] + f(-9.81) * dt
    pos = pos + vel * dt
    rest = f(floor_y) + b["radius"]
    hit = (pos[:, 1] < rest) & (vel[:, 1] < 0)
    pos[:, 1] = np.where(hit, rest + (rest - pos[:, 1]) * b["restitution"], pos[:, 1])
    vel[:, 1] = np.where(hit, (-vel[:, 1]) * b["restitution"], vel[:, 1])
    h = f(0.5) * dt
    w = b["angular_velocity"] * h
    wx, wy, wz = w[:, 0], w[:, 1], w[:, 2]
    qw, qx, qy, qz = (e["rotation"][:, k] for k in range(4))
    nw = qw + (((-wx) * qx - wy * qy) - wz * qz)
    nx = qx + ((wx * qw + wy * qz) - wz * qy)
    ny = qy + ((wy * qw - wx * qz) + wz * qx)
    nz = qz + ((wz * qw + wx * qy) - wy * qx)
    inv = f(1.0) / np.sqrt((nw * nw + nx * nx) + (ny * ny + nz * nz))
    rot = np.stack([nw * inv, nx * inv, ny * inv, nz * inv], 1).astype(f)
    e["position"] = np.where(dyn[:, None], pos, e["position"])
    e["rotation"] = np.where(dyn[:, None], rot, e["rotation"])
    b["velocity"] = np.where(dyn[:, None], vel, b["velocity"])
    return e, b
