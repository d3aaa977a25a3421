"""raygun_b200 -- B200-native replacement for Raygun's per-frame ray-tracing path.

The product is the C-ABI shared library `librgb200.so` (include/rgb200.h), built from the hand-written
sm_100a CUDA kernels under raygun_b200/csrc.  This package is the thin Python host mirror used by the
tests and the benchmark: `Raytracer` keeps the method names of raygun::render::Raytracer
(raygun/render/raytracer.hpp:36-101) and calls straight through the C ABI with numpy host buffers.

There is no CPU fallback: importing works without a GPU (so the ABI can be checked), but creating a
`Raytracer` without a CUDA device, or without the built library, raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import scene  # noqa: F401  (host-side scene data)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RGB200_LIB") or os.path.join(_HERE, "librgb200.so")   # RGB200_LIB: developer override (A/B builds)

# rg_render flags (include/rgb200.h)
RG_FXAA, RG_SRGB8, RG_STRICT_IEEE, RG_DEBUG_IDS, RG_COUNT_TRAVERSAL, RG_NO_GATHER = 1, 2, 4, 8, 16, 32
RG_SCHED_LANES, RG_SCHED_POOL, RG_SCHED_AUTO = 0, 1, 2
RG_ENTITY_VISIBLE, RG_ENTITY_HAS_MODEL = 1, 2
# rg_entity (include/rgb200.h), 64 bytes: local TRS + parent + model references of one scene-graph node, DFS pre-order
ENTITY_DTYPE = np.dtype([("position", np.float32, 3), ("parent", np.int32), ("rotation", np.float32, 4), ("scaling", np.float32, 3), ("flags", np.uint32),
                         ("mesh", np.uint32), ("vtx_off", np.uint32), ("idx_off", np.uint32), ("mat_off", np.uint32)])
assert ENTITY_DTYPE.itemsize == 64
SPHERE_BODY_DTYPE = np.dtype([("velocity", np.float32, 3), ("radius", np.float32), ("angular_velocity", np.float32, 3), ("restitution", np.float32)])
assert SPHERE_BODY_DTYPE.itemsize == 32
IMG_FINAL, IMG_BASE, IMG_NORMAL, IMG_ROUGH, IMG_TRANSITIONS, IMG_ROUGH_A, IMG_ROUGH_B = range(7)

ABI_SYMBOLS = (
    "rg_create", "rg_destroy", "rg_resize", "rg_last_error", "rg_set_region", "rg_upload_geometry", "rg_upload_materials", "rg_build_blas",
    "rg_refit_blas", "rg_refit_blas_device", "rg_set_instances", "rg_set_ubo", "rg_render", "rg_sync", "rg_read_rgba8", "rg_read_image", "rg_read_ids", "rg_get_timings",
    "rg_set_instances_device", "rg_set_ubo_device", "rg_framebuffer_device_ptr", "rg_set_gather_target", "rg_gather_buffer_export",
    "rg_gather_buffer_open", "rg_gather_buffer_close", "rg_read_gathered_rgba8", "rg_debug_blas_sort", "rg_debug_tlas_sort", "rg_debug_trace_rays",
    "rg_debug_bvh_stats", "rg_debug_upload_gbuffer", "rg_debug_run_post", "rg_launch_count", "rg_timer_begin", "rg_timer_end", "rg_flush_l2", "rg_debug_last_trace_rays_ms",
    "rg_set_trace_scheduler", "rg_set_entities", "rg_set_entities_device", "rg_debug_read_instances", "rg_physics_step_spheres", "rg_set_partition", "rg_peer_export", "rg_peer_attach", "rg_peer_detach_all", "rg_sync_error",
    "rg_host_frame_register", "rg_host_frame_unregister",
)


class RgTimings(C.Structure):
    _fields_ = [("as_build_ms", C.c_float), ("rt_total_ms", C.c_float), ("rt_only_ms", C.c_float), ("rough_ms", C.c_float),
                ("postproc_ms", C.c_float), ("gather_ms", C.c_float),
                ("rays_primary", C.c_uint64), ("rays_shadow", C.c_uint64), ("rays_reflect", C.c_uint64), ("rays_refract", C.c_uint64),
                ("sky_lookups", C.c_uint64), ("nodes_visited", C.c_uint64), ("tris_tested", C.c_uint64), ("instances_entered", C.c_uint64),
                ("generic_hits", C.c_uint64), ("trace_kernel_ms", C.c_float), ("trace_scheduler", C.c_uint32)]


class RgPeerDesc(C.Structure):
    """rg_peer_desc (include/rgb200.h): one rank's G-buffer images, barrier flags, rectangle and CUDA IPC handles."""
    _fields_ = [("base", C.c_void_p), ("normal", C.c_void_p), ("rough", C.c_void_p), ("arrive_trace", C.c_void_p), ("arrive_post", C.c_void_p),
                ("x0", C.c_int32), ("y0", C.c_int32), ("w", C.c_int32), ("h", C.c_int32), ("ipc", (C.c_ubyte * 64) * 5)]


_lib = None


def load_library():
    """dlopen librgb200.so.  Raises (loudly) when the extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"raygun_b200: {LIB_PATH} is missing -- build it with `make -C raygun_b200/csrc` "
                               "(or __graft_entry__.build()); there is no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        lib.rg_last_error.restype = C.c_char_p
        lib.rg_launch_count.restype = C.c_uint64
        lib.rg_destroy.restype = None
        lib.rg_debug_last_trace_rays_ms.restype = C.c_float
        _lib = lib
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class RaygunError(RuntimeError):
    pass


class Raytracer:
    """Python mirror of raygun::render::Raytracer over the C ABI (one context = one GPU)."""

    def __init__(self, width: int, height: int, device: int = 0):
        self.lib = load_library()
        self.h = C.c_void_p()
        rc = self.lib.rg_create(C.byref(self.h), C.c_int(device), C.c_uint32(width), C.c_uint32(height))
        if rc != 0:
            raise RaygunError(f"rg_create failed with code {rc}: no usable CUDA device {device} (raygun_b200 has no CPU fallback)")
        self.width, self.height, self.device = width, height, device
        self.region = (0, 0, width, height)

    # ------------------------------------------------------------------ plumbing
    def _ck(self, rc):
        if rc != 0:
            raise RaygunError(self.lib.rg_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.rg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def region_size(self):
        x0, y0, x1, y1 = self.region
        return x1 - x0, y1 - y0

    def resize(self, width, height):
        self._ck(self.lib.rg_resize(self.h, C.c_uint32(width), C.c_uint32(height)))
        self.width, self.height = width, height
        self.region = (0, 0, width, height)

    def set_region(self, x0, y0, x1, y1):
        self._ck(self.lib.rg_set_region(self.h, C.c_uint32(x0), C.c_uint32(y0), C.c_uint32(x1), C.c_uint32(y1)))
        self.region = (x0, y0, x1, y1)

    # ------------------------------------------------------------------ scene load (raygun.cpp:289-290)
    def setupModelBuffers(self, sd):
        """RenderSystem::setupModelBuffers: upload the packed vertex / index / material buffers."""
        v = np.ascontiguousarray(sd.vertices, np.uint32); i = np.ascontiguousarray(sd.indices, np.uint32)
        m = np.ascontiguousarray(sd.meshes, np.uint32); mat = np.ascontiguousarray(sd.materials, np.uint32)
        self._ck(self.lib.rg_upload_geometry(self.h, _p(v), C.c_uint32(len(v)), _p(i), C.c_uint32(len(i)), _p(m), C.c_uint32(len(m))))
        self._ck(self.lib.rg_upload_materials(self.h, _p(mat), C.c_uint32(len(mat))))

    def updateMaterialBuffer(self, materials):
        mat = np.ascontiguousarray(materials, np.uint32)
        self._ck(self.lib.rg_upload_materials(self.h, _p(mat), C.c_uint32(len(mat))))

    def setupBottomLevelAS(self):
        self._ck(self.lib.rg_build_blas(self.h))

    def refitBottomLevelAS(self, mesh: int, new_vertices):
        v = np.ascontiguousarray(new_vertices, np.uint32)
        self._ck(self.lib.rg_refit_blas(self.h, C.c_uint32(mesh), _p(v)))

    def refitBottomLevelAS_device(self, mesh: int, d_ptr: int):
        """new vertex records (32 B each, same count as the mesh) already in device memory"""
        self._ck(self.lib.rg_refit_blas_device(self.h, C.c_uint32(mesh), C.c_void_p(d_ptr)))

    # ------------------------------------------------------------------ per frame (render_system.cpp:88-162)
    @staticmethod
    def pack_instances(inst_xform, inst_meta) -> np.ndarray:
        """(I,12) float32 + (I,4) uint32 -> (I,16) uint32 array of 64-byte rg_instance records."""
        n = len(inst_xform)
        raw = np.empty((n, 16), np.uint32)
        raw[:, :12] = np.ascontiguousarray(inst_xform, np.float32).view(np.uint32).reshape(n, 12)
        raw[:, 12:] = np.asarray(inst_meta, np.uint32).reshape(n, 4)
        return raw

    def setupTopLevelAS(self, instances_raw):
        raw = np.ascontiguousarray(instances_raw, np.uint32)
        self._ck(self.lib.rg_set_instances(self.h, _p(raw), C.c_uint32(len(raw))))

    def set_entities(self, entities) -> int:
        """Scene-graph walk on the device (TopLevelAS::TopLevelAS + Entity::globalTransform, acceleration_structure.cpp:55-85):
        entities = ENTITY_DTYPE array in DFS pre-order.  Builds the TLAS; returns the number of instances."""
        e = np.ascontiguousarray(entities, ENTITY_DTYPE)
        n = C.c_uint32()
        self._ck(self.lib.rg_set_entities(self.h, _p(e), C.c_uint32(len(e)), C.byref(n)))
        return n.value

    def set_entities_device(self, d_ptr: int, n_entities: int) -> int:
        n = C.c_uint32()
        self._ck(self.lib.rg_set_entities_device(self.h, C.c_void_p(d_ptr), C.c_uint32(n_entities), C.byref(n)))
        return n.value

    def physics_step_spheres(self, d_entities: int, d_bodies: int, n: int, dt: float, floor_y: float = 0.0):
        """One step of the rigid-sphere stand-in for PhysicsSystem::update on DEVICE arrays (ENTITY_DTYPE / SPHERE_BODY_DTYPE)."""
        self._ck(self.lib.rg_physics_step_spheres(self.h, C.c_void_p(d_entities), C.c_void_p(d_bodies), C.c_uint32(n), C.c_float(dt), C.c_float(floor_y)))

    def debug_read_instances(self) -> np.ndarray:
        """(I,16) uint32: the rg_instance records of the current TLAS."""
        n = C.c_uint32()
        self._ck(self.lib.rg_debug_read_instances(self.h, None, C.c_uint32(0), C.byref(n)))
        out = np.zeros((n.value, 16), np.uint32)
        if n.value:
            self._ck(self.lib.rg_debug_read_instances(self.h, _p(out), C.c_uint32(n.value), C.byref(n)))
        return out

    def updateRenderTarget(self, ubo):
        u = np.ascontiguousarray(ubo, np.uint32)
        assert u.nbytes == 192
        self._ck(self.lib.rg_set_ubo(self.h, _p(u)))

    def doRaytracing(self, flags=RG_FXAA):
        self._ck(self.lib.rg_render(self.h, C.c_uint32(flags)))

    def set_trace_scheduler(self, mode: int):
        """RG_SCHED_LANES / RG_SCHED_POOL / RG_SCHED_AUTO (default): which trace kernel runs; images are bit-identical."""
        self._ck(self.lib.rg_set_trace_scheduler(self.h, int(mode)))

    def sync(self):
        self._ck(self.lib.rg_sync(self.h))

    # ------------------------------------------------------------------ read-back
    def read_rgba8(self, out=None):
        w, h = self.region_size
        if out is None:
            out = np.empty((h, w, 4), np.uint8)
        self._ck(self.lib.rg_read_rgba8(self.h, _p(out)))
        return out

    def read_image(self, which):
        w, h = self.region_size
        out = np.empty((h, w), np.int8) if which == IMG_TRANSITIONS else np.empty((h, w, 4), np.uint16)
        self._ck(self.lib.rg_read_image(self.h, C.c_int(which), _p(out)))
        return out

    def read_ids(self):
        w, h = self.region_size
        inst = np.empty((h, w), np.uint32); prim = np.empty((h, w), np.uint32)
        self._ck(self.lib.rg_read_ids(self.h, _p(inst), _p(prim)))
        return inst, prim

    def timings(self) -> dict:
        t = RgTimings()
        self._ck(self.lib.rg_get_timings(self.h, C.byref(t)))
        d = {n: getattr(t, n) for n, _ in RgTimings._fields_}
        d["rays"] = d["rays_primary"] + d["rays_shadow"] + d["rays_reflect"] + d["rays_refract"]
        return d

    def timer_begin(self):
        self._ck(self.lib.rg_timer_begin(self.h))

    def timer_end(self) -> float:
        ms = C.c_float()
        self._ck(self.lib.rg_timer_end(self.h, C.byref(ms)))
        return float(ms.value)

    def flush_l2(self):
        self._ck(self.lib.rg_flush_l2(self.h))

    def launch_count(self) -> int:
        return int(self.lib.rg_launch_count(self.h))

    # ------------------------------------------------------------------ device-resident inputs / gather
    def set_instances_device(self, dptr: int, n: int):
        self._ck(self.lib.rg_set_instances_device(self.h, C.c_void_p(dptr), C.c_uint32(n)))

    def set_ubo_device(self, dptr: int):
        self._ck(self.lib.rg_set_ubo_device(self.h, C.c_void_p(dptr)))

    def framebuffer_device_ptr(self) -> int:
        p = C.c_void_p()
        self._ck(self.lib.rg_framebuffer_device_ptr(self.h, C.byref(p)))
        return p.value

    def gather_buffer_export(self):
        handle = (C.c_ubyte * 64)(); p = C.c_void_p()
        self._ck(self.lib.rg_gather_buffer_export(self.h, handle, C.byref(p)))
        return bytes(handle), p.value

    def gather_buffer_open(self, handle: bytes) -> int:
        buf = (C.c_ubyte * 64).from_buffer_copy(handle); p = C.c_void_p()
        self._ck(self.lib.rg_gather_buffer_open(self.h, buf, C.byref(p)))
        return p.value

    def gather_buffer_close(self, dptr: int):
        self._ck(self.lib.rg_gather_buffer_close(self.h, C.c_void_p(dptr)))

    # partitioned multi-GPU mode (trace shares dealt round-robin, G-buffer pixels stored into the owners' memory)
    def set_partition(self, rank: int, world: int):
        self._ck(self.lib.rg_set_partition(self.h, C.c_uint32(rank), C.c_uint32(world)))

    def peer_export(self) -> bytes:
        d = RgPeerDesc()
        self._ck(self.lib.rg_peer_export(self.h, C.byref(d)))
        return bytes(d)

    def peer_attach(self, peer_rank: int, desc: bytes, open_ipc: bool):
        d = RgPeerDesc.from_buffer_copy(desc)
        self._ck(self.lib.rg_peer_attach(self.h, C.c_uint32(peer_rank), C.byref(d), C.c_int(1 if open_ipc else 0)))

    def peer_detach_all(self):
        self._ck(self.lib.rg_peer_detach_all(self.h))

    def sync_error(self) -> int:
        return int(self.lib.rg_sync_error(self.h))

    def host_frame_register(self, host_ptr: int, nbytes: int) -> int:
        """Page-lock + map host memory (e.g. a shared-memory frame); returns the address for set_gather_target."""
        p = C.c_void_p()
        self._ck(self.lib.rg_host_frame_register(self.h, C.c_void_p(host_ptr), C.c_size_t(nbytes), C.byref(p)))
        return p.value

    def host_frame_unregister(self, host_ptr: int):
        self._ck(self.lib.rg_host_frame_unregister(self.h, C.c_void_p(host_ptr)))

    def set_gather_target(self, dptr):
        self._ck(self.lib.rg_set_gather_target(self.h, C.c_void_p(dptr or 0)))

    def read_gathered_rgba8(self, out=None):
        if out is None:
            out = np.empty((self.height, self.width, 4), np.uint8)
        self._ck(self.lib.rg_read_gathered_rgba8(self.h, _p(out)))
        return out

    # ------------------------------------------------------------------ parity / introspection
    def debug_blas_sort(self, mesh, n):
        keys = np.zeros(n, np.uint32); order = np.zeros(n, np.uint32)
        self._ck(self.lib.rg_debug_blas_sort(self.h, C.c_uint32(mesh), _p(keys), _p(order), C.c_uint32(n)))
        return keys, order

    def debug_tlas_sort(self, n):
        keys = np.zeros(n, np.uint32); order = np.zeros(n, np.uint32)
        self._ck(self.lib.rg_debug_tlas_sort(self.h, _p(keys), _p(order), C.c_uint32(n)))
        return keys, order

    def debug_trace_rays(self, rays8):
        rays8 = np.ascontiguousarray(rays8, np.float32).reshape(-1, 8)
        n = len(rays8)
        tuv = np.zeros((n, 3), np.float32); ip = np.zeros((n, 2), np.uint32)
        self._ck(self.lib.rg_debug_trace_rays(self.h, _p(rays8), C.c_uint32(n), _p(tuv), _p(ip)))
        return tuv, ip

    def debug_bvh_stats(self) -> dict:
        o = np.zeros(8, np.uint64)
        self._ck(self.lib.rg_debug_bvh_stats(self.h, _p(o)))
        keys = ("blas_nodes", "blas_tris", "blas_bytes", "tlas_nodes", "tlas_leaves", "tlas_bytes", "meshes", "instances")
        return dict(zip(keys, (int(v) for v in o)))

    def debug_upload_gbuffer(self, base, normal, rough):
        b, n, r = (np.ascontiguousarray(a, np.uint16) for a in (base, normal, rough))
        self._ck(self.lib.rg_debug_upload_gbuffer(self.h, _p(b), _p(n), _p(r)))

    def debug_run_post(self, flags=RG_FXAA):
        self._ck(self.lib.rg_debug_run_post(self.h, C.c_uint32(flags)))

    # ------------------------------------------------------------------ convenience
    def load_scene(self, sd):
        """finalizeLoadScene (raygun.cpp:285-290): buffers + BLAS, then the first TLAS."""
        self.setupModelBuffers(sd)
        self.setupBottomLevelAS()
        self.setupTopLevelAS(self.pack_instances(sd.inst_xform, sd.inst_meta))

    def render_frame(self, ubo, flags=RG_FXAA, instances_raw=None):
        """RenderSystem::render (render_system.cpp:88-162) minus swapchain / ImGui."""
        self.updateRenderTarget(ubo)
        if instances_raw is not None:
            self.setupTopLevelAS(instances_raw)
        self.doRaytracing(flags)
