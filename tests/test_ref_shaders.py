"""CPU tests (-m "not gpu"): the oracle restatement against the reference's OWN shader sources compiled for the CPU
(oracle/_ref/libref_shaders.so, recipe in oracle/ref_recipe/: the GLSL files are rewritten lexically -- no arithmetic is
touched -- and built against the reference's vendored GLM; the Vulkan driver's part, i.e. closest hit, image storage and the
sampler, is the same stand-in the oracle uses).  Bit-exact equality is required."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import oracle as O
from raygun_b200 import scene as S

REF_LIB = os.path.join(os.path.dirname(os.path.abspath(O.__file__)), "_ref", "libref_shaders.so")
pytestmark = pytest.mark.skipif(not os.path.exists(REF_LIB), reason="oracle/_ref not built (needs /root/reference; run __graft_entry__.build())")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def ref():
    lib = C.CDLL(REF_LIB)
    lib.ref_trace.restype = C.c_uint64
    return lib


def _ref_trace(ref, sd, ubo, W, H):
    v, i, m, mat, xf, meta = (np.ascontiguousarray(a) for a in (sd.vertices, sd.indices, sd.meshes, sd.materials, sd.inst_xform, sd.inst_meta))
    out = [np.zeros((H, W, 4), np.uint16) for _ in range(3)]
    u = np.ascontiguousarray(ubo, np.uint32)
    rays = ref.ref_trace(_p(v), len(v), _p(i), len(i), _p(m), len(m), _p(mat), len(mat), _p(xf), _p(meta), len(xf), _p(u), W, H, 0, _p(out[0]), _p(out[1]), _p(out[2]))
    return out, int(rays)


def _same_bits_or_both_nan(a, b):
    fa, fb = O.f16_to_f32(a), O.f16_to_f32(b)
    return (a == b) | (np.isnan(fa) & np.isnan(fb))


@pytest.mark.parametrize("W,H,ns,mr", [(96, 54, 1, 5), (64, 36, 4, 5), (80, 45, 2, 7), (48, 27, 1, 0)])
def test_ray_tracing_shaders_bit_exact(ref, example_scene, oracle_example, W, H, ns, mr):
    ubo = S.example_ubo(W, H, num_samples=ns, max_recursions=mr)
    (rb, rn, rr), rays = _ref_trace(ref, example_scene, ubo, W, H)
    o = oracle_example.trace(ubo, W, H, O.STRICT_IEEE)
    c = o["counters"]
    assert rays == c["primary"] + c["shadow"] + c["reflect"] + c["refract"]
    for name, a in (("base", rb), ("normal", rn), ("rough", rr)):
        ok = _same_bits_or_both_nan(a, o[name])
        assert ok.all(), f"{name}: {(~ok).sum()} of {ok.size} components differ from the reference shaders"
    # the default (NaN-free) mode differs from the literal shaders only where those produce NaN (SURVEY hazards 7, 8)
    d = oracle_example.trace(ubo, W, H, 0)
    for name, a in (("base", rb), ("normal", rn), ("rough", rr)):
        fa = O.f16_to_f32(a)
        pix_nan = np.isnan(fa).any(axis=2)
        assert (a[~pix_nan] == d[name][~pix_nan]).all(), name
    assert np.isnan(O.f16_to_f32(rr)).any() and not np.isnan(O.f16_to_f32(d["rough"])).any()


def test_ray_tracing_shaders_sphere_grid(ref):
    """Deep recursion through glass / mirror instances (config 3 in small): refraction, total internal reflection, shadow chains."""
    W, H = 64, 36
    sd, _ = S.sphere_grid_scene(4, flattened=False)
    cam = S.Transform(position=np.array([10, 7, -6], np.float32)); cam.look_at(np.array([3.75, 1, 3.75], np.float32))
    ubo = S.make_ubo(cam.to_mat4_colmajor(), S.proj_inverse(W, H), 1, 7)
    (rb, rn, rr), rays = _ref_trace(ref, sd, ubo, W, H)
    o = O.OracleScene(sd).trace(ubo, W, H, O.STRICT_IEEE)
    assert o["counters"]["refract"] > 500 and o["counters"]["skylookup"] > 0
    for name, a in (("base", rb), ("normal", rn), ("rough", rr)):
        assert _same_bits_or_both_nan(a, o[name]).all(), name


@pytest.mark.parametrize("fxaa", [1, 0])
@pytest.mark.parametrize("W,H", [(96, 54), (50, 37)])
def test_post_chain_shaders_bit_exact(ref, oracle_example, W, H, fxaa):
    ubo = S.example_ubo(W, H, num_samples=2, fade=(0.2, 0.1, 0.0, 0.3))
    g = oracle_example.trace(ubo, W, H, 0)
    want = O.post_chain(ubo, g["base"].copy(), g["normal"].copy(), g["rough"].copy(), O.FXAA if fxaa else 0)
    base, normal, rough = g["base"].copy(), g["normal"].copy(), g["rough"].copy()
    final = np.zeros((H, W, 4), np.uint16); A = np.zeros((H, W, 4), np.uint16); B = np.zeros((H, W, 4), np.uint16); T = np.zeros((H, W), np.int8)
    u = np.ascontiguousarray(ubo, np.uint32)
    ref.ref_post(_p(u), W, H, fxaa, _p(base), _p(normal), _p(rough), _p(final), _p(A), _p(B), _p(T))
    assert (want["transitions"] != 0).any(), "scene must exercise the blur"
    assert np.array_equal(T, want["transitions"])
    for name, a in (("roughA", A), ("roughB", B), ("final", final), ("base", base), ("normal", normal), ("rough", rough)):
        assert np.array_equal(a, want[name]), name


def test_post_chain_show_alpha_bit_exact(ref, oracle_example):
    W, H = 40, 24
    ubo = S.example_ubo(W, H, show_alpha=True)
    g = oracle_example.trace(ubo, W, H, 0)
    want = O.post_chain(ubo, g["base"].copy(), g["normal"].copy(), g["rough"].copy(), O.FXAA)
    base, normal, rough = g["base"].copy(), g["normal"].copy(), g["rough"].copy()
    final = np.zeros((H, W, 4), np.uint16); A = np.zeros((H, W, 4), np.uint16); B = np.zeros((H, W, 4), np.uint16); T = np.zeros((H, W), np.int8)
    ref.ref_post(_p(np.ascontiguousarray(ubo, np.uint32)), W, H, 1, _p(base), _p(normal), _p(rough), _p(final), _p(A), _p(B), _p(T))
    for name, a in (("final", final), ("base", base), ("normal", normal), ("rough", rough), ("roughA", A), ("roughB", B)):
        assert np.array_equal(a, want[name]), name
    assert np.array_equal(T, want["transitions"])
