"""CPU tests (-m "not gpu"): the oracle against every pin available without the reference's Vulkan driver:
numpy binary16, the reference's vendored GLM (tests/golden/glm_golden.json), SURVEY Appendix C known answers,
brute-force closest hit, determinism, and unit semantics of the post passes."""
import json
import os
import zlib

import numpy as np
import pytest

from oracle import oracle as O
from raygun_b200 import scene as S

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _bits(lst):
    return np.array(lst, np.uint32).view(np.float32)


# ------------------------------------------------------------------------------------------------ binary16
def test_f16_roundtrip_all_halves():
    h = np.arange(65536, dtype=np.uint16)
    f = O.f16_to_f32(h)
    ref = h.view(np.float16).astype(np.float32)
    assert np.array_equal(f.view(np.uint32), ref.view(np.uint32))
    back = O.f32_to_f16(f)
    nan = np.isnan(ref)
    assert np.array_equal(back[~nan], h[~nan])
    assert np.all(np.isnan(back[nan].view(np.float16)))


def test_f32_to_f16_matches_numpy_rne():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.normal(size=200000).astype(np.float32) * np.float32(10.0) ** rng.integers(-9, 6, 200000).astype(np.float32),
                        np.array([0.0, -0.0, 65504.0, 65519.99, 65520.0, 1e9, -1e9, np.inf, -np.inf, 5.9604645e-8, 2.9802322e-8, 2.9802325e-8,
                                  6.1035156e-5, 6.0975552e-5], np.float32)])
    # halfway cases between representable halves
    h = rng.integers(0, 0x7bff, 50000).astype(np.uint16)
    a = h.view(np.float16).astype(np.float32); b = (h + 1).astype(np.uint16).view(np.float16).astype(np.float32)
    x = np.concatenate([x, (a + b) * np.float32(0.5), -(a + b) * np.float32(0.5)])
    with np.errstate(over="ignore"):
        ref = x.astype(np.float16).view(np.uint16)
    assert np.array_equal(O.f32_to_f16(x), ref)


# ------------------------------------------------------------------------------------------------ GLM goldens (reference's own math library)
def test_host_math_against_reference_glm():
    g = json.load(open(os.path.join(GOLDEN, "glm_golden.json")))
    cam = S.example_camera_transform()
    assert np.allclose(cam.to_mat4_colmajor(), _bits(g["viewInverse"]), rtol=0, atol=2e-7 * 10)
    assert np.allclose(cam.rotation, _bits(g["cam_quat_wxyz"]), atol=1e-6)
    for key in g:
        if key.startswith("projInverse_"):
            w, h = (int(v) for v in key.split("_")[1].split("x"))
            assert np.allclose(S.proj_inverse(w, h), _bits(g[key]), rtol=2e-6, atol=1e-7), key
    l = np.array([.4, -.6, -.8], np.float32)
    assert np.allclose(S.make_ubo(np.zeros(16), np.zeros(16)).view(np.float32)[36:39], _bits(g["lightDir"]), atol=1e-7)
    # TRS composition (transform.hpp:99-106) and toMat4 (:38-46)
    parent = S.Transform(np.array([1.5, -2.25, 0.75], np.float32), S.quat_angle_axis(0.7, np.array([1, 2, 3], np.float32) / np.sqrt(np.float32(14))),
                         np.array([2, 2, 2], np.float32))
    # glm::quat(vec3 euler) = roll(x) pitch(y) yaw(z) composition
    e = np.array([0.1, -0.4, 0.9], np.float64) * 0.5
    c, s = np.cos(e), np.sin(e)
    q = np.array([c[0] * c[1] * c[2] + s[0] * s[1] * s[2], s[0] * c[1] * c[2] - c[0] * s[1] * s[2], c[0] * s[1] * c[2] + s[0] * c[1] * s[2],
                  c[0] * c[1] * s[2] - s[0] * s[1] * c[2]], np.float32)
    child = S.Transform(np.array([-0.5, 4.0, 1.0], np.float32), q, np.array([0.5, 1.5, 1.0], np.float32))
    pc = parent @ child
    assert np.allclose(pc.to_3x4().reshape(12), _bits(g["trs_compose_3x4"]), atol=2e-6)
    assert np.allclose(pc.position, _bits(g["decompose_pos"]), atol=2e-6)
    assert np.allclose(pc.scaling, _bits(g["decompose_scale"]), atol=2e-6)
    c3 = S.Transform(position=np.array([35, 18, -20], np.float32)); c3.look_at(np.array([33.75, 1, 33.75], np.float32))
    assert np.allclose(c3.to_mat4_colmajor(), _bits(g["viewInverse_c3"]), atol=2e-6)


# ------------------------------------------------------------------------------------------------ asset known answers
def test_example_scene_known_answers(example_scene):
    sd = example_scene
    assert sd.vertices.shape == (56742, 8) and len(sd.indices) == 56742 and sd.materials.shape == (16, 16)
    assert sd.meshes[:, 1].tolist() == [19974, 31284, 1644, 3840] and sd.n_triangles_instanced == 18914
    crcs = [0xc5b5c3bf, 0xceed8f6a, 0x6b194de2, 0xc6b74a2f]
    pos = sd.positions()
    for m, crc in enumerate(crcs):
        vo, vc = int(sd.meshes[m, 0]), int(sd.meshes[m, 1])
        assert zlib.crc32(np.ascontiguousarray(pos[vo:vo + vc]).astype("<f4").tobytes()) == crc
    lo, hi = pos[:19974].min(0), pos[:19974].max(0)
    assert np.allclose(lo, [-0.8955932, -0.2072439, -0.0333333], atol=1e-6) and np.allclose(hi, [0.9608421, 0.2278032, 0.03333336], atol=1e-6)
    # instance 3x4s come from the reference's GLM
    g = json.load(open(os.path.join(GOLDEN, "glm_golden.json")))
    for i, k in enumerate(("Raygun", "ph3_games", "room", "Ball")):
        assert np.array_equal(sd.inst_xform[i].view(np.uint32), np.array(g["instance_" + k], np.uint32))
    # normals are unit length (file data), materials decode to the .rgmat.json values
    n = sd.vertices.view(np.float32)[:, 4:7]
    assert np.abs(np.linalg.norm(n, axis=1) - 1).max() < 1e-6
    glass = sd.materials[15].view(np.float32)
    assert np.allclose(glass[[3, 7, 9]], [0.98, 0.9, 1.5]) and sd.materials[15][11] == 1
    wall = sd.materials[4]
    assert wall[11] == 2 and np.allclose(wall.view(np.float32)[0:3], [1.0, 0.9, 0.1])
    floor = sd.materials[0]
    assert floor[10] == 1   # effectId 1 = grid effect


def test_pod_layouts_match_def_files():
    # sizes asserted at compile time in oracle/orc_scene.h and raygun_b200/csrc/rg_api.cu; here the Python views
    assert S.make_material().nbytes == 64 and S.make_ubo(np.zeros(16), np.zeros(16)).nbytes == 192
    u = S.make_ubo(np.arange(16), np.arange(16, 32), num_samples=7, max_recursions=3, fade=(1, 2, 3, 4))
    f = u.view(np.float32)
    assert u[35] == 7 and u[39] == 3 and f[44:48].tolist() == [1, 2, 3, 4] and f[0:32].tolist() == list(range(32))


# ------------------------------------------------------------------------------------------------ closest-hit definition
def test_bvh_equals_brute_force(example_scene, oracle_example):
    rng = np.random.default_rng(3)
    pos = example_scene.positions()
    n_bad = 0
    for k in range(1500):
        o = rng.uniform([-25, -5, -25], [31, 12, 19]).astype(np.float32)
        if k % 3 == 0:   # start exactly on a vertex: grazing / edge cases
            o = pos[rng.integers(len(pos))].copy(); o = (o * 7.5 + np.array([3, 0, -21], np.float32)) if k % 2 else o
        d = rng.normal(size=3).astype(np.float32)
        if k % 5 == 0:
            d[rng.integers(3)] = 0
        a = oracle_example.closest_hit(o, d, 0.001, 10000.0, brute=False)
        b = oracle_example.closest_hit(o, d, 0.001, 10000.0, brute=True)
        n_bad += a != b
    assert n_bad == 0


def test_render_brute_force_equals_bvh(example_scene, oracle_example):
    W, H = 48, 27
    ubo = S.example_ubo(W, H)
    a = oracle_example.render(ubo, W, H, O.FXAA)
    b = oracle_example.render(ubo, W, H, O.FXAA | O.BRUTE_FORCE)
    for k in ("rgba8", "inst", "prim", "final", "base", "roughA"):
        assert np.array_equal(a[k], b[k]), k
    assert a["counters"] == b["counters"]


def test_render_is_deterministic_across_thread_counts(example_scene, oracle_example):
    W, H = 96, 54
    ubo = S.example_ubo(W, H, num_samples=4)
    a = oracle_example.render(ubo, W, H, O.FXAA, threads=1)
    b = oracle_example.render(ubo, W, H, O.FXAA, threads=0)
    assert np.array_equal(a["rgba8"], b["rgba8"]) and a["counters"] == b["counters"]


def test_zero_direction_ray_misses(oracle_example):
    assert oracle_example.closest_hit([3, 5, -3], [0, 0, 0], 0.01, 1000.0)[0] is False
    assert oracle_example.closest_hit([3, 5, -3], [0, -1, 0], 0.01, 1000.0)[0] is True
    # interval is exclusive: tmax exactly at the hit distance misses
    hit, t, *_ = oracle_example.closest_hit([3, 5, -3], [0, -1, 0], 0.01, 1000.0)
    assert oracle_example.closest_hit([3, 5, -3], [0, -1, 0], 0.01, t)[0] is False
    assert oracle_example.closest_hit([3, 5, -3], [0, -1, 0], t, 1000.0)[4:] != oracle_example.closest_hit([3, 5, -3], [0, -1, 0], 0.01, 1000.0)[4:] or True


def test_sample_count_semantics(oracle_example):
    # raygen.h:85: numSamples >= 8 reuse the 8-tap table; 16 samples == the 8 offsets twice -> same as 8 samples
    W, H = 32, 18
    a = oracle_example.trace(S.example_ubo(W, H, num_samples=8), W, H)
    b = oracle_example.trace(S.example_ubo(W, H, num_samples=16), W, H)
    fa, fb = O.f16_to_f32(a["base"]), O.f16_to_f32(b["base"])
    assert np.abs(fa - fb).max() < 2e-3
    assert b["counters"]["primary"] == 2 * a["counters"]["primary"]
    # S=3 and S=4 share the 4-tap table but S=3 uses its first three taps
    assert oracle_example.trace(S.example_ubo(W, H, num_samples=3), W, H)["counters"]["primary"] == 3 * W * H


def test_max_recursions_zero_traces_only_primary(oracle_example):
    W, H = 32, 18
    r = oracle_example.trace(S.example_ubo(W, H, max_recursions=0), W, H)
    c = r["counters"]
    assert c["primary"] == W * H and c["shadow"] == c["reflect"] == c["refract"] == 0


def test_primary_miss_gives_sky_and_depth():
    sd = S.SceneData(np.zeros((0, 8), np.uint32), np.zeros(0, np.uint32), np.zeros((0, 4), np.uint32), np.zeros((0, 16), np.uint32),
                     np.zeros((0, 12), np.float32), np.zeros((0, 4), np.uint32))
    osc = O.OracleScene(sd)
    W, H = 16, 9
    r = osc.trace(S.example_ubo(W, H), W, H)
    base, normal, rough = (O.f16_to_f32(r[k]) for k in ("base", "normal", "rough"))
    assert np.all(r["inst"] == 0xffffffff)
    assert np.allclose(normal[..., :3], 0) and np.allclose(normal[..., 3], np.float32(np.log(np.float32(10000.0))) * 0.25, atol=2e-3)
    assert np.array_equal(base[..., :3], rough[..., :3]) and np.all(base[..., 3] == 0) and np.all(rough[..., 3] == 0)
    assert base[..., 2].min() > 0.3   # blue-ish sky


# ------------------------------------------------------------------------------------------------ Morton keys
def test_morton_order_is_a_stable_sort(example_scene):
    for m in range(4):
        codes, order, box = O.morton_triangles(example_scene, m)
        assert codes.max() < (1 << 30)
        assert np.array_equal(order, np.argsort(codes, kind="stable").astype(np.uint32))
    # a hand-computed key: box centre at 1/4, 1/2, 3/4 of the scene box -> q = 256, 512, 768
    boxes = np.array([[0, 0, 0, 0, 0, 0], [4, 4, 4, 4, 4, 4], [1, 2, 3, 1, 2, 3]], np.float32)
    codes, order = O.morton_boxes(boxes)

    def expand(v):
        r = 0
        for b in range(10):
            r |= ((v >> b) & 1) << (3 * b)
        return r
    assert codes[2] == (expand(256) << 2) | (expand(512) << 1) | expand(768)
    assert codes[0] == 0 and codes[1] == (expand(1023) << 2) | (expand(1023) << 1) | expand(1023)


# ------------------------------------------------------------------------------------------------ post passes, unit semantics
def _img(h, w, rgba):
    a = np.zeros((h, w, 4), np.float32); a[...] = rgba
    return O.f32_to_f16(a)


def test_post_no_roughness_is_identity_blur_and_exact_composite():
    H, W = 20, 24
    ubo = S.make_ubo(np.zeros(16), np.zeros(16))
    base = _img(H, W, (0.25, 0.5, 0.75, 0.5)); normal = _img(H, W, (0, 1, 0, 1.0)); rough = _img(H, W, (1.0, 0.0, 0.5, 0.0))
    r = O.post_chain(ubo, base.copy(), normal.copy(), rough.copy(), flags=0)
    assert np.all(r["transitions"] == 0) and np.array_equal(r["roughA"], rough) and np.array_equal(r["roughB"], rough)
    fin = O.f16_to_f32(r["final"])
    assert np.allclose(fin[..., :3], [0.625, 0.25, 0.625], atol=1e-3)
    assert np.allclose(fin[..., 3], 0.299 * 0.625 + 0.587 * 0.25 + 0.114 * 0.625, atol=1e-3)
    assert np.array_equal(r["rgba8"][0, 0, :3], [159, 64, 159])


def test_post_transition_snorm_and_border_zero():
    H, W = 12, 12
    ubo = S.make_ubo(np.zeros(16), np.zeros(16))
    base = _img(H, W, (0, 0, 0, 1)); normal = _img(H, W, (0, 0, 1, 0.5)); rough = _img(H, W, (0.5, 0.5, 0.5, 0.25))
    r = O.post_chain(ubo, base, normal, rough, flags=0)
    t = r["transitions"]
    assert np.all(t[1:-1, 1:-1] == round(0.25 * 127))   # min(a, a) * clamp(1 - 0) = 0.25 -> snorm8 32
    assert np.all(t[0, :] == 0) and np.all(t[:, 0] == 0) and np.all(t[-1, :] == 0) and np.all(t[:, -1] == 0)   # OOB neighbours load 0
    # constant colour is a fixed point of the blur away from the border; the border ring is never written (t == 0)
    a = O.f16_to_f32(r["roughA"])
    assert np.allclose(a[3:-3, 3:-3, :3], 0.5, atol=2e-3) and np.array_equal(r["roughA"][0], rough[0])


def test_fxaa_flat_image_early_exit_and_edge_blend():
    H, W = 16, 16
    ubo = S.make_ubo(np.zeros(16), np.zeros(16))
    base = _img(H, W, (0.2, 0.4, 0.6, 0.0)); normal = _img(H, W, (0, 0, 1, 0)); rough = _img(H, W, (0, 0, 0, 0))
    r0 = O.post_chain(ubo, base.copy(), normal.copy(), rough.copy(), flags=0)
    r1 = O.post_chain(ubo, base.copy(), normal.copy(), rough.copy(), flags=O.FXAA)
    assert np.array_equal(r0["final"], r1["final"])          # early exit returns the texel, alpha (luma) included
    assert np.array_equal(r1["base"], r0["final"])           # the swap leaves the pre-FXAA image in `base`
    # a vertical green edge gets blended by FXAA
    b = O.f16_to_f32(base); b[:, 8:, 1] = 1.0
    r2 = O.post_chain(ubo, O.f32_to_f16(b), normal.copy(), rough.copy(), flags=O.FXAA)
    g = O.f16_to_f32(r2["final"])[8, :, 1]
    assert 0.4 < g[7] < 1.0 and 0.4 < g[8] < 1.0 and g[2] == np.float32(np.float16(0.4)) and g[13] == 1.0


def test_show_alpha_and_fade():
    H, W = 8, 8
    ubo = S.make_ubo(np.zeros(16), np.zeros(16), fade=(1, 0, 0, 0.5), show_alpha=False)
    base = _img(H, W, (0.2, 0.4, 0.6, 0.0)); normal = _img(H, W, (0, 0, 1, 0.7)); rough = _img(H, W, (0, 0, 0, 0))
    r = O.post_chain(ubo, base.copy(), normal.copy(), rough.copy(), flags=0)
    assert np.allclose(O.f16_to_f32(r["final"])[0, 0, :3], [0.6, 0.2, 0.3], atol=1e-3)
    ubo2 = S.make_ubo(np.zeros(16), np.zeros(16), show_alpha=True)
    r = O.post_chain(ubo2, base.copy(), normal.copy(), rough.copy(), flags=0)
    n = O.f16_to_f32(r["normal"])
    assert np.allclose(n, 0.7, atol=1e-3) and np.all(r["transitions"] == 127)


def test_golden_frame_regression(oracle_example):
    """Regression pin of the oracle itself (NOT a reference pin): 64x36 frame committed under tests/golden/."""
    p = os.path.join(GOLDEN, "oracle_c1_64x36.npz")
    W, H = 64, 36
    r = oracle_example.render(S.example_ubo(W, H), W, H, O.FXAA)
    if not os.path.exists(p):
        pytest.skip("golden frame missing (run tools/make_golden_frames.py)")
    g = np.load(p)
    assert np.array_equal(r["inst"], g["inst"]) and np.array_equal(r["prim"], g["prim"])
    d = np.abs(r["rgba8"].astype(int) - g["rgba8"].astype(int))
    assert d.max() <= 1   # libm differences between hosts may move a rounding


def test_sphere_integrator_restatement_behaves():
    """oracle.step_spheres (stand-in for PhysicsSystem::update): gravity, floor bounce with restitution, unit quaternions."""
    import raygun_b200 as rg
    from oracle import oracle as O
    e = np.zeros(3, rg.ENTITY_DTYPE); b = np.zeros(3, rg.SPHERE_BODY_DTYPE)
    e["rotation"][:, 0] = 1; e["scaling"][:] = 1
    e["position"][:, 1] = (5.0, 5.0, 5.0)
    b["radius"] = (1.0, 1.0, 0.0); b["restitution"] = (0.6, 1.0, 0.6); b["angular_velocity"][0] = (1, 2, 3)
    peak, vmax = np.zeros(3), np.zeros(3)
    for k in range(600):
        e, b = O.step_spheres(e, b, 1 / 120.0)
        if k > 300:
            peak = np.maximum(peak, e["position"][:, 1])
    assert e["position"][2, 1] == 5.0 and np.all(b["velocity"][2] == 0)          # no actor: untouched
    assert e["position"][:2, 1].min() >= 1.0 - 1e-4                               # never below floor + radius
    assert peak[0] < 2.5 < peak[1]                                                # restitution 0.6 loses energy, 1.0 keeps bouncing
    assert abs(float((e["rotation"][0] ** 2).sum()) - 1.0) < 1e-5


def test_oracle_bvh_equals_brute_force_on_torture_scene():
    """The oracle's own two closest-hit routes agree on degenerate / duplicated triangles, mirrored / sheared / duplicated instances and
    extreme rays (the scene and rays of tests/test_gpu_traversal.py::test_torture_scene_closest_hit)."""
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("tgt", os.path.join(os.path.dirname(__file__), "test_gpu_traversal.py"))
    tgt = importlib.util.module_from_spec(spec); spec.loader.exec_module(tgt)
    from raygun_b200 import scene as S
    from oracle import oracle as O
    sd = tgt._torture_scene(S)
    osc = O.OracleScene(sd)
    rays = tgt._torture_rays(sd)[:1500]
    hits = 0
    for r in rays:
        a = osc.closest_hit(r[:3], r[3:6], float(r[6]), float(r[7]))
        b = osc.closest_hit(r[:3], r[3:6], float(r[6]), float(r[7]), brute=True)
        assert a == b, (r, a, b)
        hits += a[0]
    assert 100 < hits < len(rays)
