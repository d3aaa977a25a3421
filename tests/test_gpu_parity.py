"""-m gpu: parity of the CUDA path (through the C ABI, numpy host buffers) against the CPU oracle.

Bars (BASELINE.json north_star): Morton / sort output bit-exact; primary-hit instance / primitive ids equal on
>= 99.99 % of pixels; final 8-bit image PSNR >= 50 dB with <= 2/255 error on >= 99.9 % of pixels.  On top of that the
post chain is required to be BIT-EXACT when fed the oracle's G-buffer, and a frame split into bands must equal the
unsplit frame bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ID_BAR, PSNR_BAR, ERR_BAR = 0.9999, 50.0, 0.999


def _psnr(a, b):
    mse = float(np.mean((a[..., :3].astype(np.float64) - b[..., :3].astype(np.float64)) ** 2))
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


def _check_frame(rt, ref, rg, label):
    img = rt.read_rgba8()
    inst, prim = rt.read_ids()
    idm = float(((inst == ref["inst"]) & (prim == ref["prim"])).mean())
    d = np.abs(img[..., :3].astype(int) - ref["rgba8"][..., :3].astype(int)).max(axis=2)
    psnr, frac = _psnr(img, ref["rgba8"]), float((d <= 2).mean())
    print(f"{label}: id match {idm:.6f}  PSNR {psnr:.2f} dB  <=2/255 on {frac:.5f}")
    assert idm >= ID_BAR, f"{label}: primary ids match on {idm}"
    assert psnr >= PSNR_BAR, f"{label}: PSNR {psnr}"
    assert frac >= ERR_BAR, f"{label}: {frac} of pixels within 2/255"
    return img


@pytest.fixture(scope="module")
def rgmod():
    import raygun_b200 as rg
    return rg


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


@pytest.fixture(scope="module")
def S():
    from raygun_b200 import scene
    return scene


def test_morton_sort_bit_exact_blas(rgmod, O, example_scene):
    rt = rgmod.Raytracer(64, 36)
    rt.load_scene(example_scene)
    for m in range(len(example_scene.meshes)):
        n = int(example_scene.meshes[m, 3]) // 3
        keys, order = rt.debug_blas_sort(m, n)
        codes, ref_order, _ = O.morton_triangles(example_scene, m)
        assert np.array_equal(order, ref_order), f"mesh {m}: primitive order"
        assert np.array_equal(keys, codes[ref_order]), f"mesh {m}: sorted keys"


def test_morton_sort_bit_exact_million_triangles(rgmod, O, S):
    sd, _ = S.sphere_grid_scene(28, flattened=True)     # 1 003 522 triangles in one mesh
    rt = rgmod.Raytracer(64, 36)
    rt.load_scene(sd)
    n = int(sd.meshes[0, 3]) // 3
    keys, order = rt.debug_blas_sort(0, n)
    codes, ref_order, _ = O.morton_triangles(sd, 0)
    assert np.array_equal(order, ref_order) and np.array_equal(keys, codes[ref_order])
    assert np.all(np.diff(keys.astype(np.int64)) >= 0)


def test_morton_sort_bit_exact_tlas_10k_instances(rgmod, O, S):
    balls = S.AnimatedBalls(100)
    sd = balls.scene(0.37)
    rt = rgmod.Raytracer(64, 36)
    rt.load_scene(sd)
    n = len(sd.inst_xform)
    keys, order = rt.debug_tlas_sort(n)
    pos = sd.positions()
    boxes = np.zeros((n, 6), np.float32)
    mesh_boxes = []
    for m in range(len(sd.meshes)):
        vo, vc, io, ic = (int(v) for v in sd.meshes[m])
        p = pos[vo + sd.indices[io:io + ic]]
        mesh_boxes.append(np.concatenate([p.min(0), p.max(0)]))
    for i in range(n):
        boxes[i] = O.instance_world_box(sd.inst_xform[i], mesh_boxes[int(sd.inst_meta[i, 0])])
    codes, ref_order = O.morton_boxes(boxes)
    assert np.array_equal(order, ref_order) and np.array_equal(keys, codes[ref_order])


@pytest.mark.parametrize("sched", ["lanes", "pool"])
@pytest.mark.parametrize("W,H,ns", [(640, 360, 1), (1920, 1080, 4), (100, 60, 2), (333, 187, 3)])
def test_example_scene_frame_parity(rgmod, O, S, example_scene, oracle_example, W, H, ns, sched):
    """Both trace schedulers (one context per lane / per-warp context pools) against the oracle."""
    ubo = S.example_ubo(W, H, num_samples=ns, max_recursions=5)
    rt = rgmod.Raytracer(W, H)
    rt.set_trace_scheduler(rgmod.RG_SCHED_POOL if sched == "pool" else rgmod.RG_SCHED_LANES)
    rt.load_scene(example_scene)
    rt.render_frame(ubo, rgmod.RG_FXAA | rgmod.RG_DEBUG_IDS)
    ref = oracle_example.render(ubo, W, H, O.FXAA)
    _check_frame(rt, ref, rgmod, f"example {W}x{H} S={ns}")
    tm = rt.timings()
    c = ref["counters"]
    for a, b in (("rays_primary", "primary"), ("rays_shadow", "shadow"), ("rays_reflect", "reflect"), ("rays_refract", "refract"), ("sky_lookups", "skylookup")):
        assert abs(tm[a] - c[b]) <= max(4, 2e-4 * c[b]), (a, tm[a], c[b])
    assert tm["rays_primary"] == W * H * ns
    # intermediate images: transitions nearly identical, G-buffer normals tight
    nrm = np.abs(O.f16_to_f32(rt.read_image(rgmod.IMG_NORMAL)) - O.f16_to_f32(ref["normal"]))
    assert float((nrm <= 2e-3).mean()) >= 0.9999
    assert float((rt.read_image(rgmod.IMG_TRANSITIONS) == ref["transitions"]).mean()) >= 0.999


@pytest.mark.parametrize("flags_extra", [0, "nofxaa", "srgb"])
def test_post_chain_bit_exact_on_oracle_gbuffer(rgmod, O, S, example_scene, oracle_example, flags_extra):
    W, H = 1920, 1080
    ubo = S.example_ubo(W, H, num_samples=1, fade=(0.1, 0.2, 0.3, 0.25) if flags_extra == "srgb" else (0, 0, 0, 0))
    oflags = {0: O.FXAA, "nofxaa": 0, "srgb": O.FXAA | O.SRGB8}[flags_extra]
    gflags = {0: rgmod.RG_FXAA, "nofxaa": 0, "srgb": rgmod.RG_FXAA | rgmod.RG_SRGB8}[flags_extra]
    ref = oracle_example.render(ubo, W, H, oflags)
    rt = rgmod.Raytracer(W, H)
    rt.updateRenderTarget(ubo)
    gb = ref["gbuffer"]
    rt.debug_upload_gbuffer(gb["base"], gb["normal"], gb["rough"])
    rt.debug_run_post(gflags)
    for name, which in (("final", rgmod.IMG_FINAL), ("base", rgmod.IMG_BASE), ("roughA", rgmod.IMG_ROUGH_A), ("roughB", rgmod.IMG_ROUGH_B),
                        ("normal", rgmod.IMG_NORMAL), ("rough", rgmod.IMG_ROUGH)):
        assert np.array_equal(rt.read_image(which), ref[name]), name
    assert np.array_equal(rt.read_image(rgmod.IMG_TRANSITIONS), ref["transitions"])
    got = rt.read_rgba8()
    if flags_extra == "srgb":   # powf differs by an ulp between libm and CUDA: allow one 8-bit step on rgb
        assert np.abs(got.astype(int) - ref["rgba8"].astype(int)).max() <= 1
    else:
        assert np.array_equal(got, ref["rgba8"])


def test_show_alpha_debug_path(rgmod, O, S, example_scene, oracle_example):
    W, H = 160, 90
    ubo = S.example_ubo(W, H, show_alpha=True)
    ref = oracle_example.render(ubo, W, H, O.FXAA)
    rt = rgmod.Raytracer(W, H)
    rt.updateRenderTarget(ubo)
    gb = ref["gbuffer"]
    rt.debug_upload_gbuffer(gb["base"], gb["normal"], gb["rough"])
    rt.debug_run_post(rgmod.RG_FXAA)
    for name, which in (("final", rgmod.IMG_FINAL), ("base", rgmod.IMG_BASE), ("normal", rgmod.IMG_NORMAL), ("roughA", rgmod.IMG_ROUGH_A)):
        assert np.array_equal(rt.read_image(which), ref[name]), name
    assert np.array_equal(rt.read_image(rgmod.IMG_TRANSITIONS), ref["transitions"])


@pytest.mark.parametrize("split", ["columns", "rows"])
def test_band_split_is_bit_identical_to_full_frame(rgmod, S, example_scene, split):
    """Multi-GPU path on one GPU: three bands rendered one after the other into one gather buffer == the unsplit frame."""
    from raygun_b200.parallel import band_region
    W, H = 480, 270
    ubo = S.example_ubo(W, H, num_samples=2)
    full = rgmod.Raytracer(W, H)
    full.load_scene(example_scene)
    full.render_frame(ubo, rgmod.RG_FXAA)
    want = full.read_rgba8()
    _, target = full.gather_buffer_export()
    for r in range(3):
        band = rgmod.Raytracer(W, H)
        band.set_region(*band_region(W, H, r, 3, split))
        band.load_scene(example_scene)
        band.set_gather_target(target)
        band.render_frame(ubo, rgmod.RG_FXAA)
        band.sync()
        x0, y0, x1, y1 = band.region
        assert np.array_equal(band.read_rgba8(), want[y0:y1, x0:x1])
        band.close()
    assert np.array_equal(full.read_gathered_rgba8(), want)


@pytest.mark.parametrize("world,split,sched", [(2, "columns", "auto"), (3, "rows", "pool"), (4, "columns", "lanes"), (2, "rows", "pool")])
def test_partitioned_mode_is_bit_identical_to_full_frame(rgmod, S, example_scene, world, split, sched):
    """Partitioned multi-GPU mode with `world` contexts on one GPU: every context traces its round-robin share of tiles and
    stores the pixels into the owners' G-buffers (peer pointers), device-side barriers order trace / post / next frame;
    the gathered frame must equal the single-context frame bit for bit, for two consecutive frames."""
    from raygun_b200.parallel import band_region, attach_partition_in_process
    W, H = 400, 230
    full = rgmod.Raytracer(W, H)
    full.load_scene(example_scene)
    _, target = full.gather_buffer_export()
    rts = []
    for r in range(world):
        rt = rgmod.Raytracer(W, H)
        rt.set_region(*band_region(W, H, r, world, split))
        rt.set_trace_scheduler({"auto": rgmod.RG_SCHED_AUTO, "pool": rgmod.RG_SCHED_POOL, "lanes": rgmod.RG_SCHED_LANES}[sched])
        rt.load_scene(example_scene)
        rt.set_gather_target(target)
        rts.append(rt)
    attach_partition_in_process(rts)
    for ns in (1, 2, 2, 2) if sched == "auto" else (1, 2):   # auto: frames 3 and 4 are the two probe frames
        ubo = S.example_ubo(W, H, num_samples=ns)
        full.render_frame(ubo, rgmod.RG_FXAA | rgmod.RG_NO_GATHER)
        want = full.read_rgba8()
        for rt in rts:
            rt.render_frame(ubo, rgmod.RG_FXAA)
        for rt in rts:
            rt.sync()
            assert rt.sync_error() == 0
        for rt in rts:
            x0, y0, x1, y1 = rt.region
            assert np.array_equal(rt.read_rgba8(), want[y0:y1, x0:x1])
        assert np.array_equal(full.read_gathered_rgba8(), want)
        total = sum(rt.timings()["rays_primary"] for rt in rts)
        assert total == W * H * ns       # every pixel traced exactly once: no overdraw
    for rt in rts:
        rt.close()


def test_sphere_grid_recursion_8_parity(rgmod, O, S):
    """BASELINE config 3 in small: 6x6 mirror / glass spheres, instanced AND flattened, maxRecursions 8."""
    W, H = 320, 180
    for flat in (False, True):
        sd, vi = S.sphere_grid_scene(6, flattened=flat)
        ubo = S.make_ubo(vi, S.proj_inverse(W, H), 1, 8)
        # look at the small grid
        cam = S.Transform(position=np.array([14, 9, -8], np.float32)); cam.look_at(np.array([6.25, 1, 6.25], np.float32))
        ubo = S.make_ubo(cam.to_mat4_colmajor(), S.proj_inverse(W, H), 1, 8)
        ref = O.OracleScene(sd).render(ubo, W, H, O.FXAA)
        for sched in (rgmod.RG_SCHED_LANES, rgmod.RG_SCHED_POOL):
            rt = rgmod.Raytracer(W, H)
            rt.set_trace_scheduler(sched)
            rt.load_scene(sd)
            rt.render_frame(ubo, rgmod.RG_FXAA | rgmod.RG_DEBUG_IDS)
            _check_frame(rt, ref, rgmod, f"spheres6 flat={flat} sched={sched}")
            assert rt.timings()["trace_scheduler"] == sched
            rt.close()
        assert ref["counters"]["refract"] > 1000 and ref["counters"]["reflect"] > 10000


def test_auto_scheduler_switches_without_changing_the_image(rgmod, S):
    """RG_SCHED_AUTO times both trace kernels on consecutive frames and keeps the faster one: every frame must be bit-identical."""
    W, H = 480, 270
    sd, vi = S.sphere_grid_scene(8)
    ubo = S.make_ubo(vi, S.proj_inverse(W, H), 2, 8)
    rt = rgmod.Raytracer(W, H)
    rt.load_scene(sd)
    frames, used = [], []
    for _ in range(6):
        rt.render_frame(ubo, rgmod.RG_FXAA)
        frames.append(rt.read_rgba8().copy())
        used.append(rt.timings()["trace_scheduler"])
    assert set(used) == {rgmod.RG_SCHED_LANES, rgmod.RG_SCHED_POOL}, used   # both were probed
    for f in frames[1:]:
        assert np.array_equal(f, frames[0])
    rt.close()


def test_per_frame_tlas_rebuild_animated(rgmod, O, S):
    """BASELINE config 4 in small: 20x20 bouncing balls, TLAS rebuilt from new transforms every frame."""
    W, H = 256, 144
    balls = S.AnimatedBalls(20)
    rt = rgmod.Raytracer(W, H)
    rt.load_scene(balls.scene(0.0))
    cam = S.Transform(position=np.array([40, 25, -15], np.float32)); cam.look_at(np.array([24, 1, 24], np.float32))
    ubo = S.make_ubo(cam.to_mat4_colmajor(), S.proj_inverse(W, H), 1, 4)
    osc = O.OracleScene(balls.scene(0.0))
    for frame in (1, 7):
        t = frame / 60.0
        xf = balls.instances(t)
        rt.render_frame(ubo, rgmod.RG_FXAA | rgmod.RG_DEBUG_IDS, rt.pack_instances(xf, balls.meta))
        osc.set_instances(xf, balls.meta)
        _check_frame(rt, osc.render(ubo, W, H, O.FXAA), rgmod, f"balls t={t:.3f}")


def test_blas_refit_matches_rebuild(rgmod, O, S, example_scene):
    """Extension for config 4: refit after a vertex wobble gives the same hits as a fresh build (oracle = fresh build)."""
    W, H = 256, 144
    sd = example_scene
    ubo = S.example_ubo(W, H)
    rt = rgmod.Raytracer(W, H)
    rt.load_scene(sd)
    vo, vc = int(sd.meshes[3, 0]), int(sd.meshes[3, 1])     # the ball
    v = sd.vertices[vo:vo + vc].copy()
    p = v.view(np.float32)
    p[:, 0:3] *= (1.0 + 0.25 * np.sin(7.0 * p[:, 1:2])).astype(np.float32)
    rt.refitBottomLevelAS(3, v)
    rt.render_frame(ubo, rgmod.RG_FXAA | rgmod.RG_DEBUG_IDS, rt.pack_instances(sd.inst_xform, sd.inst_meta))
    sd2 = S.SceneData(sd.vertices.copy(), sd.indices, sd.meshes, sd.materials, sd.inst_xform, sd.inst_meta)
    sd2.vertices[vo:vo + vc] = v
    _check_frame(rt, O.OracleScene(sd2).render(ubo, W, H, O.FXAA), rgmod, "refit ball")


def test_edge_cases(rgmod, O, S, example_scene):
    W, H = 64, 36
    ubo = S.example_ubo(W, H)
    # no instances at all: sky everywhere
    rt = rgmod.Raytracer(W, H)
    rt.load_scene(example_scene)
    rt.render_frame(ubo, rgmod.RG_FXAA | rgmod.RG_DEBUG_IDS, np.zeros((0, 16), np.uint32))
    inst, _ = rt.read_ids()
    assert np.all(inst == 0xffffffff)
    empty = S.SceneData(np.zeros((0, 8), np.uint32), np.zeros(0, np.uint32), np.zeros((0, 4), np.uint32), np.zeros((0, 16), np.uint32),
                        np.zeros((0, 12), np.float32), np.zeros((0, 4), np.uint32))
    ref = O.OracleScene(empty).render(ubo, W, H, O.FXAA)
    assert np.abs(rt.read_rgba8()[..., :3].astype(int) - ref["rgba8"][..., :3].astype(int)).max() <= 1
    # a mesh with zero triangles + a single-triangle mesh + one instance of each
    v = np.zeros((3, 8), np.uint32); f = v.view(np.float32)
    f[0, 0:3] = (0, 0, -5); f[1, 0:3] = (4, 0, -5); f[2, 0:3] = (0, 4, -5); f[:, 4:7] = (0, 0, 1)
    sd = S.SceneData(v, np.array([0, 1, 2], np.uint32), np.array([(0, 0, 0, 0), (0, 3, 0, 3)], np.uint32), np.stack([S.make_material(diffuse=(0, 1, 0))]),
                     np.stack([S.Transform().to_3x4().reshape(12)] * 2), np.array([(0, 0, 0, 0), (1, 0, 0, 0)], np.uint32))
    cam = S.Transform(position=np.array([1, 1, 2], np.float32)); cam.look_at(np.array([1, 1, -5], np.float32))
    ubo2 = S.make_ubo(cam.to_mat4_colmajor(), S.proj_inverse(W, H))
    rt2 = rgmod.Raytracer(W, H)
    rt2.load_scene(sd)
    rt2.render_frame(ubo2, rgmod.RG_DEBUG_IDS)
    ref2 = O.OracleScene(sd).render(ubo2, W, H, 0)
    inst2, prim2 = rt2.read_ids()
    assert np.array_equal(inst2, ref2["inst"]) and np.array_equal(prim2, ref2["prim"]) and (inst2 == 1).sum() > 50
    # invalid arguments are rejected with a message, not a crash
    bad = np.zeros(3, rgmod.ENTITY_DTYPE); bad["parent"] = (-1, 2, 0)          # child listed before its parent
    with pytest.raises(rgmod.RaygunError):
        rt2.set_entities(bad)
    deep = np.zeros(80, rgmod.ENTITY_DTYPE); deep["parent"] = np.arange(-1, 79)  # a chain 80 levels deep
    with pytest.raises(rgmod.RaygunError):
        rt2.set_entities(deep)
    with pytest.raises(rgmod.RaygunError):
        rt2.updateRenderTarget(S.make_ubo(np.zeros(16), np.zeros(16), num_samples=0))
    with pytest.raises(rgmod.RaygunError):
        rt2.setupTopLevelAS(np.array([[0] * 12 + [9, 0, 0, 0]], np.uint32))
    with pytest.raises(rgmod.RaygunError):
        rt2.set_region(10, 10, 5, 20)


@pytest.mark.parametrize("ns,mr", [(16, 5), (8, 5), (5, 3), (4, 0), (1, 8), (2, 1)])
def test_sample_tables_and_recursion_limits(rgmod, O, S, example_scene, oracle_example, ns, mr):
    """raygen.h:37-67: numSamples >= 5 uses the 8-offset table (16 = twice, BASELINE config 5), 3 and 4 the rotated grid; and the
    recursion guard `recDepth < maxRecursions` from 0 (no secondary rays at all) to the library maximum 8.  Bit-exact image, exact
    ray counters, both schedulers."""
    W, H = 160, 90
    ubo = S.example_ubo(W, H, num_samples=ns, max_recursions=mr)
    ref = oracle_example.render(ubo, W, H, O.FXAA)
    for sched in (rgmod.RG_SCHED_LANES, rgmod.RG_SCHED_POOL):
        rt = rgmod.Raytracer(W, H)
        rt.set_trace_scheduler(sched)
        rt.load_scene(example_scene)
        rt.render_frame(ubo, rgmod.RG_FXAA | rgmod.RG_DEBUG_IDS)
        _check_frame(rt, ref, rgmod, f"S={ns} maxRec={mr} sched={sched}")
        tm, c = rt.timings(), ref["counters"]
        assert tm["rays_primary"] == W * H * ns
        for a, b in (("rays_shadow", "shadow"), ("rays_reflect", "reflect"), ("rays_refract", "refract")):
            assert abs(tm[a] - c[b]) <= max(2, 2e-4 * c[b]), (a, tm[a], c[b])
        if mr == 0:
            assert tm["rays_shadow"] == tm["rays_reflect"] == tm["rays_refract"] == 0
        rt.close()


def test_strict_ieee_switch_only_changes_diffuse_materials(rgmod, O, S, example_scene, oracle_example):
    """SURVEY hazard 8: with the 0/0 kept, NaN appears exactly where the oracle (strict) has it."""
    W, H = 160, 90
    ubo = S.example_ubo(W, H)
    rt = rgmod.Raytracer(W, H)
    rt.load_scene(example_scene)
    rt.render_frame(ubo, rgmod.RG_STRICT_IEEE)
    g = O.f16_to_f32(rt.read_image(rgmod.IMG_ROUGH))
    ref = oracle_example.trace(ubo, W, H, O.STRICT_IEEE)
    r = O.f16_to_f32(ref["rough"])
    assert np.isnan(r).any()
    assert float((np.isnan(g) == np.isnan(r)).mean()) >= 0.9999


def _random_scene_graph(rng, n, n_models):
    """Random tree flattened in DFS pre-order -> (ENTITY_DTYPE array, host-shim arrays)."""
    import raygun_b200 as rg
    children = {-1: []}
    for k in range(n):
        parent = -1 if k == 0 or rng.random() < 0.1 else int(rng.integers(0, k))
        children.setdefault(parent, []).append(k); children.setdefault(k, [])
    order = []
    def visit(k):
        order.append(k)
        for c in children[k]:
            visit(c)
    for r in children[-1]:
        visit(r)
    new_index = {old: i for i, old in enumerate(order)}
    parent_of = {c: p for p, cs in children.items() for c in cs}
    ents = np.zeros(n, rg.ENTITY_DTYPE)
    for i, old in enumerate(order):
        p = parent_of[old]
        q = rng.normal(size=4).astype(np.float32); q /= np.float32(np.sqrt((q * q).sum()))
        e = ents[i]
        e["parent"] = -1 if p < 0 else new_index[p]
        e["position"] = rng.uniform(-5, 5, 3).astype(np.float32)
        e["rotation"] = q
        e["scaling"] = rng.uniform(0.3, 2.0, 3).astype(np.float32)
        if rng.random() < 0.05:
            e["scaling"][int(rng.integers(0, 3))] = 0.0      # zero volume: prunes the subtree
        vis = rng.random() > 0.07
        has = rng.random() < 0.7
        e["flags"] = (rg.RG_ENTITY_VISIBLE if vis else 0) | (rg.RG_ENTITY_HAS_MODEL if has else 0)
        m = int(rng.integers(0, n_models))
        e["mesh"], e["vtx_off"], e["idx_off"], e["mat_off"] = m, 100 * m, 300 * m, 5 * m
    return ents


def test_device_scene_graph_walk_matches_host_walk(rgmod, example_scene):
    """rg_set_entities (TRS composition, pruning, compaction on the device) against the C++ host shim's
    Raytracer::gatherInstances over Scene / Entity / Transform (the mirror of acceleration_structure.cpp:55-85): bit-exact."""
    import ctypes as C
    import os
    host = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(rgmod.__file__)), "libraygun_host.so"))
    rng = np.random.default_rng(7)
    rt = rgmod.Raytracer(64, 36)
    rt.load_scene(example_scene)
    for n in (1, 5, 300, 4000):
        ents = _random_scene_graph(rng, n, 4)
        models = np.array([[100 * m, 1, 300 * m, 3, 5 * m, 1] for m in range(4)], np.uint32)
        he = np.stack([ents["parent"], np.where(ents["flags"] & rgmod.RG_ENTITY_HAS_MODEL, ents["mesh"].astype(np.int32), -1),
                       (ents["flags"] & rgmod.RG_ENTITY_VISIBLE).astype(np.int32)], 1).astype(np.int32)
        trs = np.concatenate([ents["position"], ents["rotation"], ents["scaling"]], 1).astype(np.float32)
        want = np.zeros((n, 16), np.uint32); cnt = C.c_uint32()
        p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        he = np.ascontiguousarray(he); trs = np.ascontiguousarray(trs)
        assert host.rgh_gather_instances(p(models), 4, p(he), p(trs), n, p(want), C.byref(cnt)) == 0
        got_n = rt.set_entities(ents)
        got = rt.debug_read_instances()
        assert got_n == cnt.value == len(got), (n, got_n, cnt.value)
        assert np.array_equal(got, want[:cnt.value]), f"{n} entities: instance records differ"
    rt.close()


def test_rigid_sphere_integrator_bit_exact_and_device_resident_frame(rgmod, O, S):
    """the stand-in for PhysicsSystem::update on device arrays against its numpy restatement (bit-exact over 200
    steps), then an animated frame with no host data: step -> rg_set_entities_device -> render equals the host-fed frame."""
    import torch
    balls = S.AnimatedBalls(12)
    ents = balls.entities(0.0)
    n = len(ents)
    rng = np.random.default_rng(3)
    bodies = np.zeros(n, rgmod.SPHERE_BODY_DTYPE)
    bodies["radius"][1:-1] = 1.0                       # root and floor: no actor
    bodies["restitution"][:] = 0.6
    bodies["velocity"][1:-1, 1] = rng.uniform(-2, 6, n - 2).astype(np.float32)
    bodies["angular_velocity"][1:-1] = rng.uniform(-3, 3, (n - 2, 3)).astype(np.float32)
    ents["position"][1:-1, 1] = rng.uniform(1.0, 6.0, n - 2).astype(np.float32)
    W, H = 256, 144
    rt = rgmod.Raytracer(W, H)
    rt.load_scene(balls.scene(0.0))
    d_e = torch.from_numpy(ents.view(np.uint8).copy()).cuda()
    d_b = torch.from_numpy(bodies.view(np.uint8).copy()).cuda()
    e_ref, b_ref = ents, bodies
    for _ in range(200):
        rt.physics_step_spheres(d_e.data_ptr(), d_b.data_ptr(), n, 1.0 / 60.0, 0.0)
        e_ref, b_ref = O.step_spheres(e_ref, b_ref, 1.0 / 60.0, 0.0)
    rt.sync()
    e_gpu = d_e.cpu().numpy().view(rgmod.ENTITY_DTYPE)
    b_gpu = d_b.cpu().numpy().view(rgmod.SPHERE_BODY_DTYPE)
    assert np.array_equal(e_gpu.view(np.uint32), e_ref.view(np.uint32))
    assert np.array_equal(b_gpu.view(np.uint32), b_ref.view(np.uint32))
    assert float(e_ref["position"][1:-1, 1].min()) >= 1.0 - 1e-3          # nobody fell through the floor
    ubo = S.make_ubo(balls.view_inverse, S.proj_inverse(W, H), 1, 5)
    rt.updateRenderTarget(ubo)
    assert rt.set_entities_device(d_e.data_ptr(), n) == n - 1
    rt.doRaytracing(rgmod.RG_FXAA)
    a = rt.read_rgba8().copy()
    assert rt.set_entities(e_ref) == n - 1
    rt.doRaytracing(rgmod.RG_FXAA)
    assert np.array_equal(rt.read_rgba8(), a)
    rt.close()


@pytest.mark.parametrize("workload", ["c3", "c5"])
def test_full_size_configs_schedulers_agree_and_frames_repeat(rgmod, workload):
    """BASELINE configs 3 and 5 at FULL size, where the CPU oracle takes too long for the GPU suite: size-independent properties --
    the two trace schedulers (different work order, different memory layout of the ray trees) must produce the same image bit for
    bit and the same ray counters, a second frame must repeat the first (per-sample scratch, tile order and work counters reset
    correctly), and every pixel is traced exactly numSamples times."""
    import bench
    desc, W, H, sd, ubo = bench.make_workload(workload)
    ns = int(ubo[35])
    frames, counters = [], []
    for sched in (rgmod.RG_SCHED_LANES, rgmod.RG_SCHED_POOL):
        rt = rgmod.Raytracer(W, H)
        rt.set_trace_scheduler(sched)
        rt.load_scene(sd)
        for _ in range(2):
            rt.render_frame(ubo, rgmod.RG_FXAA)
            frames.append(rt.read_rgba8().copy())
            tm = rt.timings()
            counters.append(tuple(tm[k] for k in ("rays_primary", "rays_shadow", "rays_reflect", "rays_refract", "sky_lookups")))
        rt.close()
    assert all(np.array_equal(f, frames[0]) for f in frames[1:])
    assert all(c == counters[0] for c in counters[1:]), counters
    assert counters[0][0] == W * H * ns
    assert len(np.unique(frames[0].reshape(-1, 4), axis=0)) > 1000      # a real image, not a constant


@pytest.mark.parametrize("n", [50, 100])
def test_large_tlas_frame_parity(rgmod, O, S, n):
    """BASELINE config 4's acceleration-structure path: 2 501 and 10 001 instances are above kTlasFusedMax (1 024), so the TLAS is built
    by the multi-kernel builder (k_prepare_instances, k_inst_boxes, radix sort, k_hierarchy, k_refit_binary, k_collapse_all) and then
    TRAVERSED: primary ids and the image are compared with the oracle for two animation frames, under both trace schedulers
    (replaces TopLevelAS, raygun/render/acceleration_structure.cpp:55-138, called every frame from raytracer.cpp:76-85)."""
    W, H = 480, 270
    balls = S.AnimatedBalls(n)
    sd0 = balls.scene(0.0)
    assert len(sd0.inst_xform) > 1024
    ubo = S.make_ubo(balls.view_inverse, S.proj_inverse(W, H), 1, 3)
    osc = O.OracleScene(sd0)
    for sched in (rgmod.RG_SCHED_LANES, rgmod.RG_SCHED_POOL):
        rt = rgmod.Raytracer(W, H)
        rt.set_trace_scheduler(sched)
        rt.load_scene(sd0)
        for frame in (3, 11):
            xf = balls.instances(frame / 60.0)
            rt.render_frame(ubo, rgmod.RG_FXAA | rgmod.RG_DEBUG_IDS, rt.pack_instances(xf, balls.meta))
            osc.set_instances(xf, balls.meta)
            ref = osc.render(ubo, W, H, O.FXAA)
            _check_frame(rt, ref, rgmod, f"balls {n}x{n} frame {frame} sched {sched}")
            inst, _ = rt.read_ids()
            assert len(np.unique(inst)) > 200      # the frame really shows hundreds of different instances
        rt.close()


@pytest.mark.parametrize("workload", ["c3", "c5"])
def test_full_size_frames_against_the_oracle(rgmod, O, workload):
    """BASELINE configs 3 (785 instances, 1 003 522 triangles, maxRecursions 8, 1920x1080) and 5 (3840x2160, numSamples 4) at FULL
    size against the CPU oracle (a few seconds per frame on the box's host cores): ids, PSNR, error bar and every ray counter."""
    import bench
    desc, W, H, sd, ubo = bench.make_workload(workload)
    ref = O.OracleScene(sd).render(ubo, W, H, O.FXAA)
    rt = rgmod.Raytracer(W, H)
    rt.load_scene(sd)
    rt.render_frame(ubo, rgmod.RG_FXAA | rgmod.RG_DEBUG_IDS)
    _check_frame(rt, ref, rgmod, f"full-size {workload}")
    tm, c = rt.timings(), ref["counters"]
    for k, o in (("rays_primary", "primary"), ("rays_shadow", "shadow"), ("rays_reflect", "reflect"), ("rays_refract", "refract")):
        assert abs(tm[k] - c[o]) <= 2e-4 * max(c[o], 1), (k, tm[k], c[o])
    rt.close()


def test_blas_refit_from_device_memory(rgmod, O, S, example_scene):
    """rg_refit_blas_device: the animated vertices are already in HBM (no 32 B / vertex host copy per frame); same result as the
    host-pointer refit and as a fresh build (the oracle)."""
    import torch
    W, H = 256, 144
    sd = example_scene
    ubo = S.example_ubo(W, H)
    vo, vc = int(sd.meshes[3, 0]), int(sd.meshes[3, 1])     # the ball
    v = sd.vertices[vo:vo + vc].copy()
    p = v.view(np.float32)
    p[:, 0:3] *= (1.0 + 0.25 * np.sin(7.0 * p[:, 1:2])).astype(np.float32)
    frames = []
    for device_ptr in (False, True):
        rt = rgmod.Raytracer(W, H)
        rt.load_scene(sd)
        if device_ptr:
            dv = torch.from_numpy(v.view(np.int32).copy()).cuda()
            torch.cuda.synchronize()
            rt.refitBottomLevelAS_device(3, dv.data_ptr())
        else:
            rt.refitBottomLevelAS(3, v)
        rt.render_frame(ubo, rgmod.RG_FXAA | rgmod.RG_DEBUG_IDS, rt.pack_instances(sd.inst_xform, sd.inst_meta))
        frames.append(rt.read_rgba8().copy())
        if device_ptr:
            sd2 = S.SceneData(sd.vertices.copy(), sd.indices, sd.meshes, sd.materials, sd.inst_xform, sd.inst_meta)
            sd2.vertices[vo:vo + vc] = v
            _check_frame(rt, O.OracleScene(sd2).render(ubo, W, H, O.FXAA), rgmod, "refit ball (device pointer)")
        rt.close()
    assert np.array_equal(frames[0], frames[1])


def test_resize_sample_change_and_material_edit_equal_a_fresh_context(rgmod, S, example_scene):
    """RenderSystem::reload (render_system.cpp:77-78: a new Raytracer at the new window size), numSamples changed from the UI
    (render_system.cpp:264) and the material editor's re-upload (gpu_material.cpp:91-93) on a LIVE context must give exactly what a
    fresh context gives."""
    def fresh(W, H, ns, sd):
        rt = rgmod.Raytracer(W, H)
        rt.load_scene(sd)
        rt.render_frame(S.example_ubo(W, H, num_samples=ns), rgmod.RG_FXAA)
        img = rt.read_rgba8().copy()
        rt.close()
        return img
    rt = rgmod.Raytracer(64, 36)
    rt.load_scene(example_scene)
    rt.render_frame(S.example_ubo(64, 36), rgmod.RG_FXAA)
    assert np.array_equal(rt.read_rgba8(), fresh(64, 36, 1, example_scene))
    rt.resize(200, 120)
    for ns in (1, 4, 2):
        rt.render_frame(S.example_ubo(200, 120, num_samples=ns), rgmod.RG_FXAA)
        assert np.array_equal(rt.read_rgba8(), fresh(200, 120, ns, example_scene)), ns
    mats = example_scene.materials.copy()
    mats[:, 0:3] = np.array([0.2, 0.7, 0.3], np.float32).view(np.uint32)      # every material turns green
    rt.updateMaterialBuffer(mats)
    rt.render_frame(S.example_ubo(200, 120, num_samples=2), rgmod.RG_FXAA)
    edited = S.SceneData(example_scene.vertices, example_scene.indices, example_scene.meshes, mats, example_scene.inst_xform, example_scene.inst_meta)
    assert np.array_equal(rt.read_rgba8(), fresh(200, 120, 2, edited))
    rt.close()


def test_api_guards_against_stale_state(rgmod, O, S, example_scene):
    """Defensive behaviour of the C ABI (no reference counterpart): a resize drops the gather target (old stride), a partitioned
    context refuses to re-allocate its images under its peers, rg_set_ubo_device validates like rg_set_ubo, and a material with
    rayConsumption 0 (gpu_material.def documents 1..5) is clamped identically in the kernels and the oracle."""
    import torch
    W, H = 128, 72
    sd = example_scene
    rt = rgmod.Raytracer(W, H)
    rt.load_scene(sd)
    ubo = S.example_ubo(W, H)
    _, own = rt.gather_buffer_export()
    rt.set_gather_target(own)
    rt.render_frame(ubo, rgmod.RG_FXAA)
    a = rt.read_gathered_rgba8().copy()
    assert np.array_equal(a, rt.read_rgba8())
    rt.resize(96, 54)                                 # frees the gather buffer; the stale target must not be written
    ubo2 = S.example_ubo(96, 54)
    rt.render_frame(ubo2, rgmod.RG_FXAA)
    rt.sync()
    _, own = rt.gather_buffer_export()
    rt.set_gather_target(own)
    rt.render_frame(ubo2, rgmod.RG_FXAA)
    assert np.array_equal(rt.read_gathered_rgba8(), rt.read_rgba8())
    # device-side UBO with numSamples 0: rejected, the previous block stays in force
    bad = ubo2.copy(); bad[35] = 0
    d_bad = torch.from_numpy(bad.view(np.int32).copy()).cuda(); torch.cuda.synchronize()
    with pytest.raises(rgmod.RaygunError):
        rt.set_ubo_device(d_bad.data_ptr())
    before = rt.read_rgba8().copy()
    rt.doRaytracing(rgmod.RG_FXAA)
    assert np.array_equal(rt.read_rgba8(), before)
    # partitioned contexts must leave the mode before their images move
    rt.set_partition(0, 2)
    with pytest.raises(rgmod.RaygunError):
        rt.set_region(0, 0, 48, 54)
    with pytest.raises(rgmod.RaygunError):
        rt.resize(64, 36)
    rt.set_partition(0, 1)
    rt.set_region(0, 0, 48, 54)
    rt.close()
    # rayConsumption 0 on every material
    mats = sd.materials.copy()
    mats[:, 11] = 0
    sd0 = S.SceneData(sd.vertices, sd.indices, sd.meshes, mats, sd.inst_xform, sd.inst_meta)
    rt = rgmod.Raytracer(W, H)
    rt.load_scene(sd0)
    rt.render_frame(ubo, rgmod.RG_FXAA | rgmod.RG_DEBUG_IDS)
    _check_frame(rt, O.OracleScene(sd0).render(ubo, W, H, O.FXAA), rgmod, "rayConsumption 0")
    rt.close()


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_materials_lights_and_cameras(rgmod, O, S, example_scene, seed):
    """The shading state machine under materials the fixtures never combine: every material of the example scene replaced by a
    seeded random one (transparent + reflective + rough + emissive, rayConsumption 1..5, ior below and above 1, grid effect),
    a random light direction (including straight down and grazing) and a random camera; both schedulers against the oracle."""
    import dataclasses
    rng = np.random.default_rng(seed)
    mats = []
    for k in range(len(example_scene.materials)):
        mats.append(S.make_material(diffuse=rng.uniform(0, 1, 3), transparency=float(rng.choice([0.0, 0.3, 0.98, 1.0])), specular=rng.uniform(0, 1, 3),
                                    reflectivity=float(rng.choice([0.0, 0.2, 0.9, 1.0])), roughness=float(rng.choice([0.0, 0.3, 1.0])),
                                    ior=float(rng.choice([0.8, 1.0, 1.33, 1.5, 2.4])), effectId=int(rng.integers(0, 2)), rayConsumption=int(rng.integers(1, 6)),
                                    emission=float(rng.choice([0.0, 0.5, 3.0]))))
    sd = dataclasses.replace(example_scene, materials=np.stack(mats))
    W, H, ns, mr = 200, 112, 2, 5 + seed % 2
    light = [np.array([0.0, -1.0, 0.0], np.float32), None, np.array([0.995, -0.0998, 0.0], np.float32)][seed % 3]
    cam = S.example_camera_transform()
    cam.position = (cam.position + rng.uniform(-1.5, 1.5, 3)).astype(np.float32)
    cam.look_at(rng.uniform(-1, 1, 3).astype(np.float32))
    ubo = S.make_ubo(cam.to_mat4_colmajor(), S.proj_inverse(W, H), ns, mr, light_dir=light)
    ref = O.OracleScene(sd).render(ubo, W, H, O.FXAA)
    c = ref["counters"]
    assert c["reflect"] > 1000 and c["refract"] > 1000 and c["shadow"] > 1000
    for sched in (rgmod.RG_SCHED_LANES, rgmod.RG_SCHED_POOL):
        rt = rgmod.Raytracer(W, H)
        rt.set_trace_scheduler(sched)
        rt.load_scene(sd)
        rt.render_frame(ubo, rgmod.RG_FXAA | rgmod.RG_DEBUG_IDS)
        _check_frame(rt, ref, rgmod, f"random materials seed {seed} sched={sched}")
        tm = rt.timings()
        for a, b in (("rays_primary", "primary"), ("rays_shadow", "shadow"), ("rays_reflect", "reflect"), ("rays_refract", "refract"), ("sky_lookups", "skylookup")):
            assert abs(tm[a] - c[b]) <= max(4, 2e-4 * c[b]), (a, tm[a], c[b])
        rt.close()


def test_frame_stored_into_pinned_host_memory(rgmod, S, example_scene):
    """rg_set_gather_target on PINNED HOST memory (unified addressing): the final kernel stores the RGBA8 frame there itself, no copy --
    bench.py's e2e path at N = 1.  The host buffer must equal a regular rg_read_rgba8 of the same frame, with and without FXAA, and a
    cleared target must stop the stores (replaces the blit + present of render_system.cpp:130-159 for a headless consumer)."""
    import torch
    rg = rgmod
    W, H = 320, 180
    rt = rg.Raytracer(W, H)
    rt.load_scene(example_scene)
    inst = rt.pack_instances(example_scene.inst_xform, example_scene.inst_meta)
    host = torch.zeros((H, W, 4), dtype=torch.uint8, pin_memory=True)
    rt.set_gather_target(host.data_ptr())
    for ns, flags in ((1, rg.RG_FXAA), (4, 0)):
        rt.render_frame(S.example_ubo(W, H, num_samples=ns), flags, inst)
        rt.sync()
        assert np.array_equal(host.numpy(), rt.read_rgba8()), (ns, flags)
    rt.set_gather_target(0)
    host.zero_()
    rt.render_frame(S.example_ubo(W, H, num_samples=1), rg.RG_FXAA, inst)
    rt.sync()
    assert not host.numpy().any()
    rt.close()
