"""The C++ host shim (raygun_b200/host: Transform / Entity / Camera / Material / ResourceManager / RenderSystem mirror).

CPU: math against the reference's vendored GLM (golden JSON), the Collada + rgmat loaders against the committed snapshot
(only where the reference's resources exist), the entity DFS (pruning rules of acceleration_structure.cpp:63-85).
GPU: a frame rendered through Scene / Entity / RenderSystem equals the frame rendered through the Python mirror."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from raygun_b200 import scene as S

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
LIB = os.path.join(ROOT, "raygun_b200", "libraygun_host.so")
REF_RES = "/root/reference/resources"


@pytest.fixture(scope="module")
def host():
    if not os.path.exists(LIB):
        pytest.fail(f"{LIB} missing: run `make -C raygun_b200/host` (or __graft_entry__.build())")
    lib = C.CDLL(LIB)
    lib.rgh_last_error.restype = C.c_char_p
    lib.rgh_example_scene_load.restype = C.c_void_p
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_math_matches_reference_glm(host):
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "glm_golden.json")))
    out = np.zeros(141, np.float32)
    host.rgh_math_golden(_p(out))
    f = lambda k: np.array(g[k], np.uint32).view(np.float32)  # noqa: E731
    want = np.concatenate([f("instance_Raygun"), f("instance_ph3_games"), f("instance_room"), f("instance_Ball"), f("viewInverse"), f("cam_quat_wxyz"),
                           f("projInverse_640x360"), f("projInverse_100x60"), f("lightDir"), f("trs_compose_3x4"), f("decompose_pos"), f("decompose_scale"),
                           f("decompose_quat_wxyz"), f("viewInverse_c3")])
    assert out.shape == want.shape
    # instance matrices of the example scene: exact
    assert np.array_equal(out[:48].view(np.uint32), want[:48].view(np.uint32))
    assert np.allclose(out, want, rtol=3e-6, atol=3e-6), np.abs(out - want).max()


@pytest.mark.skipif(not os.path.isdir(REF_RES), reason="reference resources not present (GPU box)")
def test_collada_and_material_loaders_reproduce_the_snapshot(host, example_scene):
    h = C.c_void_p(host.rgh_example_scene_load(REF_RES.encode(), C.c_uint32(640), C.c_uint32(360)))
    assert h.value, host.rgh_last_error().decode()
    cnt = np.zeros(5, np.uint32)
    host.rgh_scene_counts(h, _p(cnt))
    assert cnt.tolist() == [56742, 56742, 16, 4, 4]
    v = np.zeros((cnt[0], 8), np.uint32); i = np.zeros(cnt[1], np.uint32); m = np.zeros((cnt[2], 16), np.uint32)
    r = np.zeros((cnt[3], 4), np.uint32); inst = np.zeros((cnt[4], 16), np.uint32); ubo = np.zeros(48, np.uint32)
    host.rgh_scene_copy(h, _p(v), _p(i), _p(m), _p(r), _p(inst), _p(ubo))
    host.rgh_scene_free(h)
    sd = example_scene
    assert np.array_equal(v, sd.vertices) and np.array_equal(i, sd.indices) and np.array_equal(r, sd.meshes)
    assert np.array_equal(m, sd.materials)
    assert np.array_equal(inst[:, :12], sd.inst_xform.view(np.uint32)) and np.array_equal(inst[:, 12:], sd.inst_meta)
    want = S.example_ubo(640, 360).view(np.float32)
    assert np.allclose(ubo.view(np.float32)[:32], want[:32], rtol=3e-6, atol=3e-6)


def _entity_arrays(sd):
    """The example scene as an entity tree: root -> level(room.dae) -> 3 children; root -> Ball."""
    models = np.array([(int(sd.meshes[k, 0]), int(sd.meshes[k, 1]), int(sd.meshes[k, 2]), int(sd.meshes[k, 3]), int(sd.inst_meta[k, 3]), 5 if k < 3 else 1)
                       for k in range(4)], np.uint32)
    ents = [(-1, -1, 1)]                                   # 0: level entity, identity, no model
    trs = [(0, 0, 0, 1, 0, 0, 0, 1, 1, 1)]
    s2 = np.float32(np.sqrt(0.5))
    half = 0.5 * np.arctan2(0.7071068, 0.7071068)          # rot-Y 45 deg of the ph3_games node
    for k, t in enumerate([(3, 0, -21, 1, 0, 0, 0, 7.5, 7.5, 7.5), (-9, 0, -21, np.cos(half), 0, np.sin(half), 0, 1, 1, 1), (-24, -4, -24, 1, 0, 0, 0, 1, 1, 1)]):
        ents.append((0, k, 1)); trs.append(t)
    ents.append((-1, 3, 1)); trs.append((3, 0, -3, 1, 0, 0, 0, 1, 1, 1))
    return models, np.array(ents, np.int32), np.array(trs, np.float32)


def test_entity_dfs_order_and_pruning(host, example_scene):
    sd = example_scene
    models, ents, trs = _entity_arrays(sd)
    out = np.zeros((16, 16), np.uint32); n = C.c_uint32()
    assert host.rgh_gather_instances(_p(models), 4, _p(ents), _p(trs), len(ents), _p(out), C.byref(n)) == 0
    assert n.value == 4
    assert np.array_equal(out[:4, 12:], sd.inst_meta)                      # order Raygun, ph3_games, room, Ball + offset table
    assert np.allclose(out[:4, :12].view(np.float32), sd.inst_xform, atol=1e-6)
    # invisible subtree is skipped with its children; zero-volume scale prunes too; model-less entities still descend
    e2 = ents.copy(); e2[0, 2] = 0
    assert host.rgh_gather_instances(_p(models), 4, _p(e2), _p(trs), len(ents), _p(out), C.byref(n)) == 0 and n.value == 1
    assert out[0, 12] == 3
    t2 = trs.copy(); t2[2, 8] = 0.0                                         # ph3_games scaled flat
    assert host.rgh_gather_instances(_p(models), 4, _p(ents), _p(t2), len(ents), _p(out), C.byref(n)) == 0 and n.value == 3
    assert out[:3, 12].tolist() == [0, 2, 3]


@pytest.mark.gpu
def test_render_through_entity_tree_equals_python_path(host, example_scene):
    import raygun_b200 as rg
    sd = example_scene
    W, H = 320, 180
    models, ents, trs = _entity_arrays(sd)
    frame = np.zeros((H, W, 4), np.uint8); inst = np.zeros((16, 16), np.uint32); n = C.c_uint32(); tm = np.zeros(2, np.float32)
    cam_pos = np.array([8, 10, 7], np.float32); cam_tgt = np.array([3, 0, -3], np.float32)
    v = np.ascontiguousarray(sd.vertices); i = np.ascontiguousarray(sd.indices); m = np.ascontiguousarray(sd.materials)
    rc = host.rgh_render_entities(_p(v), len(v), _p(i), len(i), _p(m), len(m), _p(models), 4, _p(ents), _p(trs), len(ents), _p(cam_pos), _p(cam_tgt),
                                  W, H, 2, 5, 1, 0, _p(frame), _p(inst), C.byref(n), _p(tm))
    assert rc == 0, host.rgh_last_error().decode()
    assert n.value == 4 and tm[1] > 0
    rt = rg.Raytracer(W, H)
    rt.load_scene(sd)
    # same instance matrices / UBO as the C++ path produced them
    ubo = S.example_ubo(W, H, num_samples=2)
    rt.render_frame(ubo, rg.RG_FXAA, inst[:4])
    want = rt.read_rgba8()
    d = np.abs(frame[..., :3].astype(int) - want[..., :3].astype(int)).max(axis=2)
    assert float((d <= 1).mean()) > 0.999   # camera matrices agree to ~1e-7: a handful of edge pixels may move
    # the same scene with the scene-graph walk on the GPU (Raytracer::deviceSceneWalk -> rg_set_entities): identical instances and frame
    frame2 = np.zeros_like(frame); inst2 = np.zeros_like(inst); n2 = C.c_uint32()
    rc = host.rgh_render_entities(_p(v), len(v), _p(i), len(i), _p(m), len(m), _p(models), 4, _p(ents), _p(trs), len(ents), _p(cam_pos), _p(cam_tgt),
                                  W, H, 2, 5, 1 | 2, 0, _p(frame2), _p(inst2), C.byref(n2), _p(tm))
    assert rc == 0, host.rgh_last_error().decode()
    assert n2.value == n.value and np.array_equal(inst2[:4], inst[:4]) and np.array_equal(frame2, frame)


def test_text_layout_matches_text_generator_rules(host):
    """ui/text.cpp:108-138 on the committed text snapshot: one instance per printable glyph, x advances by letterPadding + glyph
    width, a space advances 5 paddings, a newline goes down one lineSpacing; MiddleCenter alignment shifts by half the bounds."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "text_scene.npz"))
    text = bytes(z["text"]).decode()
    widths = z["widths"].view(np.float32)
    inst = z["instances"]
    xf = inst[:, :12].view(np.float32).reshape(-1, 3, 4)
    glyphs = [c for c in text if c not in " \n"]
    assert len(inst) == len(glyphs) + 1                      # + the ground quad
    pad, line = np.float32(0.1), np.float32(1.0)
    x = np.float32(0); y = np.float32(0); want = []; ux = uy = np.float32(0)
    for c in text:
        if c == " ":
            x = np.float32(x + np.float32(5) * pad)
        elif c == "\n":
            x = np.float32(0); y = np.float32(y - line); continue
        if widths[ord(c)] == 0:
            continue
        want.append((x, y))
        x = np.float32(x + np.float32(pad + widths[ord(c)]))
        ux = np.float32(x - pad); uy = np.float32(y + np.float32(line * np.float32(0.66)))
    off = np.array([-ux / 2, -uy / 2], np.float32)           # Alignment::MiddleCenter
    got = xf[:len(glyphs), :2, 3] - np.array([0.0, 1.0], np.float32)   # the text entity sits at (0, 1, 0)
    assert np.allclose(got, np.array(want, np.float32) + off, atol=2e-6)
    assert np.allclose(xf[:len(glyphs), :, :3], np.eye(3, dtype=np.float32))   # glyph entities: translation only
    # every glyph mesh starts at x = 0 (loadFont shifts it) and is as wide as charWidth says
    v = z["vertices"].view(np.float32)
    for k, c in enumerate(glyphs):
        m = z["meshes"][inst[k, 12]]
        px = v[m[0]:m[0] + m[1], 0]
        assert px.min() == 0.0 and abs(px.max() - widths[ord(c)]) < 1e-6


def test_font_snapshot_against_independent_known_answers():
    """The glyph meshes and widths of the committed text snapshot against tests/golden/font_known_answers.json, which
    tools/font_known_answers.py computes from the reference's NotoSans.obj with its own numpy reader (NOT with the C++ loader):
    corner count, binary32 width and the CRC32 of the shifted positions of every glyph the text uses."""
    import json, zlib
    z = np.load(os.path.join(ROOT, "tests", "golden", "text_scene.npz"))
    ka = json.load(open(os.path.join(ROOT, "tests", "golden", "font_known_answers.json")))["glyphs"]
    assert len(ka) == 124
    text = bytes(z["text"]).decode()
    widths = z["widths"]                               # uint32 view of the 128 floats
    for code in range(128):
        want = ka.get(str(code))
        assert int(widths[code]) == (want["width_bits"] if want else 0), code
    inst, v = z["instances"], z["vertices"]
    glyphs = [c for c in text if c not in " \n"]
    for k, c in enumerate(glyphs):
        m = z["meshes"][inst[k, 12]]
        pos = np.ascontiguousarray(v[m[0]:m[0] + m[1], 0:3])
        want = ka[str(ord(c))]
        assert m[1] == want["corners"] and m[3] == want["corners"], c        # one vertex per face corner, indices 0..n-1
        assert zlib.crc32(pos.tobytes()) == want["crc32_positions"], c


@pytest.mark.parametrize("align", range(9))
def test_text_layout_rules_with_synthetic_widths(host, align):
    """ui::TextGenerator over a synthetic font (no files): pen advance, space, line feed, glyph-less code points, the extent after the
    last glyph and the nine anchor offsets, against a numpy restatement of the rules of ui/text.cpp:53-138 in binary32."""
    rng = np.random.default_rng(align)
    widths = np.zeros(128, np.float32)
    for c in "abcdefgXYZ019.,":
        widths[ord(c)] = np.float32(rng.uniform(0.2, 0.9))
    text = "ab c\nXYZ  0q9\n\n.,d"                     # 'q' has no glyph
    pad, line = np.float32(0.13), np.float32(1.25)
    out = np.zeros((64, 2), np.float32); b = np.zeros(4, np.float32)
    n = host.rgh_text_layout(text.encode(), align, _p(widths), C.c_float(pad), C.c_float(line), _p(out), 64, _p(b))
    x = y = ux = uy = np.float32(0); pens = []
    for ch in text:
        if ch == "\n":
            x = np.float32(0); y = np.float32(y - line); continue
        if ch == " ":
            x = np.float32(x + np.float32(5) * pad)
        if widths[ord(ch)] == 0:
            continue
        pens.append((x, y))
        x = np.float32(x + np.float32(pad + widths[ord(ch)]))
        ux, uy = np.float32(x - pad), np.float32(y + np.float32(line * np.float32(0.66)))
    col, row = align % 3, align // 3
    ax = np.float32(0) if col == 0 else (np.float32(-ux / 2) if col == 1 else np.float32(-ux))
    ay = np.float32(0) if row == 0 else (np.float32(-uy / 2) if row == 1 else np.float32(-uy))
    assert n == len(pens) == 11
    want = np.array([(np.float32(ax + px), np.float32(ay + py)) for px, py in pens], np.float32)
    assert np.array_equal(out[:n], want)
    assert np.array_equal(b, np.array([ax, ay, np.float32(ux + ax), np.float32(uy + ay)], np.float32))


def test_fade_envelopes(host):
    """render::FadeIn / FadeTransition (raygun/render/fade.cpp:44-95): alpha over time, callback exactly once at the peak, over()."""
    t = np.array([0.0, 0.25, 0.5, 1.0, 1.5, 2.0, 2.5], np.float64)
    col = np.array([0.2, 0.4, 0.6], np.float32)
    rgba = np.zeros((len(t), 4), np.float32); over = np.zeros(len(t), np.int32); peak = C.c_int(-2)
    assert host.rgh_fade_sample(0, C.c_double(1.0), _p(col), _p(t), len(t), _p(rgba), C.byref(peak), _p(over)) == 0
    assert np.allclose(rgba[:, :3], col) and np.allclose(rgba[:, 3], [1, 0.75, 0.5, 0, 0, 0, 0])
    assert list(over) == [0, 0, 0, 0, 1, 1, 1] and peak.value == -1
    assert host.rgh_fade_sample(1, C.c_double(1.0), _p(col), _p(t), len(t), _p(rgba), C.byref(peak), _p(over)) == 0
    assert np.allclose(rgba[:, 3], [0, 0.25, 0.5, 1.0, 0.5, 0, 0])
    assert peak.value == 3 and list(over) == [0, 0, 0, 0, 0, 0, 1]


def test_frame_writers_ppm_and_png(host, tmp_path):
    """The headless stand-in for the swapchain present (render_system.cpp:159): binary PPM and stored-deflate PNG, read back here."""
    import struct, zlib
    W, H = 37, 11
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
    ppm, png = str(tmp_path / "f.ppm"), str(tmp_path / "f.png")
    assert host.rgh_write_image(ppm.encode(), _p(img), W, H, 0) == 0
    assert host.rgh_write_image(png.encode(), _p(img), W, H, 1) == 0
    raw = open(ppm, "rb").read()
    head = f"P6\n{W} {H}\n255\n".encode()
    assert raw.startswith(head) and np.array_equal(np.frombuffer(raw[len(head):], np.uint8).reshape(H, W, 3), img[..., :3])
    d = open(png, "rb").read()
    assert d[:8] == b"\x89PNG\r\n\x1a\n"
    off, chunks = 8, {}
    while off < len(d):
        n, typ = struct.unpack(">I4s", d[off:off + 8])
        body = d[off + 8:off + 8 + n]
        assert struct.unpack(">I", d[off + 8 + n:off + 12 + n])[0] == zlib.crc32(typ + body)
        chunks[typ] = body; off += 12 + n
    assert struct.unpack(">IIBBBBB", chunks[b"IHDR"]) == (W, H, 8, 6, 0, 0, 0)
    rows = np.frombuffer(zlib.decompress(chunks[b"IDAT"]), np.uint8).reshape(H, 1 + 4 * W)
    assert np.all(rows[:, 0] == 0) and np.array_equal(rows[:, 1:].reshape(H, W, 4), img)


@pytest.mark.skipif(not os.path.isdir(REF_RES), reason="reference resources not present (GPU box)")
def test_font_loader_reproduces_the_snapshot(host):
    sys_path = os.path.join(ROOT, "tools")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_text_scene", os.path.join(sys_path, "make_text_scene.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    host.rgh_text_scene_load.restype = C.c_void_p
    d = mod.load(host)
    z = np.load(os.path.join(ROOT, "tests", "golden", "text_scene.npz"))
    for k in ("vertices", "indices", "materials", "meshes", "instances", "ubo", "widths"):
        assert np.array_equal(d[k], z[k]), k
    assert int(d["glyphs"]) == 124                            # objects in NotoSans.obj


@pytest.mark.gpu
def test_text_scene_frame_parity():
    """the ray-traced UI text (one tiny instance per glyph) rendered on the GPU against the oracle."""
    import raygun_b200 as rg
    from oracle import oracle as O
    sd, ubo48 = S.load_text_scene()
    W, H = 640, 360
    ubo = S.make_ubo(ubo48[:16].view(np.float32), ubo48[16:32].view(np.float32), 2, 5)
    ref = O.OracleScene(sd).render(ubo, W, H, O.FXAA)
    for sched in (rg.RG_SCHED_LANES, rg.RG_SCHED_POOL):
        rt = rg.Raytracer(W, H)
        rt.set_trace_scheduler(sched)
        rt.load_scene(sd)
        rt.render_frame(ubo, rg.RG_FXAA | rg.RG_DEBUG_IDS)
        img = rt.read_rgba8(); inst, prim = rt.read_ids()
        assert float(((inst == ref["inst"]) & (prim == ref["prim"])).mean()) >= 0.9999
        d = np.abs(img[..., :3].astype(int) - ref["rgba8"][..., :3].astype(int)).max(axis=2)
        assert float((d <= 2).mean()) >= 0.999
        assert (ref["inst"] < 20).mean() > 0.01               # the glyphs are actually on screen
        rt.close()
