"""-m gpu: the CUDA traversal (through the C ABI) against the oracle's closest-hit definition, ray by ray."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _random_rays(sd, n, seed):
    """Secondary-like rays: origins on random triangles of random instances (world space), random directions."""
    rng = np.random.default_rng(seed)
    pos = sd.positions()
    rays = np.zeros((n, 8), np.float32)
    for k in range(n):
        i = rng.integers(len(sd.inst_meta))
        mesh, voff, ioff, _ = (int(v) for v in sd.inst_meta[i])
        ntri = int(sd.meshes[mesh, 3]) // 3
        p = rng.integers(ntri)
        tri = pos[voff + sd.indices[ioff + 3 * p: ioff + 3 * p + 3]]
        b = rng.dirichlet((1, 1, 1)).astype(np.float32)
        o = (tri * b[:, None]).sum(0)
        m = sd.inst_xform[i].reshape(3, 4)
        o = m[:, :3] @ o + m[:, 3]
        d = rng.normal(size=3).astype(np.float32)
        if k % 7 == 0:
            d[rng.integers(3)] = 0.0          # axis-parallel component
        if k % 11 == 0:
            d *= np.float32(10.0 ** rng.uniform(-3, 3))   # un-normalised directions keep t semantics
        rays[k, :3] = o; rays[k, 3:6] = d; rays[k, 6] = 0.01; rays[k, 7] = 1000.0
    return rays


def test_closest_hit_matches_oracle(example_scene, oracle_example):
    import raygun_b200 as rg
    sd = example_scene
    rt = rg.Raytracer(64, 36)
    rt.load_scene(sd)
    rays = _random_rays(sd, 20000, 1)
    tuv, ip = rt.debug_trace_rays(rays)
    bad = 0
    for k in range(len(rays)):
        hit, t, u, v, inst, prim = oracle_example.closest_hit(rays[k, :3], rays[k, 3:6], 0.01, 1000.0)
        if hit:
            ok = ip[k, 0] == inst and ip[k, 1] == prim and tuv[k, 0] == np.float32(t) and tuv[k, 1] == np.float32(u) and tuv[k, 2] == np.float32(v)
        else:
            ok = ip[k, 0] == 0xffffffff
        bad += not ok
    assert bad == 0, f"{bad} of {len(rays)} rays differ from the oracle (bit-exact t,u,v + ids expected)"


def test_negative_zero_direction_components(example_scene, oracle_example):
    """Axis-parallel rays whose zero components carry a MINUS sign (reflections produce them): the slab test's near / far plane choice
    must follow the sign bit of the clamped reciprocal, not `d >= 0` (which is true for -0 while 1 / -1e-30 is negative)."""
    import raygun_b200 as rg
    sd = example_scene
    rt = rg.Raytracer(64, 36)
    rt.load_scene(sd)
    rays = _random_rays(sd, 4000, 7)
    rng = np.random.default_rng(8)
    for k in range(len(rays)):
        axes = rng.permutation(3)[: 1 + (k % 2)]
        rays[k, 3 + axes] = np.float32(-0.0)
    assert np.signbit(rays[:, 3:6]).any()
    tuv, ip = rt.debug_trace_rays(rays)
    bad = 0
    for k in range(len(rays)):
        hit, t, u, v, inst, prim = oracle_example.closest_hit(rays[k, :3], rays[k, 3:6], 0.01, 1000.0, brute=(k % 4 == 0))
        if hit:
            ok = ip[k, 0] == inst and ip[k, 1] == prim and tuv[k, 0] == np.float32(t) and tuv[k, 1] == np.float32(u) and tuv[k, 2] == np.float32(v)
        else:
            ok = ip[k, 0] == 0xffffffff
        bad += not ok
    assert bad == 0, f"{bad} of {len(rays)} rays with -0 direction components differ from the oracle"


def test_skewed_and_grazing_rays(example_scene, oracle_example):
    """What the packed binary16 node test (rg_trace.cu pairTest) has to get right: direction components that differ by many orders of
    magnitude (its node-local scaling follows the steepest axis, the fastest axis falls towards the subnormal range), rays that graze
    flat geometry, origins thousands of node sizes away.  Closest hits stay bit-exact; tests/test_slab_half_model.py has the arithmetic."""
    import raygun_b200 as rg
    sd = example_scene
    rt = rg.Raytracer(64, 36)
    rt.load_scene(sd)
    rays = _random_rays(sd, 9000, 21)
    rng = np.random.default_rng(22)
    n = len(rays)
    rays[:, 3:6] *= (10.0 ** -rng.uniform(0, 9, (n, 3)) * (rng.random((n, 3)) < 0.6) + (rng.random((n, 3)) >= 0.6)).astype(np.float32)
    rays[:, 3:6] = np.where(np.abs(rays[:, 3:6]).max(axis=1, keepdims=True) > 0, rays[:, 3:6], np.float32(1.0))
    graze = np.arange(0, n, 5)                      # nearly inside the plane of the triangle the origin sits on: y barely moves
    rays[graze, 4] = rays[graze, 4] * np.float32(1e-6)
    far = np.arange(3, n, 9)                        # the same line, started far behind
    nrm = np.linalg.norm(rays[far, 3:6].astype(np.float64), axis=1, keepdims=True)
    rays[far, :3] = (rays[far, :3] - rays[far, 3:6] / nrm * (10.0 ** rng.uniform(1, 4, (len(far), 1)))).astype(np.float32)
    rays[far, 7] = np.float32(3.0e38)
    tuv, ip = rt.debug_trace_rays(rays)
    bad = []
    for k in range(n):
        hit, t, u, v, inst, prim = oracle_example.closest_hit(rays[k, :3], rays[k, 3:6], 0.01, float(rays[k, 7]), brute=(k % 8 == 0))
        if hit:
            ok = ip[k, 0] == inst and ip[k, 1] == prim and tuv[k, 0] == np.float32(t) and tuv[k, 1] == np.float32(u) and tuv[k, 2] == np.float32(v)
        else:
            ok = ip[k, 0] == 0xffffffff
        if not ok:
            bad.append((k, hit, t, inst, prim, int(ip[k, 0]), int(ip[k, 1]), float(tuv[k, 0])))
    assert not bad, f"{len(bad)} of {n} skewed / grazing rays differ from the oracle, first: {bad[:5]}"


def _torture_scene(S, seed=3):
    """Three meshes (a triangle soup with degenerate, duplicated, tiny and huge triangles; a box; one triangle) under 48 instances with
    rotations, non-uniform and NEGATIVE scales, shear, pure translations, exact duplicates (ties on the instance id), tiny and large scales."""
    rng = np.random.default_rng(seed)
    F = np.float32
    def verts(p):
        v = np.zeros((len(p), 8), np.uint32)
        f = v.view(F)
        f[:, 0:3] = p; f[:, 4:7] = (0, 1, 0)
        return v
    # soup: 300 triangles
    tri = rng.uniform(-1, 1, size=(300, 3, 3)).astype(F)
    tri[:, 1:] = tri[:, :1] + rng.normal(scale=0.25, size=(300, 2, 3)).astype(F)
    tri[10, 2] = tri[10, 1]                                  # zero area: two equal vertices
    tri[11, 2] = tri[11, 0] + F(2) * (tri[11, 1] - tri[11, 0])   # collinear
    tri[12] = tri[12, 0]                                     # a point
    tri[20:30] = tri[40:50]                                  # exact duplicates (smaller primitive id must win)
    tri[30, 1:] = tri[30, :1] + rng.normal(scale=1e-4, size=(2, 3)).astype(F)   # tiny
    tri[31] *= F(300.0)                                      # huge
    soup_v = verts(tri.reshape(-1, 3)); soup_i = np.arange(900, dtype=np.uint32)
    c = np.array([[x, y, z] for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)], F) * F(0.5)
    box_i = np.array([0, 1, 3, 0, 3, 2, 4, 6, 7, 4, 7, 5, 0, 4, 5, 0, 5, 1, 2, 3, 7, 2, 7, 6, 0, 2, 6, 0, 6, 4, 1, 5, 7, 1, 7, 3], np.uint32)
    one_v = verts(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], F)); one_i = np.array([0, 1, 2], np.uint32)
    vertices = np.concatenate([soup_v, verts(c), one_v]); indices = np.concatenate([soup_i, box_i, one_i])
    meshes = np.array([(0, 900, 0, 900), (900, 8, 900, 36), (908, 3, 936, 3)], np.uint32)
    xf, meta = [], []
    def add(m, mesh):
        xf.append(np.asarray(m, F).reshape(12)); meta.append((mesh, int(meshes[mesh, 0]), int(meshes[mesh, 2]), 0))
    for k in range(44):
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        sc = np.exp(rng.uniform(-1.5, 1.5, 3)) * rng.choice([-1.0, 1.0], 3)
        a = q @ np.diag(sc)
        if k % 5 == 0: a = a + 0.3 * rng.normal(size=(3, 3))      # shear
        if k % 7 == 0: a = np.eye(3)                               # pure translation
        if k == 3: a = a * 1e-3
        if k == 4: a = a * 50.0
        t = rng.uniform(-6, 6, 3)
        add(np.concatenate([a, t[:, None]], 1), k % 3)
    for k in (1, 2, 8, 9):                                         # exact duplicates of earlier instances
        xf.append(xf[k].copy()); meta.append(meta[k])
    mats = np.stack([S.make_material()])
    return S.SceneData(vertices, indices, meshes, mats, np.stack(xf).astype(F), np.array(meta, np.uint32), name="torture")


def _torture_rays(sd):
    rays = _random_rays(sd, 6000, 5)
    rng = np.random.default_rng(6)
    n = len(rays)
    free = rng.integers(0, n, n // 3)                    # origins anywhere, not on a surface
    rays[free, :3] = rng.uniform(-8, 8, size=(len(free), 3)).astype(np.float32)
    tiny = rng.integers(0, n, n // 10)
    rays[tiny, 3 + rng.integers(0, 3, len(tiny))] = np.float32(1e-36) * rng.choice([-1, 1], len(tiny)).astype(np.float32)   # below the slab clamp
    neg0 = rng.integers(0, n, n // 10)
    rays[neg0, 3 + rng.integers(0, 3, len(neg0))] = np.float32(-0.0)
    far = rng.integers(0, n, n // 20)
    rays[far, :3] *= np.float32(1e4)                     # origins far outside
    rays[:, 6] = np.where(rng.random(n) < 0.2, 0.0, 0.01).astype(np.float32)
    rays[:, 7] = np.where(rng.random(n) < 0.2, rng.uniform(0.05, 5.0, n), 1000.0).astype(np.float32)
    return rays


def test_torture_scene_closest_hit():
    """Degenerate / duplicated geometry, mirrored and sheared instances, extreme rays: ids and t, u, v bit for bit against the oracle
    (every fourth ray against its brute-force loop over all triangles)."""
    import raygun_b200 as rg
    from raygun_b200 import scene as S
    from oracle import oracle as O
    sd = _torture_scene(S)
    osc = O.OracleScene(sd)
    rt = rg.Raytracer(64, 36)
    rt.load_scene(sd)
    rays = _torture_rays(sd)
    n = len(rays)
    tuv, ip = rt.debug_trace_rays(rays)
    bad = []
    for k in range(n):
        hit, t, u, v, inst, prim = osc.closest_hit(rays[k, :3], rays[k, 3:6], float(rays[k, 6]), float(rays[k, 7]), brute=(k % 4 == 0))
        if hit:
            ok = ip[k, 0] == inst and ip[k, 1] == prim and tuv[k, 0] == np.float32(t) and tuv[k, 1] == np.float32(u) and tuv[k, 2] == np.float32(v)
        else:
            ok = ip[k, 0] == 0xffffffff
        if not ok:
            bad.append((k, hit, t, inst, prim, int(ip[k, 0]), int(ip[k, 1]), float(tuv[k, 0])))
    assert not bad, f"{len(bad)} of {n} rays differ from the oracle, first: {bad[:5]}"
