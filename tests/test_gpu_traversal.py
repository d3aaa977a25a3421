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
