"""The binary16 slab test of the traversal (raygun_b200/csrc/rg_trace.cu travNode / pairTest, RG_HALF_SLAB) restated in numpy, operation
for operation, and checked against binary64: it may report a child box as hit that is not (a wasted visit), never the reverse -- a
missed box could lose the closest hit, and closest hits are bit-exact against the oracle by contract.

Replaces nothing of the reference (its traversal lives in the Vulkan driver: raygun/render/raytracer.cpp:99 traceRaysKHR); this is the
error analysis of the node test as an executable statement.  CPU only."""
import numpy as np

F = np.float32
K_REL_C, K_REL_A, K_ABS, K_SLACK = F(5.0e-4), F(7.7e-9), F(6.0e-8), F(7.3e-7)


def fma32(a, b, c):
    with np.errstate(over="ignore", invalid="ignore"):
        return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F)


def to_half(x):
    with np.errstate(over="ignore"):
        return x.astype(np.float16)


def hfma_exact(q, a_h, c_h):
    """q 2^-24 (a subnormal half, exact) * a + c on the binary16 operands, before the rounding of the result."""
    with np.errstate(over="ignore", invalid="ignore"):
        return q.astype(np.float64) * 2.0 ** -24 * a_h.astype(np.float64) + c_h.astype(np.float64)


def hfma(q, a_h, c_h):
    """HFMA2: ONE rounding to binary16."""
    with np.errstate(over="ignore", invalid="ignore"):
        return hfma_exact(q, a_h, c_h).astype(np.float16)


def slab_constants(p, e, o, inv_d, tmin, tmax):
    """travNode's prologue.  p, o: (n, 3) float32; e: (n, 3) int exponents; inv_d: (n, 3) float32 approximate reciprocals."""
    sx = np.ldexp(F(1.0), e).astype(F)
    a = sx * inv_d
    b = (p - o) * inv_d
    aa = np.abs(a)
    m = aa.max(axis=1)
    me = m.view(np.uint32) & np.uint32(0x7F800000)
    me = np.minimum(np.maximum(me, np.uint32(0x0A000000)), np.uint32(0x79000000))
    S = (np.uint32(0x86000000) - me).view(F)
    s = (np.uint32(0x7A000000) - me).view(F)
    t0 = np.where(aa[:, 0] <= aa[:, 1], b[:, 0], b[:, 1])
    t0 = np.where(aa[:, 2] < np.minimum(aa[:, 0], aa[:, 1]), b[:, 2], t0)
    t0s = t0 * s
    a_s = a * S[:, None]
    cs = fma32(b, s[:, None] * np.ones_like(b), -t0s[:, None] * np.ones_like(b))
    e0 = fma32(np.abs(t0s), np.full_like(t0s, K_SLACK), np.full_like(t0s, K_ABS))
    err = fma32(np.abs(cs), np.full_like(cs, K_REL_C), fma32(np.abs(a_s), np.full_like(cs, K_REL_A), e0[:, None] * np.ones_like(cs)))
    tn0 = fma32(tmin, s, -t0s)
    tf0 = fma32(tmax, s, -t0s)
    with np.errstate(invalid="ignore", over="ignore"):
        tn1 = fma32(np.abs(tn0), np.full_like(tn0, -K_REL_C), tn0) - K_ABS
        tf1 = fma32(np.abs(tf0), np.full_like(tf0, K_REL_C), tf0) + K_ABS
        c_near, c_far = to_half(cs - err), to_half(cs + err)
    return dict(a_h=to_half(a_s), c_near=c_near, c_far=c_far, tmn=to_half(tn1), tmx=to_half(tf1), t0=t0, s=s)


def make_cases(n, rng, mode):
    p = rng.uniform(-100, 100, (n, 3)).astype(F)
    e = rng.integers(-12, 3, (n, 3))
    if mode == "typical":   # a node about as large on every axis, seen from up to 100 node sizes away, no closer hit known yet
        e[:] = e[:, :1] + rng.integers(-2, 3, (n, 3))
    if mode == "flat":
        e[np.arange(n), rng.integers(0, 3, n)] = -120
    ext = 255.0 * np.ldexp(1.0, e)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    if mode == "skewed":
        d *= 10.0 ** -rng.uniform(0, 9, (n, 3))
    if mode == "scaled":   # object-space directions under a scaled instance are not unit vectors
        d *= 10.0 ** rng.uniform(-3, 3, (n, 1))
    d = d.astype(F)
    if mode == "parallel":
        z = rng.random((n, 3)) < 0.4
        z[z.all(axis=1), 0] = False
        d = np.where(z, F(0.0) * np.sign(d), d).astype(F)
    # a point inside or near the node, and an origin some way back along the ray from it
    target = p + rng.uniform(-0.3, 1.3, (n, 3)) * ext
    back = (10.0 ** rng.uniform(-3, 2.0 if mode == "typical" else 4.5, (n, 1))) * ext.max(axis=1, keepdims=True)
    if mode == "inside":
        back *= 0.0
    dn = d.astype(np.float64)
    nrm = np.linalg.norm(dn, axis=1, keepdims=True)
    o = (target - dn / np.where(nrm > 0, nrm, 1.0) * back).astype(F)
    # setupSlab: clamp, then a reciprocal good to 1 ulp
    eps = F(1e-30)
    dc = np.where(np.abs(d) > eps, d, np.copysign(eps, d)).astype(F)
    inv = (1.0 / dc.astype(np.float64)).astype(F)
    inv = np.nextafter(inv, np.where(rng.random((n, 3)) < 0.5, F(np.inf), F(-np.inf)).astype(F)).astype(F)
    if mode == "typical":   # child boxes of 64..128 quantisation steps per axis, as the eight children of a node have
        qlo = rng.integers(0, 192, (n, 3))
        qhi = np.minimum(255, qlo + rng.integers(64, 128, (n, 3)))
    else:                   # anything, flat boxes (qlo == qhi) included
        qlo = rng.integers(0, 256, (n, 3))
        qhi = np.minimum(255, qlo + rng.integers(0, 256, (n, 3)) * (rng.random((n, 3)) < 0.9))
    tmin = np.full(n, 1e-3, F)
    tmax = np.where(rng.random(n) < (1.0 if mode == "typical" else 0.5), F(3.0e38), (10.0 ** rng.uniform(-2, 5, n))).astype(F)
    return p, e, o, dc, inv, qlo, qhi, tmin, tmax


def check(mode, n=400_000, seed=1):
    rng = np.random.default_rng(seed)
    p, e, o, dc, inv, qlo, qhi, tmin, tmax = make_cases(n, rng, mode)
    K = slab_constants(p, e, o, inv, tmin, tmax)
    neg = np.signbit(inv)
    qn, qf = np.where(neg, qhi, qlo), np.where(neg, qlo, qhi)
    tn_h = hfma(qn, K["a_h"], K["c_near"]).astype(np.float64)
    tf_h = hfma(qf, K["a_h"], K["c_far"]).astype(np.float64)
    # binary64 truth on the same scale: ((p + q 2^e - o) / d - t0) s with the float inputs taken as exact
    cell = np.ldexp(1.0, e)
    t0, s = K["t0"].astype(np.float64)[:, None], K["s"].astype(np.float64)[:, None]
    with np.errstate(over="ignore", invalid="ignore"):
        Tn = ((p.astype(np.float64) + qn * cell - o.astype(np.float64)) / dc.astype(np.float64) - t0) * s
        Tf = ((p.astype(np.float64) + qf * cell - o.astype(np.float64)) / dc.astype(np.float64) - t0) * s
        Tmn = (tmin.astype(np.float64) - t0[:, 0]) * s[:, 0]
        Tmx = (tmax.astype(np.float64) - t0[:, 0]) * s[:, 0]
    assert not np.isnan(tn_h).any() and not np.isnan(tf_h).any(), mode
    # before the (monotone) rounding of the result, every near plane lies at or below its true place and every far plane at or above:
    # a true near <= far then survives the rounding, which is applied to both sides alike
    xn, xf = hfma_exact(qn, K["a_h"], K["c_near"]), hfma_exact(qf, K["a_h"], K["c_far"])
    assert (xn <= Tn).all(), (mode, "near plane moved inwards", int((xn > Tn).sum()))
    assert (xf >= Tf).all(), (mode, "far plane moved inwards", int((xf < Tf).sum()))
    tmn_h, tmx_h = K["tmn"].astype(np.float64), K["tmx"].astype(np.float64)
    assert (tmn_h <= Tmn).all() and (tmx_h >= Tmx).all(), mode
    # the decision itself, and how many boxes the rounding adds
    hit_h = np.maximum(tn_h.max(axis=1), tmn_h) <= np.minimum(tf_h.min(axis=1), tmx_h)
    hit_t = np.maximum(Tn.max(axis=1), Tmn) <= np.minimum(Tf.min(axis=1), Tmx)
    assert not (hit_t & ~hit_h).any(), (mode, "missed box")
    return int(hit_t.sum()), int((hit_h & ~hit_t).sum())


def test_half_slab_is_conservative_everywhere():
    extra = {}
    for i, mode in enumerate(["plain", "typical", "inside", "skewed", "parallel", "flat", "scaled"]):
        hits, false_pos = check(mode, seed=10 + i)
        extra[mode] = (hits, false_pos)
        assert hits > 1000, (mode, hits)   # the cases do exercise boxes that are hit
    # ordinary rays and child boxes: the rounding slack (about half a quantisation step per plane) adds ~2 % of visits, no more
    hits, false_pos = extra["typical"]
    assert false_pos < 0.03 * hits, extra


def test_byte_to_subnormal_half_is_exact():
    q = np.arange(256, dtype=np.uint16)
    assert (q.view(np.float16).astype(np.float64) == q * 2.0 ** -24).all()   # what PRMT(word, 0, 0x4140) builds per half


def test_model_constants_are_the_kernels():
    """The numpy restatement above checks what rg_trace.cu compiles only as long as the two carry the same constants."""
    import os
    import re
    src = open(os.path.join(os.path.dirname(__file__), "..", "raygun_b200", "csrc", "rg_trace.cu")).read()
    m = re.search(r"kRelC = ([0-9.e+-]+)f, kRelA = ([0-9.e+-]+)f, kAbs = ([0-9.e+-]+)f, kSlack = ([0-9.e+-]+)f;", src)
    assert m, "margin constants not found in travNode"
    assert tuple(F(float(v)) for v in m.groups()) == (K_REL_C, K_REL_A, K_ABS, K_SLACK)
    for magic in ("0x86000000u - me", "0x7A000000u - me", "0x0A000000u", "0x79000000u", "__byte_perm(nx, 0u, SEL)", "pairTest<0x4140, 0>", "pairTest<0x4342, 1>"):
        assert magic in src, magic
    types = open(os.path.join(os.path.dirname(__file__), "..", "raygun_b200", "csrc", "rg_types.cuh")).read()
    assert "#define RG_HALF_SLAB 1" in types   # the formulation under test is the one that ships
