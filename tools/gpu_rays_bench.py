"""Developer experiment: throughput of the bare traversal kernel (one thread per ray) on coherent and incoherent rays."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import raygun_b200 as rg
from raygun_b200 import scene as S
import bench
from oracle import oracle as O
for wl in ("c2", "c3"):
    desc, W, H, sd, ubo = bench.make_workload(wl)
    W, H = 1920, 1080
    rt = rg.Raytracer(W, H); rt.load_scene(sd)
    f = ubo.view(np.float32); VI = f[0:16].reshape(4, 4).T; PI = f[16:32].reshape(4, 4).T
    ys, xs = np.mgrid[0:H, 0:W]
    # tile order 8x4 like the kernel
    d = np.stack([(xs + 0.5) / W * 2 - 1, (ys + 0.5) / H * 2 - 1, np.ones_like(xs, float), np.ones_like(xs, float)], -1).reshape(-1, 4).astype(np.float32)
    tgt = d @ PI.T; t3 = tgt[:, :3] / np.linalg.norm(tgt[:, :3], axis=1, keepdims=True)
    dirs = t3 @ VI[:3, :3].T
    org = np.broadcast_to(VI[:3, 3], dirs.shape)
    rays = np.concatenate([org, dirs, np.full((len(dirs), 1), 0.001), np.full((len(dirs), 1), 10000.0)], 1).astype(np.float32)
    tuv, ip = rt.debug_trace_rays(rays)
    ms = rt.lib.rg_debug_last_trace_rays_ms(rt.h)
    print(f"{wl}: primary rays {len(rays)}: {ms:.3f} ms  {len(rays)/ms/1e3:.0f} Mrays/s  hit frac {(ip[:,0]!=0xffffffff).mean():.3f}")
    hitm = ip[:, 0] != 0xffffffff
    P = (org + dirs * tuv[:, :1])[hitm]
    rng = np.random.default_rng(0)
    for name, perm in (("secondary, pixel order", False), ("secondary, shuffled", True)):
        dd = rng.normal(size=P.shape).astype(np.float32); dd[:, 1] = np.abs(dd[:, 1])
        r2 = np.concatenate([P, dd, np.full((len(P), 1), 0.01), np.full((len(P), 1), 1000.0)], 1).astype(np.float32)
        if perm: r2 = r2[rng.permutation(len(r2))]
        r2 = np.concatenate([r2] * 4)
        rt.debug_trace_rays(r2)
        ms = rt.lib.rg_debug_last_trace_rays_ms(rt.h)
        print(f"{wl}: {name} {len(r2)}: {ms:.3f} ms  {len(r2)/ms/1e3:.0f} Mrays/s")
