#!/usr/bin/env python3
"""Builder timing on one GPU: BLAS build of the flattened 1 003 522-triangle sphere grid (host wall time around rg_build_blas,
buffers already uploaded, second build = warm allocations) and the per-frame TLAS rebuild of 10 001 instances (CUDA events)."""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import raygun_b200 as rg
from raygun_b200 import scene as S

W, H = 640, 360
res = {}
sd, vi = S.sphere_grid_scene(28, flattened=True)
rt = rg.Raytracer(W, H)
rt.setupModelBuffers(sd); rt.sync()
ts = []
for k in range(4):
    t0 = time.perf_counter(); rt.setupBottomLevelAS(); rt.sync(); ts.append((time.perf_counter() - t0) * 1e3)
res["blas_build_1M_tris_ms_wall"] = ts
res["blas"] = rt.debug_bvh_stats()
v = sd.vertices.copy()
ref = []
for n in range(6):
    rt.refitBottomLevelAS(0, v); rt.sync(); ref.append(rt.timings()["as_build_ms"])
res["blas_refit_1M_tris_ms"] = float(np.median(ref[2:]))
balls = S.AnimatedBalls(100)
rt2 = rg.Raytracer(W, H)
rt2.load_scene(balls.scene(0.0))
as_ms = []
for n in range(30):
    inst = rt2.pack_instances(balls.instances(n / 60.0), balls.meta)
    rt2.setupTopLevelAS(inst); rt2.sync()
    as_ms.append(rt2.timings()["as_build_ms"])
res["tlas_10k_rebuild_ms"] = float(np.median(as_ms[5:]))
res["tlas"] = rt2.debug_bvh_stats()
print(json.dumps(res))
