#!/usr/bin/env python3
"""BASELINE config 4 on one GPU: 10 000 bouncing ball instances, TLAS rebuilt every frame (+ optional refit of the flattened
1 003 522-triangle mesh).  Prints as_build_ms separately from the trace, as BASELINE.md asks."""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import raygun_b200 as rg
from raygun_b200 import scene as S

W, H = 1920, 1080
balls = S.AnimatedBalls(100)
rt = rg.Raytracer(W, H)
rt.load_scene(balls.scene(0.0))
ubo = S.make_ubo(balls.view_inverse, S.proj_inverse(W, H), 1, 5)
rt.updateRenderTarget(ubo)
res = {"instances": len(balls.meta), "frames": 120}
as_ms, tr_ms, tot = [], [], []
for n in range(120):
    inst = rt.pack_instances(balls.instances(n / 60.0), balls.meta)
    rt.setupTopLevelAS(inst)
    rt.doRaytracing(rg.RG_FXAA)
    tm = rt.timings()
    if n >= 5:
        as_ms.append(tm["as_build_ms"]); tr_ms.append(tm["rt_only_ms"]); tot.append(tm["rt_total_ms"])
# the same animation through the device-side scene-graph walk (rg_set_entities): host time per frame of both routes
t_inst, t_ent = [], []
for n in range(40):
    t0 = time.perf_counter(); inst = rt.pack_instances(balls.instances(n / 60.0), balls.meta); rt.setupTopLevelAS(inst); rt.sync(); t_inst.append(time.perf_counter() - t0)
    t0 = time.perf_counter(); ents = balls.entities(n / 60.0); ni = rt.set_entities(ents); rt.sync(); t_ent.append(time.perf_counter() - t0)
res.update(host_ms_instances_route=float(np.median(t_inst)) * 1e3, host_ms_entities_route=float(np.median(t_ent)) * 1e3, entities_route_instances=int(ni))
res.update(tlas_rebuild_ms=float(np.median(as_ms)), trace_ms=float(np.median(tr_ms)), rt_total_ms=float(np.median(tot)), rays=tm["rays"],
           mrays_s=tm["rays"] / float(np.median(tr_ms)) / 1e3, bvh=rt.debug_bvh_stats())
# refit: vertex-wobbled copy of the flattened sphere grid (1 003 522 triangles)
sd, vi = S.sphere_grid_scene(28, flattened=True)
rt2 = rg.Raytracer(W, H)
t0 = time.time(); rt2.load_scene(sd); rt2.sync(); res["blas_build_1M_tris_s_incl_upload"] = time.time() - t0
v = sd.vertices.copy()
ref_ms = []
for n in range(8):
    p = v.view(np.float32)
    p[:, 1] = sd.vertices.view(np.float32)[:, 1] + 0.05 * np.sin(3.0 * sd.vertices.view(np.float32)[:, 0] + n)
    rt2.refitBottomLevelAS(0, v)
    rt2.sync()
    ref_ms.append(rt2.timings()["as_build_ms"])
res["blas_refit_1M_tris_ms"] = float(np.median(ref_ms[2:]))
print(json.dumps(res))
