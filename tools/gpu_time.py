#!/usr/bin/env python3
"""Developer timing: python tools/gpu_time.py [workload] -> trace ms / Mrays/s for the library in RGB200_LIB (or default)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import raygun_b200 as rg
import bench
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
desc, W, H, sd, ubo = bench.make_workload(wl)
rt = rg.Raytracer(W, H)
rt.load_scene(sd)
rt.updateRenderTarget(ubo)
inst = rt.pack_instances(sd.inst_xform, sd.inst_meta)
ts = []
for i in range(8):
    rt.flush_l2()
    rt.setupTopLevelAS(inst)
    rt.doRaytracing(rg.RG_FXAA)
    tm = rt.timings()
    if i >= 3: ts.append((tm["rt_only_ms"], tm["postproc_ms"], tm["as_build_ms"]))
t = np.median(np.array(ts), axis=0)
print(f"{os.environ.get('RGB200_LIB','default'):40s} sched={os.environ.get('RGB200_TRACE_SCHED','auto')}->{tm['trace_scheduler']} {wl}: trace {t[0]:.3f} ms  post {t[1]:.3f} ms  as {t[2]:.3f} ms  rays {tm['rays']}  {tm['rays']/t[0]/1e3:.0f} Mrays/s")

rt.doRaytracing(rg.RG_FXAA | rg.RG_COUNT_TRAVERSAL)
tc = rt.timings()
r = max(tc["rays"], 1)
print(f"   per ray: nodes {tc['nodes_visited']/r:.2f} tris {tc['tris_tested']/r:.2f} inst {tc['instances_entered']/r:.2f} generic_hits {tc['generic_hits']/r:.2f}"
      f" | primary {tc['rays_primary']} shadow {tc['rays_shadow']} reflect {tc['rays_reflect']} refract {tc['rays_refract']} sky {tc['sky_lookups']}  bvh {rt.debug_bvh_stats()}")
