"""Developer experiment: trace-kernel time of each rank's share in partitioned mode, all contexts on one GPU (run sequentially)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import raygun_b200 as rg
import bench
from raygun_b200.parallel import band_region, attach_partition_in_process
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
desc, W, H, sd, ubo = bench.make_workload(wl)
for world in (1, 2, 4, 8):
    rts = []
    for r in range(world):
        rt = rg.Raytracer(W, H)
        if world > 1: rt.set_region(*band_region(W, H, r, world))
        rt.load_scene(sd); rt.updateRenderTarget(ubo)
        rts.append(rt)
    if world > 1: attach_partition_in_process(rts)
    for it in range(3):
        for rt in rts: rt.doRaytracing(rg.RG_FXAA)
        for rt in rts: rt.sync()
    tms = [rt.timings() for rt in rts]
    print(f"{wl} world={world}: trace_kernel_ms per rank {[round(t['trace_kernel_ms'], 3) for t in tms]}  rays {[t['rays'] for t in tms]}")
    for rt in rts: rt.close()
