#!/usr/bin/env python3
"""Multi-GPU parity check (run under torchrun, one rank per GPU): the frame gathered on rank 0 from N column bands over
NVLink peer stores must equal rank 0's own full-frame render bit for bit.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/mgpu_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import raygun_b200 as rg  # noqa: E402
from raygun_b200 import scene as S  # noqa: E402
from raygun_b200.parallel import band_region, share_gather_handle  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W, H = 1280, 720
sd, _ = S.load_example_scene()
ubo = S.example_ubo(W, H, num_samples=2)
ok = True
for split in ("columns", "rows"):
    rt = rg.Raytracer(W, H, device=local)
    rt.set_region(*band_region(W, H, rank, world, split))
    rt.load_scene(sd)
    handle, own = (rt.gather_buffer_export() if rank == 0 else (None, None))
    handle = share_gather_handle(dist, rank, handle)
    peer = None
    if rank == 0:
        rt.set_gather_target(own)
    else:
        peer = rt.gather_buffer_open(handle)
        rt.set_gather_target(peer)
    rt.render_frame(ubo, rg.RG_FXAA)
    rt.sync()
    dist.barrier()
    if rank == 0:
        got = rt.read_gathered_rgba8()
        full = rg.Raytracer(W, H, device=local)
        full.load_scene(sd)
        full.render_frame(ubo, rg.RG_FXAA)
        want = full.read_rgba8()
        same = bool(np.array_equal(got, want))
        print(f"[mgpu_check] world={world} split={split}: gathered frame == single-GPU frame: {same} "
              f"(differing bytes: {int((got != want).sum())}); band timings rank0 {rt.timings()['rt_total_ms']:.3f} ms vs full {full.timings()['rt_total_ms']:.3f} ms")
        ok &= same
    dist.barrier()
    if peer is not None:
        rt.gather_buffer_close(peer)
    rt.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
