#!/usr/bin/env python3
"""bench.py -- Mrays/s and ms/frame of the per-frame ray-tracing path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c2|c1|c3|c5|c5s16]

A step = one frame of the hot path: TLAS rebuild (the reference rebuilds it every frame, raytracer.cpp:76-85) ->
trace / shade -> rough_prepare, 20 blur sub-passes, postprocess, FXAA + 8-bit blit (+ NVLink gather for N > 1).
Default workload = BASELINE.json configs[1]: example scene, 1920x1080, numSamples 4 (2x2 SSAA), maxRecursions 5, FXAA.

  value   device-resident: instances / UBO already in HBM, nothing read back; CUDA-event time on the library's stream,
          max over ranks; L2 flushed (256 MiB memset) before every timed frame.
  e2e     the same frame through the public host API (rg_set_ubo + rg_set_instances from host memory, rg_render,
          read-back of the RGBA8 frame into pinned host memory), wall clock, max over ranks.
  N > 1   one process per GPU (torchrun); the frame is split into N column bands (+40 px halo re-traced per band so
          the post chain is bit-identical); every band is stored by the final kernel straight into rank 0's frame
          buffer over NVLink (CUDA IPC mapping).  Total work is fixed -> "scaling": "strong".
  --impl reference   the CPU restatement of the reference shaders (oracle/, kind "port": the real reference needs a
          Vulkan ray-tracing driver) on all host cores, same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (description, W, H, numSamples, maxRecursions, scene)
    "c1": ("example scene 640x360, numSamples 1, maxRecursions 5, FXAA (BASELINE configs[0])", 640, 360, 1, 5, "example"),
    "c2": ("example scene 1920x1080, numSamples 4 (2x2 SSAA), maxRecursions 5, FXAA (BASELINE configs[1])", 1920, 1080, 4, 5, "example"),
    "c3": ("28x28 mirror/glass sphere grid, 1 003 522 triangles as 785 instances, 1920x1080, numSamples 1, maxRecursions 8 (BASELINE configs[2])",
           1920, 1080, 1, 8, "spheres"),
    "c3flat": ("28x28 mirror/glass sphere grid flattened to one 1 003 522-triangle mesh, 1920x1080, numSamples 1, maxRecursions 8", 1920, 1080, 1, 8,
               "spheres_flat"),
    "c5": ("example scene 3840x2160, numSamples 4 (2x2 SSAA), maxRecursions 5, FXAA (north-star target line)", 3840, 2160, 4, 5, "example"),
    "c5s16": ("example scene 3840x2160, numSamples 16 (4x4 SSAA), maxRecursions 5, FXAA (BASELINE configs[4])", 3840, 2160, 16, 5, "example"),
}


def make_workload(name):
    from raygun_b200 import scene as S
    desc, W, H, ns, mr, kind = WORKLOADS[name]
    if kind == "example":
        sd, cam = S.load_example_scene()
        ubo = S.make_ubo(cam["view_inverse"], S.proj_inverse(W, H), ns, mr, cam["light_dir"])
    else:
        sd, vi = S.sphere_grid_scene(28, flattened=(kind == "spheres_flat"))
        ubo = S.make_ubo(vi, S.proj_inverse(W, H), ns, mr)
    return desc, W, H, sd, ubo


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        if self.is_alive():
            self.join(timeout=1.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": []}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def trace_algorithmic_bytes(tm, pixels):
    """per ray 32 B in + 16 B out, 80 B per wide node visited, 48 B per triangle tested, 48 B per instance
    transform applied, 184 B gathered per generic hit, 24 B of G-buffer per pixel."""
    return (tm["rays"] * 48 + tm["nodes_visited"] * 80 + tm["tris_tested"] * 48 + tm["instances_entered"] * 48 + tm["generic_hits"] * 184 + pixels * 24)


def run_reference(args, rank, world):
    """CPU arm: the oracle (port of the reference shaders) on all host cores; rank 0 only."""
    if rank != 0:
        return
    from oracle import oracle as O
    desc, W, H, sd, ubo = make_workload(args.workload)
    osc = O.OracleScene(sd)
    threads = O.max_threads()
    for _ in range(args.warmup):
        osc.render(ubo, W, H, O.FXAA, threads)
    t0 = time.perf_counter()
    rays = 0
    for _ in range(args.steps):
        r = osc.render(ubo, W, H, O.FXAA, threads)
        c = r["counters"]
        rays += c["primary"] + c["shadow"] + c["reflect"] + c["refract"]
    dt = time.perf_counter() - t0
    val = rays / dt / 1e6
    line = {"metric": "Mrays/s", "value": val, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": desc, "width": W, "height": H},
            "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": threads, "kind": "port",
                             "sample": f"{args.steps} full frames of the workload, CPU restatement of the reference shaders (oracle/), OpenMP"},
            "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--split", default="columns", choices=["columns", "rows"])
    ap.add_argument("--mgpu", default="partition", choices=["partition", "overdraw"],
                    help="N > 1: 'partition' = tiles traced round-robin, G-buffer pixels stored to their owners over NVLink; 'overdraw' = every band re-traces its halo")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import raygun_b200 as rg
    from raygun_b200.parallel import band_region, share_gather_handle, overdraw, attach_partition

    dist = None
    if world > 1:
        import torch.distributed as dist  # noqa: PLC0415
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)

    desc, W, H, sd, ubo = make_workload(args.workload)
    rt = rg.Raytracer(W, H, device=local_rank)
    if world > 1:
        rt.set_region(*band_region(W, H, rank, world, args.split))
    rt.setupModelBuffers(sd)
    rt.setupBottomLevelAS()
    inst_raw = rt.pack_instances(sd.inst_xform, sd.inst_meta)
    rt.setupTopLevelAS(inst_raw)
    rt.updateRenderTarget(ubo)

    if world > 1 and args.mgpu == "partition":
        attach_partition(dist, rt, rank, world)

    # gather target: rank 0's full-frame buffer, mapped into every other rank through CUDA IPC (NVLink peer stores)
    peer_ptr = None
    if world > 1:
        handle = None
        if rank == 0:
            handle, own = rt.gather_buffer_export()
            rt.set_gather_target(own)
        handle = share_gather_handle(dist, rank, handle)
        if rank != 0:
            peer_ptr = rt.gather_buffer_open(handle)
            rt.set_gather_target(peer_ptr)

    flags = rg.RG_FXAA
    d_inst = torch.from_numpy(inst_raw.view(np.int32).copy()).to(dev)
    d_ubo = torch.from_numpy(ubo.view(np.int32).copy()).to(dev)
    torch.cuda.synchronize()
    rt.set_ubo_device(d_ubo.data_ptr())

    def frame_resident():
        rt.set_instances_device(d_inst.data_ptr(), len(inst_raw))
        rt.doRaytracing(flags)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        rt.sync()

    for _ in range(args.warmup):
        frame_resident()
    barrier()

    # ---------------- timed: device-resident frames
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = rt.launch_count()
    dev_ms, sections = 0.0, {"as_build_ms": 0.0, "rt_only_ms": 0.0, "trace_kernel_ms": 0.0, "rough_ms": 0.0, "postproc_ms": 0.0}
    rays_local = 0
    barrier()
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        rt.flush_l2()
        rt.timer_begin()
        frame_resident()
        dev_ms += rt.timer_end()
        tm = rt.timings()
        rays_local += tm["rays"]
        for k in sections:
            sections[k] += tm[k]
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    launches = rt.launch_count() - launches0
    clocks = sampler.result()

    # ---------------- timed: end to end through the host API
    out_pinned = torch.empty((H, W, 4) if rank == 0 and world > 1 else (rt.region_size[1], rt.region_size[0], 4), dtype=torch.uint8, pin_memory=True)
    out_np = out_pinned.numpy()
    for _ in range(2):
        rt.render_frame(ubo, flags, inst_raw)
        rt.read_rgba8(out_np) if world == 1 else rt.sync()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rt.updateRenderTarget(ubo)          # 192 B host -> device
        rt.setupTopLevelAS(inst_raw)        # n x 64 B host -> device (pinned staging inside the library)
        rt.doRaytracing(flags)
        if world == 1:
            rt.read_rgba8(out_np)           # RGBA8 frame device -> pinned host
        else:
            rt.sync()
            dist.barrier()                  # all bands have landed in rank 0's frame buffer
            if rank == 0:
                rt.read_gathered_rgba8(out_np)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    h2d = 192 + inst_raw.nbytes
    d2h = W * H * 4

    # ---------------- reduce over ranks: max time, summed rays
    stats = torch.tensor([dev_ms, e2e_ms, wall_ms], dtype=torch.float64, device=dev)
    sums = torch.tensor([float(rays_local), float(launches)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    dev_ms, e2e_ms, wall_ms = (float(v) for v in stats.tolist())
    rays_total, launches_total = (float(v) for v in sums.tolist())

    # ---------------- roofline of the dominant kernel (the trace kernel the scheduler picked) from one instrumented, untimed frame (rank 0, N = 1 only)
    roofline = None
    # (every rank renders it: in partitioned mode a frame is a collective operation)
    rt.set_instances_device(d_inst.data_ptr(), len(inst_raw))
    rt.doRaytracing(flags | rg.RG_COUNT_TRAVERSAL)
    tmc = rt.timings()
    barrier()
    if rank == 0:
        peak, peak_src = load_peaks()
        px = rt.region_size[0] * rt.region_size[1]
        alg_bytes = trace_algorithmic_bytes(tmc, px)
        trace_ms = sections["trace_kernel_ms"] / args.steps
        achieved = alg_bytes / (trace_ms * 1e-3) / 1e9
        kernel = "k_trace_pool" if tm["trace_scheduler"] == rg.RG_SCHED_POOL else "k_trace_lanes"
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            with open(tp) as fh:
                traffic = json.load(fh).get(f"{kernel}:{args.workload}:n{world}")
        roofline = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": trace_ms,
                    "kernel_share_of_step": trace_ms / (dev_ms / args.steps),
                    "per_ray": {"nodes": tmc["nodes_visited"] / max(tmc["rays"], 1), "tris": tmc["tris_tested"] / max(tmc["rays"], 1)},
                    "note": "node / triangle fetches are served by L1/L2 (scene BVH is ~1 MB); DRAM traffic is the G-buffer"}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        osc = O.OracleScene(sd)
        threads = O.max_threads()
        osc.render(ubo, W, H, O.FXAA, threads)
        t0 = time.perf_counter(); n = 0; rays = 0
        while True:
            r = osc.render(ubo, W, H, O.FXAA, threads)
            c = r["counters"]; rays += c["primary"] + c["shadow"] + c["reflect"] + c["refract"]; n += 1
            if time.perf_counter() - t0 > 10.0 or n >= 50:
                break
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": rays / dt / 1e6, "unit": "Mrays/s", "cores": threads, "kind": "port",
                        "sample": f"{n} full frames of the same workload ({dt:.1f} s), CPU restatement of the reference shaders, OpenMP over rows",
                        "ms_per_frame": dt / n * 1e3}

    barrier()   # every rank: nobody may still be storing into rank 0's frame buffer
    sync_err = rt.sync_error()
    if peer_ptr is not None:
        rt.gather_buffer_close(peer_ptr)
    if world > 1 and args.mgpu == "partition":
        rt.peer_detach_all()
    if sync_err:
        raise RuntimeError(f"rank {rank}: cross-GPU barrier timed out waiting for rank {sync_err - 1}")
    if rank == 0:
        ms_step = dev_ms / args.steps
        line = {"metric": "Mrays/s", "value": rays_total / (dev_ms * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": desc, "width": W, "height": H, "split": (f"{world} {args.split} bands for the post chain; trace: 8x4 tiles dealt round-robin in chunks of 16, G-buffer pixels stored to their owners over NVLink"
                                     if args.mgpu == "partition" else f"{world} {args.split} bands, 40 px halo re-traced, overdraw x{overdraw(W, H, world, args.split):.3f}") if world > 1 else "none",
                           "l2": "flushed before every timed frame (256 MiB memset)", "tlas": "rebuilt every frame",
                           "trace_scheduler": ("pool" if tm["trace_scheduler"] == rg.RG_SCHED_POOL else "lanes") + " (RG_SCHED_AUTO: both timed during warm-up, faster kept)"},
                "fps": 1e3 / ms_step, "rays_per_frame": rays_total / args.steps,
                "sections_ms": {k: v / args.steps for k, v in sections.items()}, "wall_ms_per_step_incl_flush": wall_ms / args.steps,
                "clocks": clocks,
                "e2e": {"value": rays_total / (e2e_ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": e2e_ms / args.steps,
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": int(launches_total), "roofline": roofline}
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
