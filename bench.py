#!/usr/bin/env python3
"""bench.py -- Mrays/s and ms/frame of the per-frame ray-tracing path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c1|c2|c3|c3flat|c4|c5|c5s16] [--no-also]

A step = one frame of the hot path: TLAS rebuild (the reference rebuilds it every frame, raytracer.cpp:76-85) ->
trace / shade -> rough_prepare, the 20 blur sub-passes (11 launches over the active pixels), postprocess, FXAA + 8-bit blit (+ NVLink gather for N > 1).

Workload (both arms use the same rule, so their `config` objects are identical):
  N = 1   BASELINE.json configs[1] (c2): example scene, 1920x1080, numSamples 4 (2x2 SSAA), maxRecursions 5, FXAA.
  N > 1   BASELINE.json configs[4] at the north-star sample count (c5): example scene, 3840x2160, numSamples 4 -- the
          multi-GPU configuration of BASELINE.md (a 5 ms 1080p frame is the worst possible scaling subject).
  The other configs are reported beside the headline by the N = 1 run under "also" (c3, c4, c5, c5s16; measured in the same
  process, device-resident, L2 flushed, CUDA events) so that every BASELINE config has a driver-visible number.

  value   device-resident: instances / UBO already in HBM, nothing read back; CUDA-event time on the library's stream,
          max over ranks; L2 flushed (256 MiB memset) before every timed frame.
  e2e     the same frame through the public host API (rg_set_ubo + rg_set_instances from host memory, rg_render, the RGBA8
          frame in pinned host memory when the step ends), wall clock, max over ranks; L2 flushed before every frame too.
          N = 1: the final kernel stores the frame straight into the pinned host buffer (rg_set_gather_target on host memory:
          the 8.3 MB cross PCIe as the kernel's own stores, no separate copy; --e2e-copy times the cudaMemcpy read-back
          instead); after the timed loop the buffer is compared with a regular rg_read_rgba8 of the same frame.
          N > 1: ONE frame in POSIX shared memory, page-locked and mapped by every rank (rg_host_frame_register): every GPU stores
          its band into it over its own PCIe link and rank 0 reads the assembled frame without a copy; compared after the loop
          with the frame gathered over NVLink into rank 0's device buffer (--e2e-copy times that path instead).
  N > 1   one process per GPU (torchrun); the frame is split into N column bands for the post chain, the trace is dealt
          round-robin in tile chunks and every finished pixel is stored straight into its owners' G-buffers over NVLink
          (CUDA IPC mappings); every band's RGBA8 pixels are stored by the final kernel into rank 0's frame buffer.
          Total work is fixed -> "scaling": "strong".
  --impl reference   the CPU restatement of the reference shaders (oracle/, kind "port": the real reference needs a
          Vulkan ray-tracing driver, which neither this container nor the GPU box has -- profiles/r02_vulkan_probe.txt)
          on ALL host cores (the affinity mask; an inherited OMP_NUM_THREADS=1 from torchrun is ignored), same workload.
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
    "c4": ("10 000 bouncing ball instances + floor, closed-form animation, TLAS rebuilt every frame, 1920x1080, numSamples 1, maxRecursions 5 "
           "(BASELINE configs[3])", 1920, 1080, 1, 5, "balls"),
    "c5": ("example scene 3840x2160, numSamples 4 (2x2 SSAA), maxRecursions 5, FXAA (BASELINE configs[4] at the north-star sample count)", 3840, 2160, 4, 5,
           "example"),
    "c5s16": ("example scene 3840x2160, numSamples 16 (4x4 SSAA), maxRecursions 5, FXAA (BASELINE configs[4])", 3840, 2160, 16, 5, "example"),
}
ISSUE_LANES_PER_SM_CLK = 4 * 32   # 4 schedulers x 32 lanes: the SM's thread-instruction issue peak per clock


def default_workload(world):
    return "c2" if world == 1 else "c5"


def host_cores():
    """Cores this process may use (torchrun exports OMP_NUM_THREADS=1; the CPU arm must not inherit that)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def make_workload(name):
    """-> (description, W, H, SceneData, ubo); for "c4" the SceneData is frame 0 (see make_animation)."""
    from raygun_b200 import scene as S
    desc, W, H, ns, mr, kind = WORKLOADS[name]
    if kind == "example":
        sd, cam = S.load_example_scene()
        ubo = S.make_ubo(cam["view_inverse"], S.proj_inverse(W, H), ns, mr, cam["light_dir"])
    elif kind == "balls":
        balls = S.AnimatedBalls(100)
        sd = balls.scene(0.0)
        ubo = S.make_ubo(balls.view_inverse, S.proj_inverse(W, H), ns, mr)
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


def load_ncu_table():
    """profiles/traffic.json: per 'kernel:workload:nN' the ncu figures of ONE launch (dram_bytes, lts_bytes, thread_inst, warp_inst)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return {}
    with open(p) as fh:
        return json.load(fh)


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
    threads = host_cores()
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
                             "sample": f"{args.steps} full frames of the workload, CPU restatement of the reference shaders (oracle/), OpenMP over rows"},
            "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def measure_also(rg, name, device, frames=8, warmup=4):
    """One of the other BASELINE configs on ONE GPU, device-resident like `value`: -> dict for the "also" object."""
    import torch
    from raygun_b200 import scene as S
    desc, W, H, sd, ubo = make_workload(name)
    rt = rg.Raytracer(W, H, device=device)
    t0 = time.perf_counter()
    rt.load_scene(sd)
    rt.sync()
    load_s = time.perf_counter() - t0
    rt.updateRenderTarget(ubo)
    dev = torch.device("cuda", device)
    if name == "c4":   # a different set of transforms every frame (t = n / 60), all resident in HBM before the timed region
        balls = S.AnimatedBalls(100)
        sets = [torch.from_numpy(rt.pack_instances(balls.instances(n / 60.0), balls.meta).view(np.int32).copy()).to(dev) for n in range(frames + warmup)]
        n_inst = len(balls.meta)
    else:
        raw = rt.pack_instances(sd.inst_xform, sd.inst_meta)
        sets = [torch.from_numpy(raw.view(np.int32).copy()).to(dev)]
        n_inst = len(raw)
    torch.cuda.synchronize()
    ms, secs, rays = [], {"as_build_ms": [], "rt_only_ms": [], "postproc_ms": []}, 0
    for n in range(frames + warmup):
        rt.flush_l2()
        rt.timer_begin()
        rt.set_instances_device(sets[n % len(sets)].data_ptr(), n_inst)
        rt.doRaytracing(rg.RG_FXAA)
        t = rt.timer_end()
        tm = rt.timings()
        if n >= warmup:
            ms.append(t); rays = tm["rays"]
            for k in secs:
                secs[k].append(tm[k])
    m = float(np.median(ms))
    out = {"workload": desc, "ms_per_step": m, "fps": 1e3 / m, "value": rays / m / 1e3, "unit": "Mrays/s", "rays_per_frame": rays, "frames": frames,
           "trace_scheduler": "pool" if tm["trace_scheduler"] == rg.RG_SCHED_POOL else "lanes",
           "sections_ms": {k: float(np.median(v)) for k, v in secs.items()}, "instances": n_inst, "scene_load_s": load_s}
    if name in ("c3", "c3flat"):
        sec = rays - tm["rays_primary"]
        out["secondary_only_grays_s"] = sec / float(np.median(secs["rt_only_ms"])) / 1e6   # north-star: >= 1 on one GPU
    if name == "c4":
        out["as_build_ms"] = out["sections_ms"]["as_build_ms"]   # reported separately from the trace
        # per-frame BLAS refit of a vertex-wobbled copy of the flattened 1 003 522-triangle grid, vertices already in HBM
        sd2, _ = S.sphere_grid_scene(28, flattened=True)
        rt2 = rg.Raytracer(64, 64, device=device)
        t0 = time.perf_counter(); rt2.load_scene(sd2); rt2.sync()
        out["blas_build_1M_tris_s_incl_upload"] = time.perf_counter() - t0
        base = sd2.vertices.view(np.float32)
        dv = []
        for k in range(2):
            v = sd2.vertices.copy()
            v.view(np.float32)[:, 1] = base[:, 1] + 0.05 * np.sin(3.0 * base[:, 0] + k)
            dv.append(torch.from_numpy(v.view(np.int32).copy()).to(dev))
        torch.cuda.synchronize()
        ref = []
        for k in range(6):
            rt2.refitBottomLevelAS_device(0, dv[k & 1].data_ptr())
            rt2.sync()
            ref.append(rt2.timings()["as_build_ms"])
        out["blas_refit_1M_tris_ms"] = float(np.median(ref[2:]))
        rt2.close()
    rt.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--e2e-copy", action="store_true", help="time the e2e read-back as a cudaMemcpy (N > 1: of rank 0's gathered frame) instead of the zero-copy host frame")
    ap.add_argument("--split", default="columns", choices=["columns", "rows"])
    ap.add_argument("--mgpu", default="partition", choices=["partition", "overdraw"],
                    help="N > 1: 'partition' = tiles traced round-robin, G-buffer pixels stored to their owners over NVLink; 'overdraw' = every band re-traces its halo")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the other BASELINE configs (N = 1 only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    explicit_workload = args.workload is not None
    if args.workload is None:
        args.workload = default_workload(max(world, args.gpus))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import raygun_b200 as rg
    from raygun_b200.parallel import band_region, share_gather_handle, overdraw, attach_partition, open_shared_host_frame

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
    peer_ptr = own = None
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
    trace_ms_frames = []
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
        trace_ms_frames.append(tm["trace_kernel_ms"])
        for k in sections:
            sections[k] += tm[k]
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    launches = rt.launch_count() - launches0
    clocks = sampler.result()

    # ---------------- timed: end to end through the host API (same L2 flush per frame as above, so e2e >= value at every N)
    out_pinned = torch.empty((H, W, 4) if rank == 0 and world > 1 else (rt.region_size[1], rt.region_size[0], 4), dtype=torch.uint8, pin_memory=True)
    out_np = out_pinned.numpy()
    zero_copy = world == 1 and not args.e2e_copy
    shared = None
    if zero_copy:
        rt.set_gather_target(out_pinned.data_ptr())   # pinned host memory is device-addressable (UVA): k_fxaa_blit writes it directly
    elif world > 1 and not args.e2e_copy:
        # one frame in POSIX shared memory, mapped by every rank: each GPU stores its band over its own PCIe link, rank 0 reads no copy
        def agree(ok):
            t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            return bool(t.item())
        shared = open_shared_host_frame(dist, rt, rank, W, H, agree)   # None on every rank if any rank could not: the copy path below
        if shared is not None:
            rt.set_gather_target(shared.device_ptr)
    for _ in range(2):
        rt.render_frame(ubo, flags, inst_raw)
        rt.read_rgba8(out_np) if world == 1 and not zero_copy else rt.sync()
    barrier()
    t0 = time.perf_counter()
    flush_s = 0.0
    for _ in range(args.steps):
        tf = time.perf_counter()
        rt.flush_l2()
        rt.sync()
        flush_s += time.perf_counter() - tf     # the flush is hygiene, not part of the frame: its wall time is taken out again
        rt.updateRenderTarget(ubo)          # 192 B host -> device
        rt.setupTopLevelAS(inst_raw)        # n x 64 B host -> device (pinned staging inside the library)
        rt.doRaytracing(flags)
        if zero_copy:
            rt.sync()                       # the frame is in out_pinned: the last kernel stored it there
        elif world == 1:
            rt.read_rgba8(out_np)           # RGBA8 frame device -> pinned host
        elif shared is not None:
            rt.sync()
            dist.barrier()                  # all bands have landed in the shared host frame (shared.array on rank 0)
        else:
            rt.sync()
            dist.barrier()                  # all bands have landed in rank 0's frame buffer
            if rank == 0:
                rt.read_gathered_rgba8(out_np)
    barrier()
    e2e_ms = (time.perf_counter() - t0 - flush_s) * 1e3
    h2d = 192 + inst_raw.nbytes
    d2h = W * H * 4
    if zero_copy:   # untimed: the frame the kernel stored into host memory is the frame a read-back returns
        rt.set_gather_target(0)
        if not np.array_equal(out_np, rt.read_rgba8()):
            raise RuntimeError("bench: the zero-copy host frame differs from rg_read_rgba8")
    shared_used = shared is not None
    if shared is not None:   # untimed: the same frame once more through rank 0's device buffer (NVLink peer stores) and a read-back
        host_frame = shared.array.copy() if rank == 0 else None
        rt.set_gather_target(own if rank == 0 else peer_ptr)
        barrier()
        rt.render_frame(ubo, flags, inst_raw)
        barrier()
        if rank == 0:
            rt.read_gathered_rgba8(out_np)
            if not np.array_equal(out_np, host_frame):
                raise RuntimeError("bench: the shared host frame differs from the frame gathered in rank 0's device buffer")
        shared.close()

    # ---------------- reduce over ranks: max time, summed rays
    tk = float(np.mean(trace_ms_frames))
    stats = torch.tensor([dev_ms, e2e_ms, wall_ms, tk, -tk], dtype=torch.float64, device=dev)
    sums = torch.tensor([float(rays_local), float(launches)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    dev_ms, e2e_ms, wall_ms, tk_max, tk_min_neg = (float(v) for v in stats.tolist())
    rays_total, launches_total = (float(v) for v in sums.tolist())

    # ---------------- roofline of the dominant kernel (the trace kernel the scheduler picked) from one instrumented, untimed frame
    roofline = None
    # (every rank renders it: in partitioned mode a frame is a collective operation)
    rt.set_instances_device(d_inst.data_ptr(), len(inst_raw))
    rt.doRaytracing(flags | rg.RG_COUNT_TRAVERSAL)
    tmc = rt.timings()
    barrier()
    if rank == 0:
        peak, peak_src = load_peaks()
        ncu = load_ncu_table()
        px = rt.region_size[0] * rt.region_size[1]
        alg_bytes = trace_algorithmic_bytes(tmc, px)
        trace_ms = sections["trace_kernel_ms"] / args.steps
        achieved = alg_bytes / (trace_ms * 1e-3) / 1e9
        kernel = "k_trace_pool" if tm["trace_scheduler"] == rg.RG_SCHED_POOL else "k_trace_lanes"
        rec = ncu.get(f"{kernel}:{args.workload}:n{world}")
        if not isinstance(rec, dict):
            rec = {"dram_bytes": rec} if rec else {}
        sm_clk = (clocks["sm_mhz"] or 1965.0) * 1e6
        n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
        issue_peak = n_sm * ISSUE_LANES_PER_SM_CLK * sm_clk / 1e9           # G thread-instructions / s
        # The trace kernels are bound by instruction issue, not by bytes (ncu: DRAM < 10 % of peak, the 1 MB BVH lives in L1 / L2):
        # the roofline is the SM's thread-instruction issue rate; the SURVEY 8d byte figure is reported beside it.
        roofline = {"bound": "issue", "kernel": kernel, "kernel_ms": trace_ms, "kernel_share_of_step": trace_ms / (dev_ms / args.steps),
                    "unit": "Gthread-instr/s", "peak": issue_peak,
                    "peak_source": f"{n_sm} SMs x 4 schedulers x 32 lanes x {sm_clk / 1e6:.0f} MHz (SM clock sampled during the timed region)",
                    "achieved": None, "frac": None, "traffic": rec.get("dram_bytes"),
                    "bytes": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                              "algorithmic_bytes_per_launch": alg_bytes, "l2_bytes_per_launch_ncu": rec.get("lts_bytes"),
                              "note": "SURVEY 8d counted bytes; 9/10 of them are L1 / L2 hits, DRAM traffic is the G-buffer + per-sample scratch"},
                    "per_ray": {"nodes": tmc["nodes_visited"] / max(tmc["rays"], 1), "tris": tmc["tris_tested"] / max(tmc["rays"], 1),
                                "instances": tmc["instances_entered"] / max(tmc["rays"], 1)}}
        if rec.get("thread_inst"):
            roofline["achieved"] = rec["thread_inst"] / (trace_ms * 1e-3) / 1e9
            roofline["frac"] = roofline["achieved"] / issue_peak
            roofline["thread_instructions_per_launch_ncu"] = rec["thread_inst"]
            roofline["warp_instructions_per_launch_ncu"] = rec.get("warp_inst")
        # the other kernels of the step: algorithmic bytes over their CUDA-event time, ncu DRAM / L2 bytes beside them
        if world == 1:
            active = None
            try:
                active = int((rt.read_image(rg.IMG_TRANSITIONS) > 0).sum())
            except Exception:  # noqa: BLE001
                pass
            rough_ms = sections["rough_ms"] / args.steps
            post_ms = (sections["postproc_ms"] - sections["rough_ms"]) / args.steps
            as_ms = sections["as_build_ms"] / args.steps
            act = active if active is not None else px
            # rough_prepare: 33 B / pixel + 4 B per listed pixel; a blur sub-pass touches only the listed (active) pixels: 3 texels read +
            # 1 written (8 B each) + the list entry + the transition byte -- the 9 fused launches do an H and a V sub-pass each
            rb = px * 33 + act * 4 + 20 * act * (4 * 8 + 4 + 1)
            n_i = len(inst_raw)
            kernels = [
                {"kernel": "k_rough_prepare + 9 x k_rough_blur_list<fused H+V> + 2 x k_rough_blur_list<single>", "bound": "hbm", "ms": rough_ms,
                 "algorithmic_bytes": rb, "achieved": rb / (rough_ms * 1e-3) / 1e9,
                 "peak": peak, "unit": "GB/s", "frac": rb / (rough_ms * 1e-3) / 1e9 / peak, "active_blur_pixels": active,
                 "note": "12 small launches over the list of active pixels: bound by launch latency, not by bytes",
                 "ncu": {k: ncu.get(f"{k}:{args.workload}:n1") for k in ("k_rough_prepare", "k_rough_blur_list")}},
                {"kernel": "k_postprocess + k_fxaa_blit", "bound": "hbm", "ms": post_ms, "algorithmic_bytes": px * 36, "achieved": px * 36 / (post_ms * 1e-3) / 1e9,
                 "peak": peak, "unit": "GB/s", "frac": px * 36 / (post_ms * 1e-3) / 1e9 / peak,
                 "note": "k_fxaa_blit is instruction bound (ncu: 75 % of the issue slots, ~460 thread instructions per pixel: the sampler emulation)",
                 "ncu": {k: ncu.get(f"{k}:{args.workload}:n1") for k in ("k_postprocess", "k_fxaa_blit")}},
                {"kernel": "k_tlas_fused", "bound": "latency (one block)", "ms": as_ms, "algorithmic_bytes": n_i * (64 + 64 + 64 + 56 + 64 + 32 + 72 + 80),
                 "ncu": ncu.get(f"k_tlas_fused:{args.workload}:n1")},
                {"kernel": "k_order_hist + k_order_scatter", "bound": "latency", "ms": None,
                 "note": "side stream, beside the post chain: not on the frame's critical path", "ncu": ncu.get(f"k_order_tiles:{args.workload}:n1")},
            ]
            roofline["other_kernels"] = kernels

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        osc = O.OracleScene(sd)
        threads = host_cores()
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
    rt.close()

    also = None
    if rank == 0 and world == 1 and not args.no_also and not explicit_workload:
        also = {}
        for name in ("c3", "c4", "c5", "c5s16"):
            try:
                also[name] = measure_also(rg, name, local_rank)
            except Exception as e:  # noqa: BLE001  (a failing side measurement must not take the headline line with it)
                also[name] = {"error": f"{type(e).__name__}: {e}"}

    if rank == 0:
        ms_step = dev_ms / args.steps
        line = {"metric": "Mrays/s", "value": rays_total / (dev_ms * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": desc, "width": W, "height": H},
                "run": {"split": (f"{world} {args.split} bands for the post chain; trace: 8x4 tiles dealt round-robin in chunks of 16, G-buffer pixels stored to their owners over NVLink"
                                  if args.mgpu == "partition" else f"{world} {args.split} bands, 40 px halo re-traced, overdraw x{overdraw(W, H, world, args.split):.3f}") if world > 1 else "none",
                        "l2": "flushed before every timed frame, device-timed and e2e (256 MiB memset)", "tlas": "rebuilt every frame",
                        "trace_scheduler": ("pool" if tm["trace_scheduler"] == rg.RG_SCHED_POOL else "lanes") + " (RG_SCHED_AUTO: both timed during warm-up, faster kept)",
                        "trace_kernel_ms_per_rank": {"max": tk_max, "min": -tk_min_neg}},
                "fps": 1e3 / ms_step, "rays_per_frame": rays_total / args.steps,
                "sections_ms": {k: v / args.steps for k, v in sections.items()}, "wall_ms_per_step_incl_flush": wall_ms / args.steps,
                "clocks": clocks,
                "e2e": {"value": rays_total / (e2e_ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": e2e_ms / args.steps,
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "frame_to_host": ("stored by the final kernel into pinned host memory (zero-copy), verified against rg_read_rgba8 after the loop" if zero_copy
                                          else ("cudaMemcpy read-back into pinned host memory" if world == 1 else
                                                ("every rank's final kernel stores its band into ONE frame in POSIX shared memory (page-locked, mapped by every rank): no copy; "
                                                 "verified against the frame gathered over NVLink after the loop" if shared_used else
                                                 "peer stores into rank 0's frame, read back by rank 0")))},
                "gpu_launches": int(launches_total), "roofline": roofline}
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        if also:
            line["also"] = also
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
