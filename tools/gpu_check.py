#!/usr/bin/env python3
"""Developer check on a GPU box: render one config on the GPU and with the oracle, print parity + timing.
(Test infrastructure: imports oracle/.)  Usage: python tools/gpu_check.py [W H S [maxrec]]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import raygun_b200 as rg  # noqa: E402
from raygun_b200 import scene as S  # noqa: E402
from oracle import oracle as O  # noqa: E402


def psnr8(a, b):
    d = a[..., :3].astype(np.float64) - b[..., :3].astype(np.float64)
    mse = np.mean(d * d)
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


def main():
    W, H, Sn = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (640, 360, 1)
    maxrec = int(sys.argv[4]) if len(sys.argv) > 4 else 5
    out_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    sd, _ = S.load_example_scene()
    ubo = S.example_ubo(W, H, Sn, maxrec)

    rt = rg.Raytracer(W, H)
    t0 = time.time(); rt.load_scene(sd); rt.sync(); print("load_scene s", time.time() - t0, rt.debug_bvh_stats())

    # Morton / sort parity
    for m in range(len(sd.meshes)):
        n = int(sd.meshes[m, 3]) // 3
        k, o = rt.debug_blas_sort(m, n)
        ck, co, _ = O.morton_triangles(sd, m)
        print(f"mesh {m}: n={n} keys_equal={np.array_equal(k, ck[co])} order_equal={np.array_equal(o, co)}")

    flags = rg.RG_FXAA | rg.RG_DEBUG_IDS
    for it in range(3):
        rt.render_frame(ubo, flags)
        rt.sync()
    tm = rt.timings()
    print("timings", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in tm.items()})
    print("Mrays/s (trace only)", tm["rays"] / tm["rt_only_ms"] / 1e3)
    img = rt.read_rgba8()
    inst, prim = rt.read_ids()

    osc = O.OracleScene(sd)
    t0 = time.time(); ref = osc.render(ubo, W, H, O.FXAA); print("oracle s", time.time() - t0, ref["counters"])
    ok = (inst == ref["inst"]) & (prim == ref["prim"])
    print("primary id match", ok.mean(), "mismatch px", int((~ok).sum()))
    d = np.abs(img[..., :3].astype(int) - ref["rgba8"][..., :3].astype(int)).max(axis=2)
    print("final: psnr", psnr8(img, ref["rgba8"]), "frac<=2", (d <= 2).mean(), "max", d.max())
    rt_imgs = {}
    for name, which in (("base", rg.IMG_BASE), ("normal", rg.IMG_NORMAL), ("rough", rg.IMG_ROUGH), ("roughA", rg.IMG_ROUGH_A), ("final", rg.IMG_FINAL)):
        rt_imgs[name] = rt.read_image(which)
        a = O.f16_to_f32(rt_imgs[name]); b = O.f16_to_f32(ref[name])
        fin = np.isfinite(a) & np.isfinite(b)
        dd = np.abs(np.where(fin, a - b, 0))
        print(f"  {name}: max abs diff {dd.max():.5f} mean {dd.mean():.7f} nonfinite gpu/ref {int((~np.isfinite(a)).sum())}/{int((~np.isfinite(b)).sum())}")
    tr = rt.read_image(rg.IMG_TRANSITIONS)
    print("  transitions equal frac", (tr == ref["transitions"]).mean())
    print("ray counters gpu", {k: tm[k] for k in ("rays_primary", "rays_shadow", "rays_reflect", "rays_refract", "sky_lookups")})

    # post chain bit-exactness on the oracle's G-buffer
    gb = ref["gbuffer"]
    rt.debug_upload_gbuffer(gb["base"], gb["normal"], gb["rough"])
    rt.debug_run_post(rg.RG_FXAA)
    rt.sync()
    for name, which in (("final", rg.IMG_FINAL), ("base", rg.IMG_BASE), ("roughA", rg.IMG_ROUGH_A), ("roughB", rg.IMG_ROUGH_B)):
        a = rt.read_image(which)
        print(f"  post-on-oracle-gbuffer {name}: bit-equal frac {(a == ref[name]).mean():.6f}")
    print("  post-on-oracle-gbuffer transitions equal", (rt.read_image(rg.IMG_TRANSITIONS) == ref["transitions"]).mean())
    print("  post-on-oracle-gbuffer rgba8 equal", (rt.read_rgba8() == ref["rgba8"]).mean())
    if W * H <= 640 * 360:
        np.savez_compressed(os.path.join(out_dir, f"gpu_{W}x{H}_s{Sn}.npz"), rgba8=img, inst=inst, prim=prim,
                            **{n: rt_imgs[n] for n in rt_imgs})
    try:
        from PIL import Image
        Image.fromarray(img[..., :3]).save(os.path.join(out_dir, f"gpu_{W}x{H}_s{Sn}.png"))
        Image.fromarray(ref["rgba8"][..., :3]).save(os.path.join(out_dir, f"oracle_{W}x{H}_s{Sn}.png"))
        Image.fromarray((np.minimum(d, 25) * 10).astype(np.uint8)).save(os.path.join(out_dir, f"diff_{W}x{H}_s{Sn}.png"))
    except Exception as e:  # noqa: BLE001
        print("png save failed", e)


if __name__ == "__main__":
    main()
