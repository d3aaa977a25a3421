import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import raygun_b200 as rg
from raygun_b200 import scene as S
from test_gpu_traversal import _random_rays
sd, _ = S.load_example_scene()
rt = rg.Raytracer(64, 36); rt.load_scene(sd)
rays = _random_rays(sd, 20000, 1)
tuv, ip = rt.debug_trace_rays(rays)
np.savez_compressed("gpurun_out/rays.npz", rays=rays, tuv=tuv, ip=ip)
