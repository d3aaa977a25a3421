#!/usr/bin/env python3
"""Writes tests/golden/oracle_c1_64x36.npz: a small frame of the example scene rendered by the CPU oracle.
A regression pin of the oracle itself; the reference cannot render here (no Vulkan ray-tracing driver)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import oracle as O  # noqa: E402
from raygun_b200 import scene as S  # noqa: E402

W, H = 64, 36
sd, _ = S.load_example_scene()
r = O.OracleScene(sd).render(S.example_ubo(W, H), W, H, O.FXAA)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "oracle_c1_64x36.npz")
np.savez_compressed(out, rgba8=r["rgba8"], inst=r["inst"], prim=r["prim"])
print("wrote", out)
