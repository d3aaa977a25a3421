#!/usr/bin/env python3
"""Generate tests/golden/text_scene.npz: ray-traced UI text from the reference's own font fixture.

Runs ONLY in the authoring container (needs /root/reference/resources).  The C++ host shim does the work exactly as the reference
does it: ResourceManager::loadFont (resource_manager.cpp:107-135) reads resources/fonts/NotoSans.obj (one object per glyph, one
vertex per face corner, shifted to x >= 0), ui::TextGenerator (ui/text.cpp:108-138) lays the string out as one entity per glyph,
RenderSystem packs the buffers and the entity walk emits the instances.  The .npz is the committed snapshot for the GPU box."""
import ctypes as C
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..")
RES = "/root/reference/resources"
TEXT = b"Raygun B200\n1.1 Grays/s"
W, H = 640, 360


def load(lib, text=TEXT, align=4):
    widths = np.zeros(128, np.float32); n = C.c_uint32()
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    h = C.c_void_p(lib.rgh_text_scene_load(RES.encode(), b"NotoSans", b"ui/text", text, align, W, H, p(widths), C.byref(n)))
    assert h.value, lib.rgh_last_error().decode()
    cnt = np.zeros(5, np.uint32); lib.rgh_scene_counts(h, p(cnt))
    v = np.zeros((cnt[0], 8), np.uint32); i = np.zeros(cnt[1], np.uint32); m = np.zeros((cnt[2], 16), np.uint32)
    r = np.zeros((cnt[3], 4), np.uint32); inst = np.zeros((cnt[4], 16), np.uint32); ubo = np.zeros(48, np.uint32)
    lib.rgh_scene_copy(h, p(v), p(i), p(m), p(r), p(inst), p(ubo))
    lib.rgh_scene_free(h)
    return dict(vertices=v, indices=i, materials=m, meshes=r, instances=inst, ubo=ubo, widths=widths.view(np.uint32), glyphs=np.uint32(n.value))


def main():
    lib = C.CDLL(os.path.join(ROOT, "raygun_b200", "libraygun_host.so"))
    lib.rgh_last_error.restype = C.c_char_p
    lib.rgh_text_scene_load.restype = C.c_void_p
    d = load(lib)
    # known answers of the font fixture (computed here from the reference's NotoSans.obj; Assimp itself is not available)
    assert int(d["glyphs"]) == 124
    assert len(d["instances"]) == 21 and len(d["meshes"]) == 16
    print("vertices", d["vertices"].shape, "crc32", hex(zlib.crc32(d["vertices"].tobytes())), "instances", len(d["instances"]))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "text_scene.npz"), text=np.frombuffer(TEXT, np.uint8), **d)


if __name__ == "__main__":
    sys.exit(main())
