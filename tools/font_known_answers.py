#!/usr/bin/env python3
"""Known answers of the reference's font fixture, computed INDEPENDENTLY of the C++ loader (raygun_b200/host/resource_loader.cpp):
a 40-line numpy reader of resources/fonts/NotoSans.obj following what Assimp's OBJ importer gives raygun::Entity (one vertex per
face corner in file order, no welding; all faces of this file are triangles) and ResourceManager::loadFont
(raygun/resource_manager.cpp:107-135: shift every glyph so that its left edge is x = 0, width = max x - min x).
Writes tests/golden/font_known_answers.json: per glyph code the corner count, the binary32 width (as bits) and the CRC32 of the
shifted float32 positions.  Runs only where /root/reference exists; the JSON travels."""
import json, os, sys, zlib
import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
OBJ = "/root/reference/resources/fonts/NotoSans.obj"


def read_glyphs(path=OBJ):
    pos, glyphs, cur = [], {}, None
    for line in open(path):
        t = line.split()
        if not t:
            continue
        if t[0] == "o":
            cur = glyphs.setdefault(int(t[1]), [])
        elif t[0] == "v":
            pos.append([np.float32(float(x)) for x in t[1:4]])    # strtod, then rounded once to binary32
        elif t[0] == "f":
            assert len(t) == 4, "the fixture is triangulated"
            for c in t[1:]:
                k = int(c.split("/")[0])
                cur.append(k - 1 if k > 0 else len(pos) + k)
    P = np.array(pos, np.float32)
    out = {}
    for code, idx in glyphs.items():
        p = P[np.array(idx)].copy()
        lo = p[:, 0].min()
        p[:, 0] = p[:, 0] - lo                                   # binary32 subtraction, as the reference's loop does
        width = np.float32(p[:, 0].max() - p[:, 0].min())
        out[code] = {"corners": len(idx), "width_bits": int(width.view(np.uint32)), "crc32_positions": zlib.crc32(np.ascontiguousarray(p).tobytes())}
    return out


if __name__ == "__main__":
    g = read_glyphs()
    json.dump({"file": "resources/fonts/NotoSans.obj", "glyphs": {str(k): v for k, v in sorted(g.items())}}, open(os.path.join(ROOT, "tests", "golden", "font_known_answers.json"), "w"), indent=0)
    print(len(g), "glyphs; corners", sum(v["corners"] for v in g.values()))
