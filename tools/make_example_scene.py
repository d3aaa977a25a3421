#!/usr/bin/env python3
"""Generate tests/golden/example_scene.npz from the reference's own asset fixtures.

Runs ONLY in the authoring container (needs /root/reference); the .npz it writes is the
committed fixture that travels to the GPU box.  It reproduces what the reference hands
the GPU for the example scene:

  * Collada ingestion as Assimp does it with aiProcess_Triangulate only
    (raygun/entity.cpp:88-90): one output vertex per <p> index tuple (3 per triangle,
    no welding), one aiMesh per <triangles> group in file order, merged per node
    (raygun/entity.cpp:66-81, raygun/render/mesh.cpp:52-62).  Vertex = 32 bytes
    (resources/shaders/vertex.def:3-7), matIndex = aiMesh::mMaterialIndex
    (raygun/entity.cpp:44).
  * Materials by name -> resources/materials/<name>.rgmat.json
    (raygun/resource_manager.cpp:35-50, raygun/material.cpp:83-107), 64-byte
    gpu::Material (resources/shaders/gpu_material.def:11-26).  Every model of a file
    carries ALL materials of that file (raygun/entity.cpp:109-116).
  * Packing into one vertex / index / material buffer
    (raygun/render/render_system.cpp:270-330) and the instance list + offset table
    (raygun/render/acceleration_structure.cpp:34-85).  Instance transforms and camera
    matrices come from tests/golden/glm_golden.json (made by the reference's vendored
    GLM, see oracle/ref_recipe/glm_golden.cpp).

It also asserts the known-answer data of SURVEY.md Appendix C (triangle counts, CRC32
of expanded positions, AABBs).
"""
import json
import os
import struct
import sys
import xml.etree.ElementTree as ET
import zlib

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden", "example_scene.npz")
NS = {"c": "http://www.collada.org/2005/11/COLLADASchema"}

MAT_DEFAULT = dict(diffuse=[1.0, 0.0, 1.0], transparency=0.0, specular=[1.0, 1.0, 1.0], reflectivity=0.0,
                   roughness=0.0, ior=1.0, effectId=0, rayConsumption=1, emission=0.0)


def pack_material(d):
    m = dict(MAT_DEFAULT)
    for k, v in d.items():
        if k in ("type", "basedOn", "staticFriction", "dynamicFriction"):
            continue
        assert k in m, k
        m[k] = v
    f = np.float32
    return struct.pack("<3ff3ffffIIffff", *map(f, m["diffuse"]), f(m["transparency"]), *map(f, m["specular"]),
                       f(m["reflectivity"]), f(m["roughness"]), f(m["ior"]), int(m["effectId"]),
                       int(m["rayConsumption"]), f(m["emission"]), 0.0, 0.0, 0.0)


def load_material(name):
    path = os.path.join(REF, "resources", "materials", name + ".rgmat.json")
    if not os.path.exists(path):  # underscore fallback, resource_manager.cpp:41-50
        i = name.find("_")
        path = os.path.join(REF, "resources", "materials", name[:i], name[i + 1:] + ".rgmat.json")
    with open(path) as fh:
        return pack_material(json.load(fh))


def parse_dae(path):
    """-> (material names in Assimp order, [node dict(name, matrix16, positions, normals, mat_index, groups)])"""
    root = ET.parse(path).getroot()
    # Assimp's ColladaParser keeps the material library in a std::map keyed by the
    # material ID, so aiScene::mMaterials comes out sorted by ID (render-invariant).
    mats = sorted(((m.get("id"), m.get("name")) for m in root.findall(".//c:library_materials/c:material", NS)))
    mat_ids = [i for i, _ in mats]
    geoms = {}
    for g in root.findall(".//c:library_geometries/c:geometry", NS):
        mesh = g.find("c:mesh", NS)
        sources = {}
        for s in mesh.findall("c:source", NS):
            fa = s.find("c:float_array", NS)
            stride = int(s.find("c:technique_common/c:accessor", NS).get("stride"))
            sources["#" + s.get("id")] = np.array(fa.text.split(), dtype=np.float32).reshape(-1, stride)
        verts = mesh.find("c:vertices", NS)
        vpos = {"#" + verts.get("id"): verts.find("c:input[@semantic='POSITION']", NS).get("source")}
        groups = []
        for tri in mesh.findall("c:triangles", NS):
            inputs = {i.get("semantic"): (int(i.get("offset")), i.get("source")) for i in tri.findall("c:input", NS)}
            stride = max(o for o, _ in inputs.values()) + 1
            p = np.array(tri.find("c:p", NS).text.split(), dtype=np.int64).reshape(-1, stride)
            assert p.shape[0] == 3 * int(tri.get("count"))
            pos = sources[vpos[inputs["VERTEX"][1]]][p[:, inputs["VERTEX"][0]]]
            nrm = sources[inputs["NORMAL"][1]][p[:, inputs["NORMAL"][0]]]
            groups.append(dict(material=tri.get("material"), positions=pos, normals=nrm))
        geoms["#" + g.get("id")] = groups
    nodes = []
    for n in root.findall(".//c:library_visual_scenes/c:visual_scene/c:node", NS):
        mat16 = np.array(n.find("c:matrix", NS).text.split(), dtype=np.float32)
        ig = n.find("c:instance_geometry", NS)
        bind = {im.get("symbol"): im.get("target")[1:] for im in ig.findall(".//c:instance_material", NS)}
        pos, nrm, mi, counts = [], [], [], []
        for grp in geoms[ig.get("url")]:
            pos.append(grp["positions"]); nrm.append(grp["normals"])
            idx = mat_ids.index(bind[grp["material"]])
            mi.append(np.full(len(grp["positions"]), idx, np.uint32))
            counts.append(len(grp["positions"]) // 3)
        nodes.append(dict(name=n.get("name"), matrix=mat16, positions=np.concatenate(pos), normals=np.concatenate(nrm),
                          mat_index=np.concatenate(mi), group_tris=counts))
    return [n for _, n in mats], nodes


def bits_to_f32(lst):
    return np.array(lst, dtype=np.uint32).view(np.float32)


def main():
    glm = json.load(open(os.path.join(HERE, "..", "tests", "golden", "glm_golden.json")))
    room_mats, room_nodes = parse_dae(os.path.join(REF, "resources", "models", "room.dae"))
    ball_mats, ball_nodes = parse_dae(os.path.join(REF, "resources", "models", "ball.dae"))

    # ---- known answers, SURVEY.md Appendix C
    kat = {"Raygun": (19974, 0xc5b5c3bf, [6658]), "ph3_games": (31284, 0xceed8f6a, [4686, 5742]),
           "room": (1644, 0x6b194de2, [396, 152]), "ball": (3840, 0xc6b74a2f, [1280])}
    for n in room_nodes + ball_nodes:
        nv, crc, groups = kat[n["name"]]
        assert len(n["positions"]) == nv, (n["name"], len(n["positions"]))
        assert zlib.crc32(n["positions"].astype("<f4").tobytes()) == crc, n["name"]
        assert n["group_tris"] == groups

    models = room_nodes + ball_nodes            # registration order (entity.cpp:104-120, ball.cpp:11-22)
    model_mats = [room_mats] * 3 + [ball_mats]  # every child model carries all materials of its file
    V, I, meshes, M = [], [], [], []
    mesh_names, mat_off, voff, ioff = [], [], 0, 0
    for n, mats in zip(models, model_mats):
        nv = len(n["positions"])
        v = np.zeros((nv, 8), np.float32)
        v[:, 0:3] = n["positions"]; v[:, 3] = n["mat_index"].view(np.float32); v[:, 4:7] = n["normals"]
        V.append(v); I.append(np.arange(nv, dtype=np.uint32))
        meshes.append((voff, nv, ioff, nv)); voff += nv; ioff += nv
        mat_off.append(len(M)); M += [load_material(m) for m in mats]
        mesh_names.append(n["name"])
    assert len(M) == 16 and voff == 56742

    inst_names = ["Raygun", "ph3_games", "room", "Ball"]  # DFS order, acceleration_structure.cpp:63-85
    xforms = np.stack([bits_to_f32(glm["instance_" + k]) for k in inst_names])
    inst = np.array([(i, meshes[i][0], meshes[i][2], mat_off[i]) for i in range(4)], np.uint32)

    np.savez_compressed(
        OUT,
        vertices=np.concatenate(V).view(np.uint32),      # (N, 8) raw 32-byte Vertex records
        indices=np.concatenate(I),
        meshes=np.array(meshes, np.uint32),              # (vtx_off, vtx_cnt, idx_off, idx_cnt) in elements
        materials=np.frombuffer(b"".join(M), np.uint8).reshape(-1, 64).copy(),
        instance_xform=xforms.view(np.uint32),           # (4, 12) row-major 3x4 object->world
        instance_mesh_voff_ioff_moff=inst,
        view_inverse=np.array(glm["viewInverse"], np.uint32),   # column-major mat4 bit patterns
        light_dir=np.array(glm["lightDir"], np.uint32),
        mesh_names=np.array(mesh_names), instance_names=np.array(inst_names),
    )
    print("wrote", os.path.normpath(OUT), os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    sys.exit(main())
