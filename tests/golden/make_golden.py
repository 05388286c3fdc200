#!/usr/bin/env python
"""Regenerates tests/golden/*.npz — run in the build container only (needs /root/reference and oracle/_ref).

Each fixture is one of the reference's committed golden modal models (glTF_PhysicalAudio/samples/**/*.gltf, the
`KHR_audio_rigid_bodies.modalModels` entries written by the macOS reference binary MeshEditorModalSolve,
tests/ModalSolveTool.cpp:47-124) together with the exact INPUTS that solve saw, replayed here:

  surface mesh  : the reference's own generator functions, imported from glTF_PhysicalAudio/samples/generate.py
                  (_grid_box / sphere / union_surface, :110-145,:252-275,:623-660), vertices written with Python repr
                  and re-read as float32 the way tests/LoadObj.h:28-34 does, welded in first-seen order;
  tet mesh      : the UNMODIFIED reference tetrahedralizer (oracle/_ref/libme_ref_tet.so), Quality off (ModalSolveTool.cpp:72);
  golden arrays : frequencies, decayRates, positions, mode-major shapes, massProperties decoded from the glTF buffers.

The fixtures carry no reference source, only data. /root/reference does not exist on the GPU box; tests read the .npz.
"""
import base64
import ctypes as C
import importlib.util
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/glTF_PhysicalAudio/samples"


def load_generator():
    spec = importlib.util.spec_from_file_location("ref_generate", os.path.join(REF, "generate.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def obj_roundtrip(verts, tris):
    """f"v {x} {y} {z}" with Python repr -> tinyobj float32 -> weld identical positions, first seen (LoadObj.h:28-39)."""
    v32 = np.array([[np.float32(float(repr(float(c)))) for c in p] for p in verts], np.float32)
    welded, remap, out = {}, np.zeros(len(v32), np.uint32), []
    for i, p in enumerate(v32):
        key = tuple(p.tolist())
        if key not in welded:
            welded[key] = len(out)
            out.append(p)
        remap[i] = welded[key]
    tri = remap[np.asarray(tris, np.int64).ravel()].astype(np.uint32)
    return np.asarray(out, np.float32), tri


def tetrahedralize(points32, tri, quality=False):
    L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libme_ref_tet.so"))
    L.ref_tet_error.restype = C.c_char_p
    pts = np.ascontiguousarray(points32.astype(np.float64))  # GenerateTets promotes float positions to double (Tets.cpp:265-268)
    tri = np.ascontiguousarray(tri, np.uint32)
    rc = L.ref_tetrahedralize(pts.ctypes.data_as(C.c_void_p), C.c_uint32(len(pts)), tri.ctypes.data_as(C.c_void_p), C.c_uint32(len(tri)), C.c_int(int(quality)))
    if rc != 0:
        raise RuntimeError(L.ref_tet_error().decode())
    nv, nt = L.ref_tet_point_count(), L.ref_tet_count()
    out_p, out_t = np.zeros((nv, 3), np.float64), np.zeros((nt, 4), np.uint32)
    L.ref_tet_copy(out_p.ctypes.data_as(C.c_void_p), out_t.ctypes.data_as(C.c_void_p))
    return out_p, out_t


def read_accessor(g, idx):
    acc = g["accessors"][idx]
    view = g["bufferViews"][acc["bufferView"]]
    raw = base64.b64decode(g["buffers"][view["buffer"]]["uri"].split(",", 1)[1])
    comp = {5126: np.float32, 5123: np.uint16, 5125: np.uint32}[acc["componentType"]]
    width = {"SCALAR": 1, "VEC3": 3}[acc["type"]]
    off = view.get("byteOffset", 0) + acc.get("byteOffset", 0)
    a = np.frombuffer(raw, comp, acc["count"] * width, off)
    return a.reshape(-1, width) if width > 1 else a.copy()


def golden_model(path, name):
    g = json.load(open(os.path.join(REF, path)))
    ext = g["extensions"]["KHR_audio_rigid_bodies"]
    model = next(m for m in ext["modalModels"] if m["name"] == name)
    material = ext["acousticMaterials"][model["material"]]
    freqs = read_accessor(g, model["frequencies"])
    pos = read_accessor(g, model["positions"])
    shapes = read_accessor(g, model["shapes"]).reshape(len(freqs), len(pos), 3)  # mode-major
    mp = model["massProperties"]
    return dict(
        golden_freqs=freqs.astype(np.float32), golden_decay=read_accessor(g, model["decayRates"]).astype(np.float32),
        golden_positions=pos.astype(np.float32), golden_shapes=shapes.astype(np.float32), golden_indices=read_accessor(g, model["indices"]).astype(np.uint32),
        golden_mass=np.float64(mp["mass"]), golden_com=np.asarray(mp["centerOfMass"], np.float32), golden_inertia=np.asarray(mp["inertiaDiagonal"], np.float32),
        material=np.array([material["density"], material["youngsModulus"], material["poissonRatio"], material["alpha"], material["beta"]], np.float64),
    )


def icosphere(radius, recursion_level):
    """IcoSphere primitive restated from src/mesh/Primitives.h:60-110 in float32 (positions, triangle indices)."""
    f32 = np.float32
    t = f32((1 + 5 ** 0.5) / 2)
    verts = [np.array(v, f32) for v in [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t), (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]]

    def normalize(v):
        return (v * (f32(1) / np.sqrt(f32(np.dot(v, v))))).astype(f32)

    verts = [normalize(v) for v in verts]
    tris = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
            (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    cache = {}

    def mid(a, b):
        key = (min(a, b), max(a, b))
        if key not in cache:
            verts.append(normalize(((verts[a] + verts[b]) / f32(2)).astype(f32)))
            cache[key] = len(verts) - 1
        return cache[key]

    for _ in range(recursion_level):
        new = []
        for a, b, c in tris:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            new += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        tris = new
    return (np.asarray(verts, f32) * f32(radius)).astype(f32), np.asarray(tris, np.uint32).ravel()


def make_config1():
    """BASELINE.json configs[0]: IcoSphere radius 0.1 m, 4 subdivisions, reference tetrahedralizer with
    Quality on. Mesh only (no golden answer exists for it); the parity tests solve it with the oracle."""
    surf, tri = icosphere(0.1, 4)
    points, tets = tetrahedralize(surf, tri, quality=True)
    np.savez_compressed(os.path.join(HERE, "..", "meshes", "icosphere_c1.npz"), points=points, tets=tets, surface=surf, triangles=tri)
    print(f"icosphere_c1: {len(surf)} surface verts, {len(points)} points, {len(tets)} tets")


def main():
    os.makedirs(os.path.join(HERE, "..", "meshes"), exist_ok=True)
    make_config1()
    gen = load_generator()
    sphere = lambda r: (lambda p, n, u, idx: (p, [tuple(idx[t:t + 3]) for t in range(0, len(idx), 3)]))(*gen.sphere(r))
    cases = [
        # fixture, gltf, model name, surface builder, modes, max_freq
        ("bar_ceramic", "test/ContactDuration/a_5g.gltf", None, lambda: gen._grid_box(*gen.BarHalf, 0.02), 10, 16000.0),
        ("bar_steel", "Pile.gltf", "Bar", lambda: gen._grid_box(*gen.BarHalf, 0.02), 30, 16000.0),
        ("cube_ceramic", "Pile.gltf", "Cube", lambda: gen._grid_box(0.04, 0.04, 0.04, 0.01), 30, 60000.0),
        ("slab_ceramic", "Pile.gltf", "Slab", lambda: gen._grid_box(0.07, 0.012, 0.05, 0.012), 30, 60000.0),
        ("platform_ceramic", "Pile.gltf", "Platform", lambda: gen._grid_box(0.3, 0.03, 0.3, 0.05), 30, 16000.0),
        ("bracket_steel", "Pile.gltf", "Bracket", lambda: gen.union_surface(gen.BracketBoxes, 0.005), 30, 60000.0),
        ("marble_glass", "Pile.gltf", "Marble", lambda: sphere(gen.SphereR), 30, 60000.0),
        # the remaining solved models of the sample tree: the glass bar, the machined ground plate, the ceramic sphere
        ("bar_glass", "test/surface/PressedRing/a_NoGrip.gltf", "Solved box", lambda: gen._grid_box(*gen.BarHalf, 0.02), 30, 16000.0),
        ("ground_ceramic", "test/surface/SurfaceRadiates/a_Scrape.gltf", "Ground Machined", lambda: gen._grid_box(*gen.GroundHalf, 0.04), 30, 16000.0),
        ("sphere_ceramic", "test/PatchGrowth/a_SphereLight.gltf", "Solved sphere", lambda: sphere(gen.SphereR), 30, 60000.0),
        # the steel bead (its document links the model to acoustic material 0, Ceramic, but the generator solved it as the scene's STEEL probe,
        #  generate.py:1048 `probe=solved_modes_sphere(SphereR, STEEL, max_freq=60000.0)`: the solve's material is what the fixture keeps)
        ("sphere_steel", "test/AccelerationNoise/a_SteelBead.gltf", "Solved sphere", lambda: sphere(gen.SphereR), 30, 60000.0, gen.STEEL),
    ]
    for fixture, gltf, name, build, modes, max_freq, *solved_as in cases:
        if name is None:
            g = json.load(open(os.path.join(REF, gltf)))
            name = g["extensions"]["KHR_audio_rigid_bodies"]["modalModels"][0]["name"]
        gold = golden_model(gltf, name)
        if solved_as:
            m = solved_as[0]
            gold["material"] = np.array([m["density"], m["youngsModulus"], m["poissonRatio"], m["alpha"], m["beta"]], np.float64)
        verts, tris = build()
        surf, tri = obj_roundtrip(verts, tris)
        points, tets = tetrahedralize(surf, tri)
        # `modes` requested = the generator's default 30 (generate.py:316); the bar's 10 kept modes are what survived the 16 kHz window.
        np.savez_compressed(os.path.join(HERE, fixture + ".npz"), points=points, tets=tets, surface=surf, triangles=tri, num_modes=np.int32(30),
                            min_freq=np.float32(20.0), max_freq=np.float32(max_freq), source=np.array(f"{gltf}#{name}"), **gold)
        print(f"{fixture}: {len(surf)} surface verts, {len(points)} points, {len(tets)} tets, {len(gold['golden_freqs'])} golden modes, f1 {gold['golden_freqs'][0]:.2f} Hz")


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    main()
