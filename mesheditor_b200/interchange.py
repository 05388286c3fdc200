"""Model interchange over the C ABI (include/me_modal.h, "Model interchange"): the reference's `.modal` files
(src/audio/ModalModelFile.{h,cpp}) and the JSON of MeshEditorModalSolve (tests/ModalSolveTool.cpp:84-123) that
glTF_PhysicalAudio embeds as KHR_audio_rigid_bodies modal models; and the generation-job glue either side of the solve
(include/me_modal.h, "Generation-job glue": src/audio/AudioSystem.cpp:838-862, src/mesh/Tets.cpp:268-293). Host-only."""
from __future__ import annotations

import ctypes as C
import json

import numpy as np

from ._lib import ME_OK, MeModalFileExtras, check, lib
from .modal import _read_result, material, solve_handle

LN1000 = float(np.float32(3) * np.float32(np.log(np.float32(10.0))))


def _u32(a):
    return np.ascontiguousarray(a if a is not None else [], np.uint32).reshape(-1)


def _take_u32(ptr, n):
    try:
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint32)), (n,)).copy() if n else np.zeros(0, np.uint32)
    finally:
        lib().me_bytes_free(ptr)


def desired_solve_vertices(requested, num_vertices) -> np.ndarray:
    """DesiredSolveVertices (AudioSystem.cpp:667-671): evenly spaced excitation vertices."""
    out, n = C.c_void_p(), C.c_uint32()
    check(lib().me_desired_solve_vertices(int(requested), int(num_vertices), C.byref(out), C.byref(n)))
    return _take_u32(out, n.value)


def sample_surface_triangles(triangle_indices, vertex_count, excitation_vertices) -> np.ndarray:
    """SampleSurfaceTriangles (AudioSystem.cpp:701-746): triangles over the excitation vertices (as indices into
    `excitation_vertices`), from the mesh's own triangulation collapsed onto them."""
    tri, ex, out, n = _u32(triangle_indices), _u32(excitation_vertices), C.c_void_p(), C.c_uint32()
    check(lib().me_sample_surface_triangles(tri.ctypes.data, len(tri), int(vertex_count), ex.ctypes.data, len(ex), C.byref(out), C.byref(n)))
    return _take_u32(out, n.value)


def compact_excitation_vertices(vertices, sample_point_of) -> np.ndarray:
    """CompactExcitationVertices (AudioSystem.cpp:750-757): the first excitation vertex of every sample point."""
    v, sp, out, n = _u32(vertices), _u32(sample_point_of), C.c_void_p(), C.c_uint32()
    check(lib().me_compact_excitation_vertices(v.ctypes.data, len(v), sp.ctypes.data, len(sp), C.byref(out), C.byref(n)))
    return _take_u32(out, n.value)


def relabel_sample_triangles(triangles, sample_point_of) -> np.ndarray:
    """RelabelSampleTriangles (AudioSystem.cpp:761-769): the sample surface over the sample points the solve merged."""
    tri, sp, out, n = _u32(triangles), _u32(sample_point_of), C.c_void_p(), C.c_uint32()
    check(lib().me_relabel_sample_triangles(tri.ctypes.data, len(tri), sp.ctypes.data, len(sp), C.byref(out), C.byref(n)))
    return _take_u32(out, n.value)


def build_tet_mesh_data(points, tets, scale=(1.0, 1.0, 1.0)):
    """BuildTetMeshData (Tets.cpp:268-293) -> (positions float32 [V][3] in the node's local frame, edge index pairs uint32)."""
    pts = np.ascontiguousarray(points, np.float64).reshape(-1, 3)
    tt = np.ascontiguousarray(tets, np.uint32).reshape(-1, 4)
    sc, pos, edges, n = (C.c_float * 3)(*scale), C.c_void_p(), C.c_void_p(), C.c_uint32()
    check(lib().me_build_tet_mesh_data(pts.ctypes.data, len(pts), tt.ctypes.data, len(tt), sc, C.byref(pos), C.byref(edges), C.byref(n)))
    try:
        positions = np.ctypeslib.as_array(C.cast(pos, C.POINTER(C.c_float)), (3 * len(pts),)).reshape(-1, 3).copy() if len(pts) else np.zeros((0, 3), np.float32)
    finally:
        lib().me_bytes_free(pos)
    return positions, _take_u32(edges, n.value)


class ModalModel:
    """A MeModalResult handle plus what ModalModelData carries beyond it (ModalModelFile.h:13-20)."""

    def __init__(self, handle, status=ME_OK, *, vertices=None, indices=None, baked_scale=(1.0, 1.0, 1.0), tet_positions=None, tet_edge_indices=None, solved_material=None,
                 solved_min_mode_freq=20.0, solved_max_mode_freq=16000.0, solved_num_modes=30, tet_inputs_hash=0, solved_vertices=None, file_handle=None):
        self._h, self._file, self.status = handle, file_handle, status
        self.vertices, self.indices, self.solved_vertices, self.tet_edge_indices = _u32(vertices), _u32(indices), _u32(solved_vertices), _u32(tet_edge_indices)
        self.tet_positions = np.ascontiguousarray(tet_positions if tet_positions is not None else np.zeros((0, 3)), np.float32).reshape(-1, 3)
        self.baked_scale = tuple(float(v) for v in baked_scale)
        self.solved_material = material(solved_material) if solved_material is not None else material((0, 0, 0, 0, 0))
        self.solved_min_mode_freq, self.solved_max_mode_freq = float(solved_min_mode_freq), float(solved_max_mode_freq)
        self.solved_num_modes, self.tet_inputs_hash = int(solved_num_modes), int(tet_inputs_hash)
        self.result = _read_result(handle, status)

    def __del__(self):
        L = lib()
        if getattr(self, "_file", None):
            L.me_modal_file_free(self._file)
            self._file = None
        if getattr(self, "_h", None):
            L.me_modal_result_free(self._h)
            self._h = None

    @classmethod
    def solve(cls, points, tets, mat, excite_positions, baked_scale=(1.0, 1.0, 1.0), config=None, **extras):
        """modal::mesh2modes on the GPU, keeping the result for serialisation (AudioSystem.cpp:850 -> SaveModalModelFile)."""
        h, status = solve_handle(points, tets, mat, excite_positions, baked_scale, config)
        return cls(h, status, baked_scale=baked_scale, solved_material=mat, **extras)

    @classmethod
    def generate(cls, tet_points, tets, mat, surface_positions, triangle_indices, vertices, node_scale=(1.0, 1.0, 1.0), config=None, tet_inputs_hash=0):
        """The modal generation job after tetrahedralization (AudioSystem.cpp:830-862): solve at the excitation vertices'
        positions, then fill everything SaveModalModelFile stores - ModalModes::Vertices / Indices / BakedScale, the eigen
        summary's solve settings, the display TetMeshData - so that to_bytes() is the `.modal` file of the job."""
        from .modal import solver_config

        config = config if config is not None else solver_config()
        surface = np.ascontiguousarray(surface_positions, np.float32).reshape(-1, 3)
        vertices = _u32(vertices)
        sample_triangles = sample_surface_triangles(triangle_indices, len(surface), vertices)
        h, status = solve_handle(tet_points, tets, mat, surface[vertices], node_scale, config)
        count = C.c_uint32()
        sp = lib().me_modal_result_sample_point_of_excitation(h, C.byref(count))
        sample_point_of = np.ctypeslib.as_array(sp, (count.value,)).astype(np.uint32) if count.value else np.zeros(0, np.uint32)
        tet_positions, tet_edges = build_tet_mesh_data(tet_points, tets, node_scale)
        return cls(h, status, vertices=compact_excitation_vertices(vertices, sample_point_of), indices=relabel_sample_triangles(sample_triangles, sample_point_of), baked_scale=node_scale,
                   tet_positions=tet_positions, tet_edge_indices=tet_edges, solved_material=mat, solved_min_mode_freq=config.min_mode_freq, solved_max_mode_freq=config.max_mode_freq,
                   solved_num_modes=config.num_modes, tet_inputs_hash=tet_inputs_hash, solved_vertices=vertices)

    @classmethod
    def from_bytes(cls, data: bytes):
        """LoadModalModelFile (ModalModelFile.cpp:52-58)."""
        buf = np.frombuffer(data, np.uint8)
        h, f = C.c_void_p(), C.c_void_p()
        check(lib().me_modal_file_parse(buf.ctypes.data, len(buf), C.byref(h), C.byref(f)))
        x = MeModalFileExtras()
        check(lib().me_modal_file_extras(f, C.byref(x)))

        def arr(ptr, n, dtype):
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint32 if dtype == np.uint32 else C.c_float)), (n,)).astype(dtype).copy() if n else np.zeros(0, dtype)

        m = x.solved_material
        return cls(h, ME_OK, vertices=arr(x.vertices, x.n_vertices, np.uint32), indices=arr(x.indices, x.n_indices, np.uint32), baked_scale=tuple(x.baked_scale),
                   tet_positions=arr(x.tet_positions_xyz, 3 * x.n_tet_positions, np.float32), tet_edge_indices=arr(x.tet_edge_indices, x.n_tet_edge_indices, np.uint32),
                   solved_material=(m.density, m.young_modulus, m.poisson_ratio, m.alpha, m.beta), solved_min_mode_freq=x.solved_min_mode_freq, solved_max_mode_freq=x.solved_max_mode_freq,
                   solved_num_modes=x.solved_num_modes, tet_inputs_hash=x.tet_inputs_hash, solved_vertices=arr(x.solved_vertices, x.n_solved_vertices, np.uint32), file_handle=f)

    def _extras(self):
        x = MeModalFileExtras()
        x.vertices, x.n_vertices = self.vertices.ctypes.data, len(self.vertices)
        x.indices, x.n_indices = self.indices.ctypes.data, len(self.indices)
        x.baked_scale = (C.c_float * 3)(*self.baked_scale)
        x.tet_positions_xyz, x.n_tet_positions = self.tet_positions.ctypes.data, len(self.tet_positions)
        x.tet_edge_indices, x.n_tet_edge_indices = self.tet_edge_indices.ctypes.data, len(self.tet_edge_indices)
        x.solved_material = self.solved_material
        x.solved_min_mode_freq, x.solved_max_mode_freq = self.solved_min_mode_freq, self.solved_max_mode_freq
        x.solved_num_modes, x.tet_inputs_hash = self.solved_num_modes, self.tet_inputs_hash
        x.solved_vertices, x.n_solved_vertices = self.solved_vertices.ctypes.data, len(self.solved_vertices)
        return x

    def to_bytes(self) -> bytes:
        """The bytes SaveModalModelFile writes (ModalModelFile.cpp:15-22)."""
        x, out, size = self._extras(), C.c_void_p(), C.c_uint64()
        check(lib().me_modal_file_serialize(self._h, C.byref(x), C.byref(out), C.byref(size)))
        try:
            return C.string_at(out, size.value)
        finally:
            lib().me_bytes_free(out)

    def rescaled(self, mat, config=None):
        """modal::RescaleModes (mesh2modes.h:88, AudioSystem.cpp:595-616): this model's solved eigenpairs re-derived under `mat`
        (same Poisson ratio), as a new ModalModel that still carries the SOLVED summary and material, like the reference's
        ModalModelData after a material edit. Raises MeError where the reference returns nullopt."""
        from .modal import solver_config

        cfg = config if config is not None else solver_config(num_modes=self.solved_num_modes, min_mode_freq=self.solved_min_mode_freq, max_mode_freq=self.solved_max_mode_freq)
        m, h = material(mat), C.c_void_p()
        check(lib().me_rescale_modes(self._h, C.byref(self.solved_material), C.byref(m), C.byref(cfg), C.byref(h)))
        sm = self.solved_material
        return ModalModel(h, ME_OK, vertices=self.vertices, indices=self.indices, baked_scale=self.baked_scale, tet_positions=self.tet_positions, tet_edge_indices=self.tet_edge_indices,
                          solved_material=(sm.density, sm.young_modulus, sm.poisson_ratio, sm.alpha, sm.beta), solved_min_mode_freq=self.solved_min_mode_freq,
                          solved_max_mode_freq=self.solved_max_mode_freq, solved_num_modes=self.solved_num_modes, tet_inputs_hash=self.tet_inputs_hash, solved_vertices=self.solved_vertices)

    def solve_json(self, triangle_indices=()) -> str:
        """What MeshEditorModalSolve prints for this model (tests/ModalSolveTool.cpp:84-123)."""
        tri, out = _u32(triangle_indices), C.c_void_p()
        check(lib().me_modal_solve_json(self._h, tri.ctypes.data, len(tri), C.byref(out)))
        try:
            return C.string_at(out).decode()
        finally:
            lib().me_bytes_free(out)


def khr_modal_model(solve_json: str) -> dict:
    """The KHR_audio_rigid_bodies modalModel fields glTF_PhysicalAudio's generator takes from that JSON (generate.py:623-687):
    float32 arrays, shapes mode-major [mode][point][3]."""
    d = json.loads(solve_json)
    freqs = np.asarray(d["frequencies"], np.float32)
    positions = np.asarray(d["positions"], np.float32).reshape(-1, 3)
    return dict(frequencies=freqs, decayRates=np.asarray(d["decayRates"], np.float32), positions=positions,
                shapes=np.asarray(d["shapes"], np.float32).reshape(len(freqs), len(positions), 3), indices=np.asarray(d["indices"], np.uint32), mass=float(d["mass"]),
                centerOfMass=np.asarray(d["centerOfMass"], np.float32), inertiaDiagonal=np.asarray(d["inertiaDiagonal"], np.float32))


def bank_modes(model: dict):
    """A modal model as the synthesis bank takes it (AddModalObject + TuneModalObject): T60 = ln 1000 / decayRate,
    shapes back to [point][mode][3]."""
    rates = model["decayRates"]
    t60s = np.where(rates > 0, np.float32(LN1000) / np.where(rates > 0, rates, 1).astype(np.float32), np.float32(0)).astype(np.float32)
    return dict(freqs=model["frequencies"], t60s=t60s, shapes=np.ascontiguousarray(np.transpose(model["shapes"], (1, 0, 2))), positions=model["positions"], indices=model["indices"])


# ---- KHR_audio_rigid_bodies glTF documents (glTF_PhysicalAudio/extensions/2.0/Khronos/KHR_audio_rigid_bodies) ------------------
# The reference reads and writes these through fastgltf in its scene code (out of this path's scope); the generator of the
# golden samples (glTF_PhysicalAudio/samples/generate.py:623-687) writes them from MeshEditorModalSolve's JSON. These two
# functions are that last hop, so a solve on the GPU can be carried into a document the reference loads, and the reference's
# committed golden models can be read back (tests/golden/make_golden.py reads the same fields).

_COMPONENT = {5126: np.float32, 5123: np.uint16, 5125: np.uint32}
_WIDTH = {"SCALAR": 1, "VEC3": 3}


def _accessor(doc, index, buffers):
    import base64

    acc = doc["accessors"][index]
    view = doc["bufferViews"][acc["bufferView"]]
    b = view["buffer"]
    if b not in buffers:
        uri = doc["buffers"][b]["uri"]
        if not uri.startswith("data:"):
            raise ValueError("only embedded (data: URI) buffers are read")
        buffers[b] = base64.b64decode(uri.split(",", 1)[1])
    width = _WIDTH[acc["type"]]
    a = np.frombuffer(buffers[b], _COMPONENT[acc["componentType"]], acc["count"] * width, view.get("byteOffset", 0) + acc.get("byteOffset", 0))
    return a.reshape(-1, width).copy() if width > 1 else a.copy()


def read_gltf_modal_models(doc) -> list:
    """The modalModels of a glTF document (a dict, or a path to a .gltf with embedded buffers), each in the form
    khr_modal_model() returns plus `name` and `material` (the acousticMaterials entry: density, youngsModulus, poissonRatio,
    alpha, beta). Required fields per the schema: frequencies, decayRates, positions, shapes."""
    if not isinstance(doc, dict):
        with open(doc) as f:
            doc = json.load(f)
    ext = doc.get("extensions", {}).get("KHR_audio_rigid_bodies", {})
    buffers, out = {}, []
    for m in ext.get("modalModels", []):
        freqs = _accessor(doc, m["frequencies"], buffers).astype(np.float32)
        positions = _accessor(doc, m["positions"], buffers).astype(np.float32).reshape(-1, 3)
        model = dict(name=m.get("name", ""), frequencies=freqs, decayRates=_accessor(doc, m["decayRates"], buffers).astype(np.float32), positions=positions,
                     shapes=_accessor(doc, m["shapes"], buffers).astype(np.float32).reshape(len(freqs), len(positions), 3),
                     indices=_accessor(doc, m["indices"], buffers).astype(np.uint32).reshape(-1) if "indices" in m else np.zeros(0, np.uint32))
        if "material" in m:
            model["material"] = dict(ext["acousticMaterials"][m["material"]])
        mp = m.get("massProperties")
        if mp:
            model.update(mass=float(mp["mass"]), centerOfMass=np.asarray(mp.get("centerOfMass", (0, 0, 0)), np.float32), inertiaDiagonal=np.asarray(mp.get("inertiaDiagonal", (0, 0, 0)), np.float32))
        out.append(model)
    return out


def write_gltf_modal_models(models, asset_generator="mesheditor_b200") -> dict:
    """A minimal glTF 2.0 document carrying `models` (dicts as read_gltf_modal_models / khr_modal_model give them, each
    optionally with `name` and a `material` dict) in extensions.KHR_audio_rigid_bodies, all arrays in one embedded buffer:
    FLOAT SCALAR frequencies / decayRates, FLOAT VEC3 positions / mode-major shapes, UNSIGNED_INT SCALAR indices."""
    import base64

    blob, views, accessors, materials, out_models = bytearray(), [], [], [], []

    def add(array, component, kind):
        a = np.ascontiguousarray(array, _COMPONENT[component]).reshape(-1)
        while len(blob) % 4:
            blob.append(0)
        views.append(dict(buffer=0, byteOffset=len(blob), byteLength=a.nbytes))
        blob.extend(a.tobytes())
        accessors.append(dict(bufferView=len(views) - 1, componentType=component, count=a.size // _WIDTH[kind], type=kind))
        return len(accessors) - 1

    for m in models:
        entry = dict(name=m.get("name", ""), frequencies=add(m["frequencies"], 5126, "SCALAR"), decayRates=add(m["decayRates"], 5126, "SCALAR"), positions=add(m["positions"], 5126, "VEC3"),
                     shapes=add(m["shapes"], 5126, "VEC3"))
        if len(m.get("indices", ())):
            entry["indices"] = add(m["indices"], 5125, "SCALAR")
        if "material" in m:
            if m["material"] not in materials:
                materials.append(m["material"])
            entry["material"] = materials.index(m["material"])
        if "mass" in m:
            entry["massProperties"] = dict(mass=float(m["mass"]), centerOfMass=[float(v) for v in m["centerOfMass"]], inertiaDiagonal=[float(v) for v in m["inertiaDiagonal"]])
        out_models.append(entry)
    ext = dict(modalModels=out_models)
    if materials:
        ext["acousticMaterials"] = materials
    return dict(asset=dict(version="2.0", generator=asset_generator), extensionsUsed=["KHR_audio_rigid_bodies"], buffers=[dict(byteLength=len(blob), uri="data:application/octet-stream;base64," + base64.b64encode(bytes(blob)).decode())],
                bufferViews=views, accessors=accessors, extensions=dict(KHR_audio_rigid_bodies=ext))
