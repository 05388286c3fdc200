"""Model interchange over the C ABI (include/me_modal.h, "Model interchange"): the reference's `.modal` files
(src/audio/ModalModelFile.{h,cpp}) and the JSON of MeshEditorModalSolve (tests/ModalSolveTool.cpp:84-123) that
glTF_PhysicalAudio embeds as KHR_audio_rigid_bodies modal models. Host-only."""
from __future__ import annotations

import ctypes as C
import json

import numpy as np

from ._lib import ME_OK, MeModalFileExtras, check, lib
from .modal import _read_result, material, solve_handle

LN1000 = float(np.float32(3) * np.float32(np.log(np.float32(10.0))))


def _u32(a):
    return np.ascontiguousarray(a if a is not None else [], np.uint32).reshape(-1)


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
