"""TEST INFRASTRUCTURE ONLY (oracle). The reference's own `.modal` serialisation: ModalModelData through zpp::bits with the
reference's struct definitions and glm hooks, compiled into oracle/_ref/libme_ref_audio.so (oracle/ref_audio_driver.cpp).
Parity pinned: it IS the reference's archive; tests/golden/interchange/ holds its output for boxes without _ref."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

REF_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libme_ref_audio.so")


def have_ref():
    return os.path.exists(REF_SO)


class RefModalFlat(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("n_modes", "n_points", "n_vertices", "n_indices", "n_tet_positions", "n_tet_edges", "n_eigen", "n_solved_vertices")] + [
        (n, C.c_void_p) for n in ("freqs", "t60s", "shapes", "positions", "vertices", "indices")] + [
        ("original_fundamental", C.c_float), ("baked_scale", C.c_float * 3), ("mass", C.c_double), ("com", C.c_float * 3), ("inertia", C.c_float * 3), ("quat_wxyz", C.c_float * 4),
        ("tet_positions", C.c_void_p), ("tet_edges", C.c_void_p), ("eigenvalues", C.c_void_p), ("summary_shapes", C.c_void_p), ("material", C.c_double * 5), ("min_freq", C.c_float),
        ("max_freq", C.c_float), ("num_modes", C.c_uint32), ("tet_hash", C.c_uint64), ("solved_vertices", C.c_void_p)]


def random_model(seed, n_modes=7, n_points=5, n_eigen=11):
    """A seeded ModalModelData as plain arrays (every field populated, sizes deliberately unequal)."""
    rng = np.random.default_rng(seed)
    f32 = np.float32
    q = rng.standard_normal(4)
    return dict(
        freqs=np.sort(rng.uniform(50, 9000, n_modes)).astype(f32), t60s=rng.uniform(0.01, 3, n_modes).astype(f32), shapes=rng.standard_normal((n_points, n_modes, 3)).astype(f32),
        positions=rng.standard_normal((n_points, 3)).astype(f32), vertices=rng.integers(0, 1000, n_points).astype(np.uint32), indices=rng.integers(0, n_points, 3 * max(n_points - 2, 0)).astype(np.uint32),
        original_fundamental=f32(rng.uniform(50, 500)), baked_scale=rng.uniform(0.5, 2, 3).astype(f32), mass=float(rng.uniform(0.01, 20)), com=rng.standard_normal(3).astype(f32),
        inertia=rng.uniform(1e-4, 1, 3).astype(f32), quat_wxyz=(q / np.linalg.norm(q)).astype(f32), tet_positions=rng.standard_normal((9, 3)).astype(f32),
        tet_edges=rng.integers(0, 9, 14).astype(np.uint32), eigenvalues=np.sort(rng.uniform(1e5, 1e10, n_eigen)), summary_shapes=rng.standard_normal((n_points, n_eigen, 3)).astype(f32),
        material=np.array([2700.0, 7.2e10, 0.19, 6.0, 1e-7]), min_freq=f32(20), max_freq=f32(16000), num_modes=30, tet_hash=int(rng.integers(0, 2**63)), solved_vertices=rng.integers(0, 1000, n_points + 2).astype(np.uint32))


def _flat(m):
    keep = {k: np.ascontiguousarray(v) for k, v in m.items() if isinstance(v, np.ndarray)}
    f = RefModalFlat()
    f.n_modes, f.n_points, f.n_eigen = len(m["freqs"]), len(m["positions"]), len(m["eigenvalues"])
    f.n_vertices, f.n_indices, f.n_tet_positions, f.n_tet_edges, f.n_solved_vertices = len(m["vertices"]), len(m["indices"]), len(m["tet_positions"]), len(m["tet_edges"]), len(m["solved_vertices"])
    for k in ("freqs", "t60s", "shapes", "positions", "vertices", "indices", "tet_positions", "tet_edges", "eigenvalues", "summary_shapes", "solved_vertices"):
        setattr(f, k, keep[k].ctypes.data)
    f.original_fundamental, f.mass, f.min_freq, f.max_freq, f.num_modes, f.tet_hash = float(m["original_fundamental"]), m["mass"], float(m["min_freq"]), float(m["max_freq"]), m["num_modes"], m["tet_hash"]
    f.baked_scale, f.com, f.inertia = (C.c_float * 3)(*m["baked_scale"]), (C.c_float * 3)(*m["com"]), (C.c_float * 3)(*m["inertia"])
    f.quat_wxyz, f.material = (C.c_float * 4)(*m["quat_wxyz"]), (C.c_double * 5)(*m["material"])
    return f, keep


def _lib():
    L = C.CDLL(REF_SO)
    L.ref_modal_file_serialize.restype = C.c_uint64
    L.ref_modal_file_serialize.argtypes = [C.POINTER(RefModalFlat), C.c_void_p, C.c_uint64]
    L.ref_modal_file_roundtrip_equal.restype = C.c_int
    L.ref_modal_file_roundtrip_equal.argtypes = [C.POINTER(RefModalFlat), C.c_void_p, C.c_uint64]
    return L


def serialize(m) -> bytes:
    """Serialize(ModalModelData) of the reference (ModalModelFile.cpp:15-22)."""
    f, _keep = _flat(m)
    L = _lib()
    n = L.ref_modal_file_serialize(C.byref(f), None, 0)
    buf = (C.c_uint8 * n)()
    assert L.ref_modal_file_serialize(C.byref(f), buf, n) == n
    return bytes(buf)


def parses_to(m, data: bytes) -> int:
    """1 when the reference's archive decodes `data` to exactly `m`, 0 when to something else, -1 on a decode failure."""
    f, _keep = _flat(m)
    buf = np.frombuffer(data, np.uint8)
    return _lib().ref_modal_file_roundtrip_equal(C.byref(f), buf.ctypes.data, len(buf))
