"""Python face of the analysis C ABI, shaped like the reference's API (modal::mesh2modes / PostprocessModes / RescaleModes,
src/audio/mesh2modes.h:77-88; the inner operators of src/audio/CholeskyShiftInvert.h:18-23) so the parity tests read like
the reference's own tests (tests/ModalSolverTest.cpp, tests/ModalSolverBench.cpp). Nothing here computes: every number comes
from libme_modal.so."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from ._lib import (MeSymbolicInfo, ME_CANCELLED, ME_NO_MODES, ME_NOT_CONVERGED, ME_OK, MeError, MeFactorInfo, MeFemInfo, MeJobMonitor, MeMassProperties, MeMaterial, MeSolveProfile,
                   MeSolverConfig, check, lib, struct_dict)

# materials::acoustic (src/audio/AcousticMaterial.h:33-40): density, Young, Poisson, alpha, beta
MATERIALS = {
    "Ceramic": (2700, 7.2e10, 0.19, 6, 1e-7), "Glass": (2600, 6.2e10, 0.20, 1, 1e-7), "Wood": (750, 1.1e10, 0.25, 60, 2e-6), "Plastic": (1070, 1.4e9, 0.35, 30, 1e-6),
    "Iron": (8000, 2.1e11, 0.28, 5, 1e-7), "Polycarbonate": (1190, 2.4e9, 0.37, 0.5, 4e-7), "Steel": (7850, 2.0e11, 0.29, 5, 3e-8),
}


def material(m):
    if isinstance(m, MeMaterial):
        return m
    if isinstance(m, str):
        m = MATERIALS[m]
    if hasattr(m, "density"):
        m = (m.density, m.young, m.poisson, m.alpha, m.beta)
    return MeMaterial(*[float(x) for x in m])


def solver_config(num_modes=30, num_fem_modes=None, min_mode_freq=20.0, max_mode_freq=16000.0, tolerance=1e-8, max_restarts=100, fundamental_freq=None, element_order=2, device=0):
    c = MeSolverConfig()
    lib().me_solver_config_default(C.byref(c))
    c.num_modes, c.num_fem_modes = num_modes, num_modes + 15 if num_fem_modes is None else num_fem_modes
    c.min_mode_freq, c.max_mode_freq, c.tolerance, c.max_restarts = min_mode_freq, max_mode_freq, tolerance, max_restarts
    c.has_fundamental_freq, c.fundamental_freq = (1, fundamental_freq) if fundamental_freq else (0, 0.0)
    c.element_order, c.device = element_order, device
    return c


@dataclass
class ModalResult:
    """modal::ModalResult (mesh2modes.h:52-62)."""

    status: int = ME_OK
    freqs: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float32))
    t60s: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float32))
    shapes: np.ndarray = field(default_factory=lambda: np.zeros((0, 0, 3), np.float32))      # [point][mode][3]
    positions: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.float32))
    original_fundamental: float = 0.0
    sample_point_of_excitation: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint32))
    eigenvalues: np.ndarray = field(default_factory=lambda: np.zeros(0))
    summary_shapes: np.ndarray = field(default_factory=lambda: np.zeros((0, 0, 3), np.float32))  # [point][eigenpair][3]
    mass_props: dict = field(default_factory=dict)
    profile: dict = field(default_factory=dict)
    basis: np.ndarray | None = None

    @property
    def empty(self):
        return len(self.freqs) == 0


def _read_result(h, status):
    L = lib()
    r = ModalResult(status=status)
    m, p, e = L.me_modal_result_mode_count(h), L.me_modal_result_point_count(h), L.me_modal_result_eigenpair_count(h)

    def arr(ptr, shape, dtype):
        n = int(np.prod(shape))
        return np.ctypeslib.as_array(ptr, (n,)).astype(dtype).reshape(shape).copy() if n else np.zeros(shape, dtype)

    r.freqs, r.t60s = arr(L.me_modal_result_freqs(h), (m,), np.float32), arr(L.me_modal_result_t60s(h), (m,), np.float32)
    r.shapes = arr(L.me_modal_result_shapes(h), (p, m, 3), np.float32) if m else np.zeros((p, 0, 3), np.float32)
    r.positions = arr(L.me_modal_result_positions(h), (p, 3), np.float32) if m else np.zeros((0, 3), np.float32)
    r.original_fundamental = float(L.me_modal_result_original_fundamental(h))
    count = C.c_uint32()
    sp = L.me_modal_result_sample_point_of_excitation(h, C.byref(count))
    r.sample_point_of_excitation = arr(sp, (count.value,), np.uint32)
    r.eigenvalues = arr(L.me_modal_result_eigenvalues(h), (e,), np.float64)
    r.summary_shapes = arr(L.me_modal_result_summary_shapes(h), (p, e, 3), np.float32) if e else np.zeros((p, 0, 3), np.float32)
    mp, prof = MeMassProperties(), MeSolveProfile()
    check(L.me_modal_result_mass_properties(h, C.byref(mp)))
    check(L.me_modal_result_profile(h, C.byref(prof)))
    r.mass_props, r.profile = struct_dict(mp), struct_dict(prof)
    rows, cols = C.c_uint32(), C.c_uint32()
    b = L.me_modal_result_basis(h, C.byref(rows), C.byref(cols))
    if b and rows.value:
        r.basis = np.ctypeslib.as_array(b, (rows.value * cols.value,)).reshape(cols.value, rows.value).T.copy()
    return r


def mesh2modes(points, tets, mat, excite_positions, baked_scale=(1.0, 1.0, 1.0), config=None, keep_basis=False, monitor=None, seed_basis=None):
    """modal::mesh2modes (mesh2modes.h:77). seed_basis (n x cols, a prior result's `basis`) selects the warm re-solve
    (SolveReuse::SeedBasis). Like the reference, a cancelled / non-converged / no-mode solve returns an EMPTY
    result (status says which); a failed factorisation raises (the reference throws std::runtime_error)."""
    h, status = solve_handle(points, tets, mat, excite_positions, baked_scale, config, keep_basis, monitor, seed_basis)
    try:
        return _read_result(h, status)
    finally:
        lib().me_modal_result_free(h)


def solve_handle(points, tets, mat, excite_positions, baked_scale=(1.0, 1.0, 1.0), config=None, keep_basis=False, monitor=None, seed_basis=None):
    """me_modal_solve, returning the library-owned MeModalResult handle and the status (interchange.py keeps the handle)."""
    L = lib()
    pts = np.ascontiguousarray(points, np.float64)
    tt = np.ascontiguousarray(tets, np.uint32)
    ex = np.ascontiguousarray(excite_positions, np.float32).reshape(-1, 3)
    scale = (C.c_float * 3)(*baked_scale)
    cfg = config or solver_config()
    m = material(mat)
    h = C.c_void_p()
    seed, seed_rows, seed_cols = None, 0, 0
    if seed_basis is not None:
        seed_arr = np.asfortranarray(seed_basis, np.float32)  # column-major n x cols
        seed, (seed_rows, seed_cols) = seed_arr.ctypes.data, seed_arr.shape
    status = L.me_modal_solve(pts.ctypes.data, len(pts), tt.ctypes.data, len(tt), C.byref(m), ex.ctypes.data, len(ex), scale, C.byref(cfg), seed, seed_rows, seed_cols, int(keep_basis),
                              C.byref(monitor) if monitor is not None else None, C.byref(h))
    if status not in (ME_OK, ME_CANCELLED, ME_NOT_CONVERGED, ME_NO_MODES):
        raise MeError(status, L.me_last_error().decode())
    return h, status


def postprocess_modes(eigenvalues, shapes, shape_scale, mat, config, positions):
    """modal::PostprocessModes (mesh2modes.h:82). shapes: [point][eigenpair][3]."""
    L = lib()
    ev = np.ascontiguousarray(eigenvalues, np.float64)
    sh = np.ascontiguousarray(shapes, np.float32)
    pos = np.ascontiguousarray(positions, np.float32)
    m = material(mat)
    h = C.c_void_p()
    check(L.me_postprocess_modes(ev.ctypes.data, len(ev), sh.ctypes.data, sh.shape[0], shape_scale, C.byref(m), C.byref(config), pos.ctypes.data, C.byref(h)))
    try:
        return _read_result(h, ME_OK)
    finally:
        L.me_modal_result_free(h)


class FemSystem:
    """The assembled pencil (K, M) on the device: FilterDegenerate + BuildQuadMesh + AssembleQuadratic (mesh2modes.cpp:42-60,246-327)."""

    def __init__(self, points, tets, mat, element_order=2, device=0):
        pts = np.ascontiguousarray(points, np.float64)
        tt = np.ascontiguousarray(tets, np.uint32)
        m = material(mat)
        self._h = C.c_void_p()
        check(lib().me_fem_assemble(pts.ctypes.data, len(pts), tt.ctypes.data, len(tt), C.byref(m), element_order, device, C.byref(self._h)))
        info = MeFemInfo()
        check(lib().me_fem_info(self._h, C.byref(info)))
        self.info = struct_dict(info)

    def close(self):
        if getattr(self, "_h", None):
            lib().me_fem_free(self._h)
            self._h = None

    __del__ = close

    def element_nodes(self):
        out = np.zeros((self.info["tets_kept"], self.info["nodes_per_element"]), np.uint32)
        check(lib().me_fem_get_element_nodes(self._h, out.ctypes.data))
        return out

    def csc(self, which):
        """Eigen-layout lower-triangular CSC of K ("K") or M ("M"): (colptr, rowidx, values)."""
        w = {"K": 0, "M": 1}[which]
        nnz = self.info["nnz_stiffness" if w == 0 else "nnz_mass"]
        colptr, rowidx, values = np.zeros(self.info["dofs"] + 1, np.uint64), np.zeros(nnz, np.uint32), np.zeros(nnz, np.float64)
        check(lib().me_fem_get_csc(self._h, w, colptr.ctypes.data, rowidx.ctypes.data, values.ctypes.data))
        return colptr, rowidx, values

    def colour_elements(self):
        out, n = np.zeros(self.info["tets_kept"], np.uint32), C.c_uint32()
        check(lib().me_fem_colour_elements(self._h, out.ctypes.data, C.byref(n)))
        return out, n.value

    def spmv(self, which, x, repeats=1):
        x = np.ascontiguousarray(x, np.float64)
        y, ms = np.zeros_like(x), C.c_float()
        check(lib().me_fem_spmv(self._h, {"K": 0, "M": 1}[which], x.ctypes.data, y.ctypes.data, repeats, C.byref(ms)))
        self.last_spmv_ms = ms.value
        return y


class Factor:
    """CholeskyShiftInvert (src/audio/CholeskyShiftInvert.h:11-30): y = (K - sigma M)^-1 x on the device."""

    def __init__(self, fem, sigma):
        self.fem = fem
        self._h = C.c_void_p()
        check(lib().me_factor_create(fem._h, float(sigma), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            lib().me_factor_free(self._h)
            self._h = None

    __del__ = close

    def solve(self, b):
        b = np.asfortranarray(b, np.float64)
        width = 1 if b.ndim == 1 else b.shape[1]
        x = np.zeros_like(b, order="F")
        check(lib().me_factor_solve(self._h, b.ctypes.data, x.ctypes.data, width))
        return x

    @property
    def info(self):
        i = MeFactorInfo()
        check(lib().me_factor_info(self._h, C.byref(i)))
        return struct_dict(i)


def measure_fp64_rate(device=0, mode=1, iters=5):
    out = C.c_double()
    check(lib().me_measure_fp64_rate(device, mode, iters, C.byref(out)))
    return out.value


def symbolic_analyse(rowptr, col, xyz):
    """The host-side ordering + supernodal structure of the sparse Cholesky (no device needed): (perm, info dict)."""
    rowptr, col, xyz = np.ascontiguousarray(rowptr, np.uint32), np.ascontiguousarray(col, np.uint32), np.ascontiguousarray(xyz, np.float32)
    n = len(rowptr) - 1
    perm, info = np.zeros(n, np.uint32), MeSymbolicInfo()
    check(lib().me_symbolic_analyse(n, rowptr.ctypes.data, col.ctypes.data, xyz.ctypes.data, perm.ctypes.data, C.byref(info)))
    return perm, struct_dict(info)


def effective_modal_material(props, solved, solve_mass, body_mass=0.0) -> MeMaterial:
    """EffectiveModalMaterial (AudioSystem.cpp:595-601); body_mass <= 0: not an authoritative dynamic rigid body."""
    p, s, out = material(props), material(solved), MeMaterial()
    check(lib().me_effective_modal_material(C.byref(p), C.byref(s), solve_mass, body_mass, C.byref(out)))
    return out


def pinned_fundamental(freqs, original_fundamental):
    """RescaledModes' rule (AudioSystem.cpp:612-616): the fundamental to keep pinned through a rescale, or None."""
    f, out = np.ascontiguousarray(freqs, np.float32), C.c_float()
    return out.value if lib().me_pinned_fundamental(f.ctypes.data, len(f), original_fundamental, C.byref(out)) else None


def impact_spectrum(frames, sample_rate=48000):
    """ComputeFft (AudioSystem.cpp:553-558) -> complex64 spectrum of the windowed segment (n_real/2 + 1 bins), n_real."""
    f, n = np.ascontiguousarray(frames, np.float32), C.c_uint64()
    check(lib().me_impact_spectrum(f.ctypes.data, len(f), int(sample_rate), None, C.byref(n)))
    out = np.zeros((n.value // 2 + 1, 2), np.float32)
    check(lib().me_impact_spectrum(f.ctypes.data, len(f), int(sample_rate), out.ctypes.data, C.byref(n)))
    return out.view(np.complex64).reshape(-1), n.value


def estimate_fundamental_from_spectrum(spectrum, n_real, sample_rate=48000):
    """EstimateFundamentalFrequency (AudioSystem.cpp:522-550); None where the reference returns nullopt."""
    s, hz = np.ascontiguousarray(spectrum, np.complex64), C.c_float()
    return hz.value if lib().me_estimate_fundamental_from_spectrum(s.ctypes.data, int(n_real), int(sample_rate), C.byref(hz)) else None


def estimate_fundamental(frames, sample_rate=48000):
    """The fundamental LaunchModalSolve matches the model to (AudioSystem.cpp:821-829), or None."""
    f, hz = np.ascontiguousarray(frames, np.float32), C.c_float()
    return hz.value if lib().me_estimate_fundamental(f.ctypes.data, len(f), int(sample_rate), C.byref(hz)) else None
