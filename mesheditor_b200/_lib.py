"""Loads libme_modal.so (the C ABI declared in include/me_modal.h) through ctypes.

There is no CPU fallback and no alternative backend: if the library is missing or there is no CUDA device,
calls fail loudly (ImportError here, MeError with ME_CUDA_ERROR from the compute entry points).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libme_modal.so")

ME_OK, ME_BAD_ARG, ME_CUDA_ERROR, ME_CANCELLED, ME_NOT_CONVERGED, ME_NO_MODES, ME_FACTOR_FAILED, ME_QUEUE_FULL, ME_OUT_OF_MEMORY = range(9)
STATUS_NAMES = ["ME_OK", "ME_BAD_ARG", "ME_CUDA_ERROR", "ME_CANCELLED", "ME_NOT_CONVERGED", "ME_NO_MODES", "ME_FACTOR_FAILED", "ME_QUEUE_FULL", "ME_OUT_OF_MEMORY"]


class MeError(RuntimeError):
    def __init__(self, status, text):
        super().__init__(f"{STATUS_NAMES[status] if 0 <= status < len(STATUS_NAMES) else status}: {text}")
        self.status = status


class MeModalEvent(C.Structure):
    """include/me_modal.h MeModalEvent == the reference's ModalEvent (ModalAudio.h:28-37)."""

    _fields_ = [("kind", C.c_uint32), ("object", C.c_uint32), ("ex_pos", C.c_uint32)] + [
        (n, C.c_float) for n in ("jx", "jy", "jz", "pulse_step", "pulse_gamma", "accel_amp", "click_b0", "click_a1", "click_a2")
    ]


class MeRenderStats(C.Structure):
    _fields_ = [
        ("kernel_launches", C.c_uint32), ("resonator_kernel_ms", C.c_float), ("total_device_ms", C.c_float),
        ("mode_samples", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("time_segments", C.c_uint32), ("scan_fallbacks", C.c_uint32),
        ("tensor_windows", C.c_uint32), ("partial_rows", C.c_uint32), ("walk_kernel_ms", C.c_float), ("tensor_mix_kernel_ms", C.c_float), ("host_plan_ms", C.c_float), ("pulse_kernels_ms", C.c_float),
    ]


class MeMaterial(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("density", "young_modulus", "poisson_ratio", "alpha", "beta")]


class MeModalFileExtras(C.Structure):
    _fields_ = [("vertices", C.c_void_p), ("n_vertices", C.c_uint32), ("indices", C.c_void_p), ("n_indices", C.c_uint32), ("baked_scale", C.c_float * 3),
                ("tet_positions_xyz", C.c_void_p), ("n_tet_positions", C.c_uint32), ("tet_edge_indices", C.c_void_p), ("n_tet_edge_indices", C.c_uint32), ("solved_material", MeMaterial),
                ("solved_min_mode_freq", C.c_float), ("solved_max_mode_freq", C.c_float), ("solved_num_modes", C.c_uint32), ("tet_inputs_hash", C.c_uint64),
                ("solved_vertices", C.c_void_p), ("n_solved_vertices", C.c_uint32)]


class MeStriker(C.Structure):
    _fields_ = [("material", MeMaterial), ("tip_radius", C.c_float), ("length", C.c_float)]


class MeImpactor(C.Structure):
    _fields_ = [("material", MeMaterial), ("curvature", C.c_double), ("inv_mass", C.c_double)]


class MeContactDynamics(C.Structure):
    _fields_ = [("mass", C.c_double), ("inverse_inertia", C.c_float * 9), ("contact_arm_xyz", C.c_void_p), ("arm_count", C.c_uint32)]


class MeStrike(C.Structure):
    _fields_ = [("object", C.c_uint32), ("excitable_index", C.c_uint32), ("force", C.c_float), ("contact_speed", C.c_float), ("direction", C.c_float * 3), ("is_collision", C.c_int32),
                ("resultant_index", C.c_uint32), ("dynamics", C.POINTER(MeContactDynamics)), ("elastic", C.POINTER(MeMaterial)), ("impactor", MeImpactor), ("curvature", C.c_double),
                ("nominal_area", C.c_double), ("scale_ratio", C.c_double), ("roughness", C.c_double), ("displaced_volume", C.c_double), ("radiant_radius", C.c_float), ("sample_rate", C.c_float)]


class MeSolverConfig(C.Structure):
    _fields_ = [("min_mode_freq", C.c_float), ("max_mode_freq", C.c_float), ("num_modes", C.c_uint32), ("num_fem_modes", C.c_uint32), ("tolerance", C.c_double),
                ("warm_tolerance", C.c_double), ("max_restarts", C.c_uint32), ("has_fundamental_freq", C.c_int32), ("fundamental_freq", C.c_float),
                ("element_order", C.c_uint32), ("device", C.c_int32)]


class MeJobMonitor(C.Structure):
    _fields_ = [("progress", C.c_float), ("cancelled", C.c_int32)]


class MeSolveProfile(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("mass_props", "quad_mesh", "assemble", "sample_excite", "factorize", "iterate", "op_solve", "extract")] + [
        (n, C.c_uint32) for n in ("dofs", "stiffness_nonzeros", "op_applications", "restarts")] + [("analyse", C.c_double), ("factor_flops", C.c_double),
        ("factor_nonzeros", C.c_uint64), ("assemble_kernel_ms", C.c_float), ("factor_device_ms", C.c_float)] + [
        (n, C.c_uint32) for n in ("supernodes", "levels", "kernel_launches", "tets_kept")]


class MeRetune(C.Structure):
    """include/me_modal.h MeRetune: RetuneModalObject's inputs after its scene lookups (AudioSystem.cpp:263-311)."""

    _fields_ = [("scale", C.c_float), ("fundamental", C.c_float), ("t60_scale", C.c_float), ("has_alpha", C.c_int), ("alpha", C.c_double), ("modal_level", C.c_float), ("gain", C.c_float)]


class MeMassProperties(C.Structure):
    _fields_ = [("mass", C.c_double), ("center_of_mass", C.c_float * 3), ("inertia_diagonal", C.c_float * 3), ("inertia_orientation", C.c_float * 4)]


class MeFemInfo(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("tets_kept", "node_count", "dofs", "nodes_per_element")] + [
        (n, C.c_uint64) for n in ("nnz_stiffness", "nnz_mass", "node_blocks_lower", "node_blocks_full")] + [("assemble_kernel_ms", C.c_float), ("kernel_launches", C.c_uint32)]


class MeFactorInfo(C.Structure):
    _fields_ = [("analyse_seconds", C.c_double), ("factor_flops", C.c_double), ("factor_device_ms", C.c_float), ("last_solve_device_ms", C.c_float),
                ("factor_nonzeros", C.c_uint64)] + [(n, C.c_uint32) for n in ("supernodes", "levels", "dofs", "kernel_launches")]


class MeSymbolicInfo(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("supernodes", "levels", "max_panel_columns", "max_panel_rows")] + [
        (n, C.c_uint64) for n in ("factor_nonzeros", "update_tiles", "panel_tiles", "violations")] + [(n, C.c_double) for n in ("factor_flops", "ordering_seconds", "structure_seconds")]


def struct_dict(s):
    out = {}
    for name, _ in s._fields_:
        v = getattr(s, name)
        out[name] = list(v) if hasattr(v, "__len__") else v
    return out


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m mesheditor_b200.build` (nvcc, sm_100a). There is no fallback path.")
    L = C.CDLL(LIB_PATH)
    vp, u32, u64, f32, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_float, C.c_int
    L.me_last_error.restype = C.c_char_p
    L.me_build_info.restype = C.c_char_p
    L.me_device_count.restype = i32
    sig = {
        "me_bank_create": [f32, i32, C.POINTER(vp)],
        "me_bank_add_object": [vp, u32, u32, vp, vp, vp, u32, C.POINTER(u32)],
        "me_bank_tune_object": [vp, u32, vp, vp, u32, f32],
        "me_bank_set_object_shapes": [vp, u32, u32, u32, vp],
        "me_bank_set_gain": [vp, u32, f32, f32],
        "me_bank_set_click_gain": [vp, f32],
        "me_bank_set_max_impacts": [vp, u32],
        "me_bank_set_time_segments": [vp, u32],
        "me_bank_set_render_path": [vp, u32],
        "me_bank_install": [vp],
        "me_bank_enqueue": [vp, C.POINTER(MeModalEvent)],
        "me_bank_render": [vp, vp, u32],
        "me_bank_render_offline": [vp, vp, vp, u32, u64, u32, vp],
        "me_bank_render_offline_device": [vp, vp, vp, u32, u64, u32, vp, vp],
        "me_bank_get_mode_column": [vp, i32, vp],
        "me_bank_get_object_layout": [vp, u32, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32), C.POINTER(f32)],
        "me_bank_get_object_status": [vp, u32, C.POINTER(u32), C.POINTER(u32)],
        "me_bank_last_render_stats": [vp, C.POINTER(MeRenderStats)],
        "me_measure_fp32_fma_rate": [i32, i32, i32, C.POINTER(C.c_double)],
        "me_measure_fp64_rate": [i32, i32, i32, C.POINTER(C.c_double)],
        "me_debug_tensor_mix": [i32, vp, vp, u32, u32, u32, u32, u32, u32, vp, C.POINTER(f32)],
        "me_modal_solve": [vp, u32, vp, u32, C.POINTER(MeMaterial), vp, u32, vp, C.POINTER(MeSolverConfig), vp, u32, u32, i32, C.POINTER(MeJobMonitor), C.POINTER(vp)],
        "me_modal_result_mass_properties": [vp, C.POINTER(MeMassProperties)],
        "me_modal_result_profile": [vp, C.POINTER(MeSolveProfile)],
        "me_postprocess_modes": [vp, u32, vp, u32, f32, C.POINTER(MeMaterial), C.POINTER(MeSolverConfig), vp, C.POINTER(vp)],
        "me_deal_objects": [vp, u32, u32, vp, vp],
        "me_rescale_modes": [vp, C.POINTER(MeMaterial), C.POINTER(MeMaterial), C.POINTER(MeSolverConfig), C.POINTER(vp)],
        "me_fem_assemble": [vp, u32, vp, u32, C.POINTER(MeMaterial), u32, i32, C.POINTER(vp)],
        "me_fem_info": [vp, C.POINTER(MeFemInfo)],
        "me_fem_get_element_nodes": [vp, vp],
        "me_fem_get_csc": [vp, i32, vp, vp, vp],
        "me_fem_colour_elements": [vp, vp, C.POINTER(u32)],
        "me_fem_spmv": [vp, i32, vp, vp, u32, C.POINTER(f32)],
        "me_factor_create": [vp, C.c_double, C.POINTER(vp)],
        "me_factor_solve": [vp, vp, vp, u32],
        "me_factor_info": [vp, C.POINTER(MeFactorInfo)],
        "me_symbolic_analyse": [u32, vp, vp, vp, vp, C.POINTER(MeSymbolicInfo)],
    }
    sig.update({
        "me_modal_file_serialize": [vp, C.POINTER(MeModalFileExtras), C.POINTER(vp), C.POINTER(u64)],
        "me_modal_file_parse": [vp, u64, C.POINTER(vp), C.POINTER(vp)],
        "me_modal_file_extras": [vp, C.POINTER(MeModalFileExtras)],
        "me_modal_solve_json": [vp, vp, u32, C.POINTER(vp)],
        "me_striker_impactor": [C.POINTER(MeStriker), C.POINTER(MeImpactor)],
        "me_inverse_inertia_tensor": [C.POINTER(MeMassProperties), vp],
        "me_make_strike_event": [C.POINTER(MeStrike), C.POINTER(MeModalEvent)],
        "me_contact_dynamics": [C.POINTER(MeMassProperties), C.c_double, vp, u32, vp, C.POINTER(C.c_double), vp, vp],
        "me_effective_modal_material": [C.POINTER(MeMaterial), C.POINTER(MeMaterial), C.c_double, C.c_double, C.POINTER(MeMaterial)],
        "me_pinned_fundamental": [vp, u32, f32, C.POINTER(f32)],
        "me_wav_encode": [vp, u64, u32, f32, C.POINTER(vp), C.POINTER(u64)],
        "me_wav_decode": [vp, u64, C.POINTER(vp), C.POINTER(u64), C.POINTER(u32)],
        "me_impact_spectrum": [vp, u64, u32, vp, C.POINTER(u64)],
        "me_monitor_frames": [vp, u64, f32, C.POINTER(f32)],
        "me_retune_modes": [vp, vp, u32, C.POINTER(MeRetune), vp, vp],
        "me_bank_retune_object": [vp, u32, vp, vp, u32, C.POINTER(MeRetune)],
        "me_desired_solve_vertices": [u32, u32, C.POINTER(vp), C.POINTER(u32)],
        "me_sample_surface_triangles": [vp, u32, u32, vp, u32, C.POINTER(vp), C.POINTER(u32)],
        "me_compact_excitation_vertices": [vp, u32, vp, u32, C.POINTER(vp), C.POINTER(u32)],
        "me_relabel_sample_triangles": [vp, u32, vp, u32, C.POINTER(vp), C.POINTER(u32)],
        "me_build_tet_mesh_data": [vp, u32, vp, u32, vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(u32)],
    })
    for name, args in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = i32
    f64 = C.c_double
    for name, args in {
        "me_striker_mass": [C.POINTER(MeStriker)],
        "me_reduced_contact_mass": [C.POINTER(MeContactDynamics), u32, vp, C.POINTER(MeImpactor)],
        "me_estimate_contact_time": [C.POINTER(MeContactDynamics), u32, vp, f64, C.POINTER(MeMaterial), f64, f64, C.POINTER(MeImpactor), f64, f64],
        "me_contact_constant": [i32, C.POINTER(MeMaterial), C.POINTER(MeMaterial), f64, f64, f64],
        "me_sphere_equivalent_curvature": [f64, f64],
    }.items():
        getattr(L, name).argtypes = args
        getattr(L, name).restype = f64
    for name in ("me_modal_file_free", "me_bytes_free"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = None
    for name, args in {"me_modal_out_gain": [C.POINTER(MeRetune)], "me_uniform_scale_ratio": [vp, vp], "me_listener_gain": [f32]}.items():
        getattr(L, name).argtypes = args
        getattr(L, name).restype = f32
    L.me_estimate_fundamental_from_spectrum.argtypes, L.me_estimate_fundamental_from_spectrum.restype = [vp, u64, u32, C.POINTER(f32)], i32
    L.me_estimate_fundamental.argtypes, L.me_estimate_fundamental.restype = [vp, u64, u32, C.POINTER(f32)], i32
    L.me_tilt_along_normal.argtypes, L.me_tilt_along_normal.restype = [vp, vp, vp], None
    L.me_recoil_click_filter.argtypes = [f64, f64, f64, f64, vp]
    L.me_recoil_click_filter.restype = None
    for name in ("me_bank_free", "me_modal_result_free", "me_fem_free", "me_factor_free"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = None
    L.me_solver_config_default.argtypes = [C.POINTER(MeSolverConfig)]
    L.me_solver_config_default.restype = None
    for name in ("me_modal_result_mode_count", "me_modal_result_point_count", "me_modal_result_eigenpair_count"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = u32
    for name in ("me_modal_result_freqs", "me_modal_result_t60s", "me_modal_result_shapes", "me_modal_result_positions", "me_modal_result_summary_shapes"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = C.POINTER(f32)
    L.me_modal_result_eigenvalues.argtypes = [vp]
    L.me_modal_result_eigenvalues.restype = C.POINTER(C.c_double)
    L.me_modal_result_original_fundamental.argtypes = [vp]
    L.me_modal_result_original_fundamental.restype = f32
    L.me_modal_result_sample_point_of_excitation.argtypes = [vp, C.POINTER(u32)]
    L.me_modal_result_sample_point_of_excitation.restype = C.POINTER(u32)
    L.me_modal_result_basis.argtypes = [vp, C.POINTER(u32), C.POINTER(u32)]
    L.me_modal_result_basis.restype = C.POINTER(f32)
    for name in ("me_bank_object_count", "me_bank_mode_total", "me_bank_active_impacts"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = u32
    L.me_bank_events_dropped.argtypes = [vp]
    L.me_bank_events_dropped.restype = u64
    _lib = L
    return L


def check(status):
    if status != ME_OK:
        raise MeError(status, lib().me_last_error().decode())
