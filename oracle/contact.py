"""TEST INFRASTRUCTURE ONLY (oracle). ctypes front-end of the UNMODIFIED reference ContactModel.cpp / RecoilClickFilter
compiled into oracle/_ref/libme_ref_audio.so (oracle/Makefile, oracle/ref_audio_driver.cpp), plus a Python restatement
of the arithmetic of TriggerModalStrike (src/audio/AudioSystem.cpp:400-465) composed from those reference functions.
Parity pinned: these ARE the reference's functions; tests/golden/strike/contact.npz holds their outputs for boxes without _ref."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

REF_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libme_ref_audio.so")
F32P = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
F64P = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_L = None


def have_ref():
    return os.path.exists(REF_SO)


def ref():
    global _L
    if _L is None:
        L = C.CDLL(REF_SO)
        d, f, u = C.c_double, C.c_float, C.c_uint32
        L.ref_striker_mass.restype = d
        L.ref_striker_mass.argtypes = [F64P, f, f]
        L.ref_striker_impactor.argtypes = [F64P, f, f, F64P]
        L.ref_inverse_inertia.argtypes = [d, F32P, F32P, F32P]
        L.ref_reduced_contact_mass.restype = d
        L.ref_reduced_contact_mass.argtypes = [d, F32P, F32P, u, u, F32P, F64P, d, d]
        L.ref_estimate_contact_time.restype = d
        L.ref_estimate_contact_time.argtypes = [d, F32P, F32P, u, u, F32P, d, F64P, d, d, F64P, d, d, d, d]
        L.ref_inv_effective_modulus.restype = d
        L.ref_inv_effective_modulus.argtypes = [F64P, F64P]
        L.ref_contact_scalar.restype = d
        L.ref_contact_scalar.argtypes = [C.c_int, d, d, d]
        L.ref_click_filter.argtypes = [d, d, d, d, F32P]
        _L = L
    return _L


def _m(mat):
    return np.ascontiguousarray(mat, np.float64)


def _f(a):
    return np.ascontiguousarray(a, np.float32)


def striker_mass(mat, tip_radius, length):
    return ref().ref_striker_mass(_m(mat), tip_radius, length)


def striker_impactor(mat, tip_radius, length):
    out = np.zeros(2)
    ref().ref_striker_impactor(_m(mat), tip_radius, length, out)
    return float(out[0]), float(out[1])


def inverse_inertia(mass, inertia_diagonal, quat_wxyz):
    out = np.zeros(9, np.float32)
    ref().ref_inverse_inertia(mass, _f(inertia_diagonal), _f(quat_wxyz), out)
    return out


def reduced_contact_mass(mass, inv_inertia, arms, i, direction, imp_mat, imp_curvature, imp_inv_mass):
    arms = _f(arms).reshape(-1, 3)
    return ref().ref_reduced_contact_mass(mass, _f(inv_inertia).reshape(-1), arms, len(arms), i, _f(direction), _m(imp_mat), imp_curvature, imp_inv_mass)


def estimate_contact_time(mass, inv_inertia, arms, i, direction, speed, obj_mat, obj_curvature, nominal_area, imp_mat, imp_curvature, imp_inv_mass, scale_ratio, roughness):
    arms = _f(arms).reshape(-1, 3)
    return ref().ref_estimate_contact_time(mass, _f(inv_inertia).reshape(-1), arms, len(arms), i, _f(direction), speed, _m(obj_mat), obj_curvature, nominal_area, _m(imp_mat), imp_curvature, imp_inv_mass,
                                           scale_ratio, roughness)


def inv_effective_modulus(a, b):
    return ref().ref_inv_effective_modulus(_m(a), _m(b))


SCALARS = {"combined_curvature": 0, "stiffness": 1, "patch_radius": 2, "static_penetration": 3, "saturation_penetration": 4, "punch_stiffness": 5}


def contact_scalar(name, x, y, z=0.0):
    return ref().ref_contact_scalar(SCALARS[name], x, y, z)


def click_filter(radius, volume, mass, sample_rate):
    out = np.zeros(3, np.float32)
    ref().ref_click_filter(radius, volume, mass, sample_rate, out)
    return out


def trigger_modal_strike(slot, excitable_index, force, contact_speed, direction, dyn, elastic, imp, *, is_collision, resultant_index, curvature, nominal_area, scale_ratio, roughness, displaced_volume,
                         radiant_radius, sample_rate):
    """The event TriggerModalStrike enqueues (AudioSystem.cpp:400-465), scene lookups replaced by arguments.
    dyn = (mass, inv_inertia[9], arms[n][3]) or None; imp = (material, curvature, inv_mass). Returns the 12 event fields."""
    direction = _f(direction)
    if is_collision:
        direction = (direction * np.float32(1.0 / np.sqrt(np.float32(np.dot(direction, direction))))).astype(np.float32)  # glm::normalize
    tau, click_amp, click = 1e-4, np.float32(0), np.zeros(3, np.float32)
    if dyn is not None and elastic is not None:
        mass, inv_inertia, arms = dyn
        tau = estimate_contact_time(mass, inv_inertia, arms, resultant_index if is_collision else excitable_index, direction, contact_speed, elastic, curvature, nominal_area if is_collision else 0.0,
                                    imp[0], imp[1], imp[2], scale_ratio, roughness)
        volume = displaced_volume
        radius = np.cbrt(3.0 * volume / (4.0 * np.pi)) if volume > 0 else float(np.float32(radiant_radius) * np.float32(scale_ratio))
        click = click_filter(radius, volume, mass, sample_rate)
        impulse = float(np.float32(force)) if is_collision else reduced_contact_mass(mass, inv_inertia, arms, excitable_index, direction, imp[0], imp[1], imp[2]) * abs(float(np.float32(contact_speed)))
        click_amp = np.float32(impulse * float(np.float32(sample_rate)))
    step = np.float32(1.0 / (tau * float(np.float32(sample_rate))))
    f = np.float32(force)
    return dict(kind=0, object=slot, ex_pos=excitable_index, jx=direction[0] * f, jy=direction[1] * f, jz=direction[2] * f, pulse_step=step, pulse_gamma=np.float32(2) * step, accel_amp=click_amp,
                click_b0=click[0], click_a1=click[1], click_a2=click[2])
