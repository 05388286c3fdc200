"""Strike front-end over the C ABI (include/me_modal.h, "Strike front-end"): Hertz contact time, reduced contact mass,
recoil click filter and the ModalEvent of a strike. Mirrors src/audio/ContactModel.h and TriggerModalStrike
(src/audio/AudioSystem.cpp:400-465). Host-only."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import MeContactDynamics, MeImpactor, MeMassProperties, MeMaterial, MeModalEvent, MeStrike, MeStriker, check, lib

STEEL = (7850.0, 2.0e11, 0.29, 5.0, 3.0e-8)  # materials::acoustic::Steel (AcousticMaterial.h:38)
CONSTANTS = {"inv_effective_modulus": 0, "combined_curvature": 1, "stiffness": 2, "patch_radius": 3, "static_penetration": 4, "saturation_penetration": 5, "punch_stiffness": 6}


def material(m) -> MeMaterial:
    return m if isinstance(m, MeMaterial) else MeMaterial(*[float(x) for x in m])


def striker(mat=STEEL, tip_radius=0.01, length=0.19) -> MeStriker:
    """Striker{} (ContactModel.h:36-40)."""
    return MeStriker(material(mat), tip_radius, length)


def striker_mass(s: MeStriker) -> float:
    return lib().me_striker_mass(C.byref(s))


def striker_impactor(s: MeStriker) -> MeImpactor:
    out = MeImpactor()
    check(lib().me_striker_impactor(C.byref(s), C.byref(out)))
    return out


def impactor(mat, curvature, inv_mass) -> MeImpactor:
    return MeImpactor(material(mat), curvature, inv_mass)


def inverse_inertia_tensor(mass, inertia_diagonal, quat_wxyz) -> np.ndarray:
    mp = MeMassProperties(mass, (C.c_float * 3)(0, 0, 0), (C.c_float * 3)(*inertia_diagonal), (C.c_float * 4)(*quat_wxyz))
    out = np.zeros(9, np.float32)
    check(lib().me_inverse_inertia_tensor(C.byref(mp), out.ctypes.data))
    return out


class ContactDynamics:
    """ContactDynamics (ContactModel.h:28-32); keeps the arm array alive for the struct that borrows it."""

    def __init__(self, mass, inverse_inertia, arms):
        self.arms = np.ascontiguousarray(arms, np.float32).reshape(-1, 3)
        self.c = MeContactDynamics(mass, (C.c_float * 9)(*np.asarray(inverse_inertia, np.float32).reshape(-1)), self.arms.ctypes.data, len(self.arms))


def contact_dynamics(mass_props: dict, positions, baked_scale=(1.0, 1.0, 1.0), mass_scale=1.0) -> ContactDynamics:
    """UpdateContactDynamics (ContactDynamics.cpp:19-46): from a solve's mass properties (ModalResult.mass_props) and sample points."""
    mp = MeMassProperties(mass_props["mass"], (C.c_float * 3)(*mass_props["center_of_mass"]), (C.c_float * 3)(*mass_props["inertia_diagonal"]), (C.c_float * 4)(*mass_props["inertia_orientation"]))
    pos = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
    mass, inverse, arms = C.c_double(), np.zeros(9, np.float32), np.zeros_like(pos)
    check(lib().me_contact_dynamics(C.byref(mp), mass_scale, pos.ctypes.data, len(pos), (C.c_float * 3)(*baked_scale), C.byref(mass), inverse.ctypes.data, arms.ctypes.data))
    return ContactDynamics(mass.value, inverse, arms)


def _dir(d):
    return np.ascontiguousarray(d, np.float32)


def reduced_contact_mass(d: ContactDynamics, i, direction, imp: MeImpactor) -> float:
    direction = _dir(direction)
    return lib().me_reduced_contact_mass(C.byref(d.c), i, direction.ctypes.data, C.byref(imp))


def estimate_contact_time(d: ContactDynamics, i, direction, contact_speed, object_material, object_curvature, nominal_area, imp: MeImpactor, scale_ratio=1.0, combined_roughness=0.0) -> float:
    direction = _dir(direction)
    m = material(object_material)
    return lib().me_estimate_contact_time(C.byref(d.c), i, direction.ctypes.data, contact_speed, C.byref(m), object_curvature, nominal_area, C.byref(imp), scale_ratio, combined_roughness)


def contact_constant(name, a=None, b=None, x=0.0, y=0.0, z=0.0) -> float:
    ma, mb = (material(a), material(b)) if a is not None else (None, None)
    return lib().me_contact_constant(CONSTANTS[name], C.byref(ma) if ma else None, C.byref(mb) if mb else None, x, y, z)


def recoil_click_filter(radius, volume, mass, sample_rate):
    out = np.zeros(3, np.float32)
    lib().me_recoil_click_filter(radius, volume, mass, sample_rate, out.ctypes.data)
    return tuple(float(v) for v in out)


def make_strike_event(object_slot, excitable_index, force, contact_speed, direction, *, dynamics: ContactDynamics | None = None, elastic=None, imp: MeImpactor | None = None, is_collision=False,
                      resultant_index=0, curvature=0.0, nominal_area=0.0, scale_ratio=1.0, roughness=0.0, displaced_volume=0.0, radiant_radius=0.0, sample_rate=48000.0) -> MeModalEvent:
    """TriggerModalStrike (AudioSystem.cpp:400-465) after its scene lookups; a mallet strike defaults to Striker{}."""
    el = material(elastic) if elastic is not None else None
    s = MeStrike(object_slot, excitable_index, force, contact_speed, (C.c_float * 3)(*direction), int(is_collision), resultant_index, C.pointer(dynamics.c) if dynamics else None,
                 C.pointer(el) if el is not None else None, imp if imp is not None else striker_impactor(striker()), curvature, nominal_area, scale_ratio, roughness, displaced_volume, radiant_radius,
                 sample_rate)
    out = MeModalEvent()
    check(lib().me_make_strike_event(C.byref(s), C.byref(out)))
    return out


def tilt_along_normal(normal, joystick=(0.0, 0.0)) -> np.ndarray:
    """TiltAlongNormal (AudioSystem.cpp:359-371): a manual strike's direction."""
    n, j, out = np.ascontiguousarray(normal, np.float32), np.ascontiguousarray(joystick, np.float32), np.zeros(3, np.float32)
    lib().me_tilt_along_normal(n.ctypes.data, j.ctypes.data, out.ctypes.data)
    return out


def sphere_equivalent_curvature(density, inv_mass) -> float:
    """SphereEquivalentCurvature (AudioSystem.cpp:379-380)."""
    return lib().me_sphere_equivalent_curvature(density, inv_mass)
