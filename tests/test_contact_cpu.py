"""Strike front-end through the C ABI against the UNMODIFIED reference ContactModel.cpp / RecoilClickFilter:
the committed golden vectors (tests/golden/strike/contact.npz, made by tests/golden/make_contact_golden.py from oracle/_ref) always,
and oracle/_ref live where it is built. Host-only: no CUDA device needed.
Tolerances: 1e-12 relative for the FP64 formulas, 2e-6 where glm float arithmetic sits in the path (compiler contraction)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_contact_golden import cases  # noqa: E402

from mesheditor_b200 import contact as mc  # noqa: E402
from oracle import contact as oc  # noqa: E402

pytestmark = pytest.mark.usefixtures("built_lib")  # builds libme_modal.so on demand (tests/conftest.py)

GOLDEN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "strike", "contact.npz"))
FIELDS = ("jx", "jy", "jz", "pulse_step", "pulse_gamma", "accel_amp", "click_b0", "click_a1", "click_a2")


def close(a, b, rel):
    a, b = np.atleast_1d(np.asarray(a, np.float64)), np.atleast_1d(np.asarray(b, np.float64))
    same = a == b  # covers equal infinities
    with np.errstate(invalid="ignore"):
        return bool(np.all(same | (np.abs(a - b) <= rel * np.maximum(np.abs(b), 1e-300))))


def ours(c):
    imp = mc.striker_impactor(mc.striker(c["imp_mat"], float(c["tip_radius"]), float(c["length"])))
    inv_inertia = mc.inverse_inertia_tensor(c["mass"], c["inertia"], c["quat"])
    dyn = mc.ContactDynamics(c["mass"], inv_inertia, c["arms"])
    volume = c["mass"] / c["obj_mat"][0] if c["mass"] > 0 else 0.0
    r = dict(
        striker_mass=mc.striker_mass(mc.striker(c["imp_mat"], float(c["tip_radius"]), float(c["length"]))), imp_curvature=imp.curvature, imp_inv_mass=imp.inv_mass, inv_inertia=inv_inertia,
        reduced_mass=mc.reduced_contact_mass(dyn, c["index"], c["direction"], imp),
        tau=mc.estimate_contact_time(dyn, c["index"], c["direction"], c["speed"], c["obj_mat"], c["curvature"], c["area"], imp, c["scale"], c["roughness"]),
        inv_modulus=mc.contact_constant("inv_effective_modulus", c["obj_mat"], c["imp_mat"]), click=np.array(mc.recoil_click_filter(c["radius"], volume, c["mass"], c["rate"])),
    )
    e, k = r["inv_modulus"], mc.contact_constant("combined_curvature", x=c["curvature"], y=imp.curvature)
    stiff = mc.contact_constant("stiffness", x=e, y=k)
    r["scalars"] = np.array([k, stiff, mc.contact_constant("patch_radius", x=c["force"] * 100.0, y=e, z=k), mc.contact_constant("static_penetration", x=c["force"] * 100.0, y=stiff),
                             mc.contact_constant("saturation_penetration", x=k, y=c["area"]), mc.contact_constant("punch_stiffness", x=e, y=c["area"])])
    for is_collision in (False, True):
        ev = mc.make_strike_event(3, min(c["index"], len(c["arms"]) - 1), float(c["force"]), c["speed"], c["direction"] if is_collision else c["direction"] / np.linalg.norm(c["direction"]),
                                  dynamics=dyn if c["mass"] > 0 else None, elastic=c["obj_mat"], imp=imp, is_collision=is_collision, resultant_index=c["index"] % len(c["arms"]), curvature=c["curvature"],
                                  nominal_area=c["area"], scale_ratio=c["scale"], roughness=c["roughness"], displaced_volume=volume if int(c["mass"] * 1e6) % 2 else 0.0, radiant_radius=c["radius"],
                                  sample_rate=c["rate"])
        assert (ev.kind, ev.object, ev.ex_pos) == (0, 3, min(c["index"], len(c["arms"]) - 1))
        r["event_collision" if is_collision else "event_mallet"] = np.array([getattr(ev, f) for f in FIELDS], np.float64)
    return r


def compare(mine, want, where):
    for key in ("striker_mass", "imp_curvature", "imp_inv_mass", "inv_modulus"):
        assert close(mine[key], want[key], 1e-12), (where, key, mine[key], want[key])
    assert close(mine["scalars"], want["scalars"], 1e-12), (where, "scalars", mine["scalars"], want["scalars"])
    scale = max(float(np.abs(want["inv_inertia"]).max()), 1e-30)
    assert np.abs(mine["inv_inertia"] - want["inv_inertia"]).max() <= 2e-6 * scale, (where, "inv_inertia")
    for key in ("reduced_mass", "tau"):
        assert close(mine[key], want[key], 2e-6), (where, key, mine[key], want[key])
    assert close(mine["click"], want["click"], 1e-6), (where, "click", mine["click"], want["click"])
    for key in ("event_mallet", "event_collision"):
        assert np.all(np.abs(mine[key] - want[key]) <= 4e-6 * np.maximum(np.abs(want[key]), 1e-30)), (where, key, mine[key], want[key])


def test_against_golden_vectors_of_the_reference():
    cs = cases()
    assert int(GOLDEN["n"]) == len(cs)
    taus = []
    for i, c in enumerate(cs):
        want = {k.split("/", 1)[1]: GOLDEN[k] for k in GOLDEN.files if k.startswith(f"{i}/")}
        compare(ours(c), want, f"case {i}")
        taus.append(float(want["tau"]))
    taus = np.array(taus)
    assert ((taus > 2e-5) & (taus < 5e-2)).sum() >= len(cs) // 2  # most cases exercise the integral, not the clamps


@pytest.mark.skipif(not oc.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_against_the_reference_live():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_contact_golden import evaluate

    for i, c in enumerate(cases(n=40, seed=99)):
        compare(ours(c), evaluate(c), f"live case {i}")


def test_hertz_limits_and_defaults():
    """The two closed-form limits the reference cites (ContactModel.h:69-70) and the documented fallbacks."""
    steel = mc.STEEL
    imp = mc.striker_impactor(mc.striker())
    r, l = float(np.float32(0.01)), float(np.float32(0.19))  # Striker's members are floats
    assert abs(mc.striker_mass(mc.striker()) - 7850.0 * np.pi * (r * r * l + 4.0 / 3.0 * r**3)) < 1e-12
    dyn = mc.ContactDynamics(1e9, np.zeros(9, np.float32), [[0, 0, 0]])  # immovable, no lever: m* = striker mass
    m_star = mc.reduced_contact_mass(dyn, 0, [0, 0, 1], imp)
    assert abs(m_star - 1.0 / (1e-9 + imp.inv_mass)) < 1e-9 * m_star
    v, curv = 1.0, 0.0
    e_inv = mc.contact_constant("inv_effective_modulus", steel, steel)
    r_star = 1.0 / mc.contact_constant("combined_curvature", x=curv, y=imp.curvature)
    hertz = 2.87 * (m_star**2 * e_inv**2 / (r_star * v)) ** 0.2
    tau = mc.estimate_contact_time(dyn, 0, [0, 0, 1], v, steel, curv, 0.0, imp)
    assert abs(tau - hertz) < 0.01 * hertz  # Hertz's 2.87 (m*^2 / (E*^2 R* v))^(1/5)
    assert mc.estimate_contact_time(dyn, 5, [0, 0, 1], v, steel, curv, 0.0, imp) == 2e-5  # index out of range -> MinContactTime
    ev = mc.make_strike_event(0, 0, 1.0, 1.0, [0, 0, 1], dynamics=None, elastic=None)  # no dynamics: 1e-4 s, no click
    assert abs(ev.pulse_step - np.float32(1.0 / (1e-4 * 48000.0))) < 1e-7 and ev.accel_amp == 0.0 and ev.click_b0 == 0.0
