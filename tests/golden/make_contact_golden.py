"""Generates tests/golden/strike/contact.npz: seeded inputs and the outputs of the UNMODIFIED reference ContactModel.cpp /
RecoilClickFilter (through oracle/_ref/libme_ref_audio.so, built by `make -C oracle ref` where /root/reference exists).
Run: python tests/golden/make_contact_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import contact as oc  # noqa: E402

MATERIALS = np.array([  # materials::acoustic::All (AcousticMaterial.h), {Density, YoungModulus, PoissonRatio, Alpha, Beta}
    [2700.0, 7.2e10, 0.19, 6.0, 1e-7], [8000.0, 2.0e11, 0.29, 5.0, 3e-8], [750.0, 1.1e10, 0.25, 60.0, 2e-6], [1070.0, 1.4e9, 0.35, 30.0, 1e-6], [2600.0, 6.2e10, 0.20, 1.0, 1e-7],
])


def cases(n=64, seed=20260710):
    rng = np.random.default_rng(seed)
    out = []
    for k in range(n):
        q = rng.standard_normal(4)
        q /= np.linalg.norm(q)
        n_arms = int(rng.integers(1, 6))
        c = dict(
            obj_mat=MATERIALS[rng.integers(0, len(MATERIALS))], imp_mat=MATERIALS[rng.integers(0, len(MATERIALS))],
            tip_radius=np.float32(rng.uniform(0.002, 0.05)), length=np.float32(rng.uniform(0.02, 0.4)),
            mass=float(np.exp(rng.uniform(np.log(1e-3), np.log(50)))), inertia=np.exp(rng.uniform(np.log(1e-6), np.log(1.0), 3)).astype(np.float32), quat=q.astype(np.float32),
            arms=(rng.standard_normal((n_arms, 3)) * rng.uniform(0.01, 0.5)).astype(np.float32), index=int(rng.integers(0, n_arms + (1 if k % 9 == 0 else 0))),
            direction=rng.standard_normal(3).astype(np.float32), speed=float(rng.choice([-1, 1]) * np.exp(rng.uniform(np.log(1e-3), np.log(20)))),
            curvature=float(rng.choice([0.0, rng.uniform(0.5, 200)])), area=float(rng.choice([0.0, np.exp(rng.uniform(np.log(1e-8), np.log(1e-2)))])),
            scale=float(rng.uniform(0.25, 4)), roughness=float(rng.choice([0.0, np.exp(rng.uniform(np.log(1e-7), np.log(1e-4)))])),
            radius=float(rng.uniform(0.005, 0.5)), rate=float(rng.choice([44100.0, 48000.0, 96000.0])), force=np.float32(rng.uniform(0.05, 5)),
        )
        if k % 11 == 0:
            c["mass"] = 0.0  # degenerate: MinContactTime / zero reduced mass
        out.append(c)
    return out


def evaluate(c):
    curv, inv_mass = oc.striker_impactor(c["imp_mat"], c["tip_radius"], c["length"])
    inv_inertia = oc.inverse_inertia(c["mass"], c["inertia"], c["quat"])
    volume = c["mass"] / c["obj_mat"][0] if c["mass"] > 0 else 0.0
    r = dict(
        striker_mass=oc.striker_mass(c["imp_mat"], c["tip_radius"], c["length"]), imp_curvature=curv, imp_inv_mass=inv_mass, inv_inertia=inv_inertia,
        reduced_mass=oc.reduced_contact_mass(c["mass"], inv_inertia, c["arms"], c["index"], c["direction"], c["imp_mat"], curv, inv_mass),
        tau=oc.estimate_contact_time(c["mass"], inv_inertia, c["arms"], c["index"], c["direction"], c["speed"], c["obj_mat"], c["curvature"], c["area"], c["imp_mat"], curv, inv_mass, c["scale"], c["roughness"]),
        inv_modulus=oc.inv_effective_modulus(c["obj_mat"], c["imp_mat"]), click=oc.click_filter(c["radius"], volume, c["mass"], c["rate"]),
    )
    e = r["inv_modulus"]
    k = oc.contact_scalar("combined_curvature", c["curvature"], curv)
    r["scalars"] = np.array([k, oc.contact_scalar("stiffness", e, k), oc.contact_scalar("patch_radius", c["force"] * 100.0, e, k), oc.contact_scalar("static_penetration", c["force"] * 100.0, oc.contact_scalar("stiffness", e, k)),
                             oc.contact_scalar("saturation_penetration", k, c["area"]), oc.contact_scalar("punch_stiffness", e, c["area"])])
    for is_collision in (False, True):
        ev = oc.trigger_modal_strike(3, min(c["index"], len(c["arms"]) - 1), c["force"], c["speed"], c["direction"] if is_collision else c["direction"] / np.linalg.norm(c["direction"]),
                                     (c["mass"], inv_inertia, c["arms"]) if c["mass"] > 0 else None, c["obj_mat"], (c["imp_mat"], curv, inv_mass), is_collision=is_collision, resultant_index=c["index"] % len(c["arms"]),
                                     curvature=c["curvature"], nominal_area=c["area"], scale_ratio=c["scale"], roughness=c["roughness"], displaced_volume=volume if int(c["mass"] * 1e6) % 2 else 0.0,
                                     radiant_radius=c["radius"], sample_rate=c["rate"])
        r["event_collision" if is_collision else "event_mallet"] = np.array([ev[k] for k in ("jx", "jy", "jz", "pulse_step", "pulse_gamma", "accel_amp", "click_b0", "click_a1", "click_a2")], np.float64)
    return r


if __name__ == "__main__":
    cs = cases()
    rs = [evaluate(c) for c in cs]
    flat = {}
    for i, (c, r) in enumerate(zip(cs, rs)):
        for k, v in r.items():
            flat[f"{i}/{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "strike", "contact.npz"), n=len(cs), **flat)
    print("wrote", len(cs), "cases; tau range", min(r["tau"] for r in rs), max(r["tau"] for r in rs))
