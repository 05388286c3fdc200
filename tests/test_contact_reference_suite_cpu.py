"""The reference's own contact-model suite (tests/ContactModelTest.cpp) restated against the C ABI's strike front-end
(me_inverse_inertia_tensor, me_estimate_contact_time, me_contact_constant, me_striker_impactor): same bodies, same strikes, same
expectations and tolerances, test for test. Host-only."""
import math

import numpy as np
import pytest

from mesheditor_b200 import contact as mc

pytestmark = pytest.mark.usefixtures("built_lib")

POLYMER = (1000.0, 1e9, 0.3, 0.0, 0.0)  # ContactModelTest.cpp:22
CERAMIC = (2700.0, 7.2e10, 0.19, 0.0, 0.0)  # :23
MIN_CONTACT_TIME, MAX_CONTACT_TIME = 2e-5, 5e-2  # ContactModel.h:98
IDENTITY, ZERO = np.eye(3, dtype=np.float32), np.zeros((3, 3), np.float32)


def null_striker():
    """ContactModelTest.cpp:16-20: mass, stiffness and tip curvature all vanish from the harmonic sums."""
    return mc.striker((1e6, 1e30, 0.0, 0.0, 0.0), tip_radius=1e6, length=1e6)


def body(mass, inverse_inertia, arm=(0.0, 0.0, 0.0)):
    """ContactModelTest.cpp:27-33."""
    return mc.ContactDynamics(mass, inverse_inertia, [arm])


def contact_time(d, material, curvature, area=0.0, speed=1.0, scale=1.0, striker=None):
    """ContactModelTest.cpp:36-38: a strike along +z."""
    return mc.estimate_contact_time(d, 0, [0, 0, 1], speed, material, curvature, area, mc.striker_impactor(striker if striker is not None else null_striker()), scale)


def near(a, b, rel=1e-6):
    """tests/Near.h: relative to the larger magnitude."""
    return abs(a - b) <= rel * max(abs(a), abs(b), 1e-300)


def test_inverse_inertia_round_trips_a_principal_decomposition():
    q = np.array([0.3, 0.1, -0.5, 0.8])  # w, x, y, z
    q = (q / np.linalg.norm(q)).astype(np.float32)
    w, x, y, z = (float(v) for v in q)
    rot = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    inertia = rot @ np.diag([2.0, 5.0, 9.0]) @ rot.T
    inverse = mc.inverse_inertia_tensor(1.0, [2.0, 5.0, 9.0], q).reshape(3, 3)  # symmetric: row- or column-major alike
    np.testing.assert_allclose(inertia @ inverse, np.eye(3), atol=1e-4)


def test_contact_time_matches_the_hertz_formula():
    tau = contact_time(body(1.0, IDENTITY), POLYMER, 100)
    assert near(tau, 1.744e-3, 2e-2)  # hand-computed: 2.87 ((1 (1 - 0.09) / 1e9)^2 100)^0.2


def test_effective_mass_drops_with_an_off_centre_strike():
    assert contact_time(body(1.0, IDENTITY, (0.2, 0, 0)), POLYMER, 100) < contact_time(body(1.0, IDENTITY), POLYMER, 100)


def test_scale_ratio_and_clamping():
    d = body(1.0, IDENTITY)
    tau = lambda scale: contact_time(d, POLYMER, 100, 0, 1, scale)  # noqa: E731
    assert near(tau(2.0), 2 * tau(1.0), 1e-6)
    assert near(tau(100.0), MAX_CONTACT_TIME) and near(tau(1e-6), MIN_CONTACT_TIME)


def test_the_contact_time_reaches_both_of_its_limits():
    d = body(1.0, ZERO)
    inv_modulus = 0.91 / 1e9
    tau = lambda curvature, area, speed: contact_time(d, POLYMER, curvature, area, speed)  # noqa: E731
    curvature, area = 100.0, 1e-4
    hertz = 2.868 * (inv_modulus**2 * curvature) ** 0.2
    assert near(tau(curvature, 0.0, 1.0), hertz, 1e-3)
    punch = math.pi * math.sqrt(inv_modulus / (2 * math.sqrt(area / math.pi)))
    assert near(tau(0.0, area, 1.0), punch, 1e-3)
    assert near(tau(curvature, 0.0, 32.0) / tau(curvature, 0.0, 1.0), 32.0**-0.2, 1e-3)
    assert near(tau(0.0, area, 32.0) / tau(0.0, area, 1.0), 1.0, 1e-3)


def test_filling_the_patch_stops_the_contact_stiffening():
    d = body(0.5, ZERO)
    curvature, area = 10.0, 1e-5
    tau = lambda a, speed: contact_time(d, CERAMIC, curvature, a, speed)  # noqa: E731
    assert near(mc.contact_constant("saturation_penetration", x=curvature, y=area), 3.183e-5, 1e-3)
    assert near(tau(area, 0.1), tau(0.0, 0.1), 1e-6)
    assert near(tau(1.0, 3.0), tau(0.0, 3.0), 1e-6)
    assert tau(area, 3.0) > tau(0.0, 3.0)
    assert tau(area, 3.0) > math.pi * math.sqrt(0.5 / mc.contact_constant("punch_stiffness", x=0.91 / 7.2e10, y=area))
    assert near(tau(1.7e-5, 1.0), tau(1.5e-5, 1.0), 1e-3)
    hertz_ratio, saturating_ratio = tau(0.0, 3.0) / tau(0.0, 0.1), tau(area, 3.0) / tau(area, 0.1)
    assert near(hertz_ratio, 30.0**-0.2, 1e-3)
    assert hertz_ratio < saturating_ratio < 1.0


def test_a_lighter_striker_shortens_the_contact_against_a_heavy_object():
    d = body(1000.0, ZERO)
    light, heavy = mc.striker(length=0.05), mc.striker(length=5.0)
    assert contact_time(d, CERAMIC, 5, 0, 1, 1, light) < contact_time(d, CERAMIC, 5, 0, 1, 1, heavy)
