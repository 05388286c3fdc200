"""me_modal_solve (the drop-in for modal::mesh2modes) against the reference's golden modal models and against the oracle.
Tolerances are north_star's: eigenvalues 1e-6 relative, mode shapes 1e-4 subspace sine per degenerate cluster (the golden
arrays are float32, which adds ~1e-7 / ~1e-4 of their own; the float64 gates run against the oracle)."""
import numpy as np
import pytest

from golden_util import clusters, compare_shapes, decay_rates, golden_names, load_golden, subspace_sine
from oracle import modal as om

pytestmark = pytest.mark.gpu


def _config(g, **kw):
    from mesheditor_b200 import solver_config

    return solver_config(num_modes=int(g["num_modes"]), min_mode_freq=float(g["min_freq"]), max_mode_freq=float(g["max_freq"]), **kw)


@pytest.mark.parametrize("name", golden_names())
def test_golden_models(name):
    from mesheditor_b200 import mesh2modes

    g = load_golden(name)
    r = mesh2modes(g["points"], g["tets"], tuple(g["material"].tolist()), g["surface"], config=_config(g))
    assert r.status == 0 and len(r.freqs) == len(g["golden_freqs"])
    np.testing.assert_allclose(r.freqs, g["golden_freqs"], rtol=1e-6)          # eigenvalues within 1e-6 => frequencies within 5e-7 (+ float32)
    np.testing.assert_allclose(decay_rates(r.t60s), g["golden_decay"], rtol=4e-6)
    np.testing.assert_array_equal(r.positions, g["golden_positions"])
    np.testing.assert_array_equal(r.sample_point_of_excitation, np.arange(len(g["surface"]), dtype=np.uint32))
    assert abs(r.mass_props["mass"] - float(g["golden_mass"])) <= 1e-6 * float(g["golden_mass"])
    np.testing.assert_allclose(r.mass_props["inertia_diagonal"], g["golden_inertia"], rtol=2e-5)
    sine, norm = compare_shapes(r.shapes, g["golden_shapes"], g["golden_freqs"])
    assert sine < 2e-4 and norm < 1e-4, (sine, norm)
    p = r.profile
    assert p["dofs"] > 0 and p["op_applications"] >= 65 and p["restarts"] >= 1 and p["kernel_launches"] > 0


@pytest.mark.parametrize("form", ["block", "single"])
@pytest.mark.parametrize("order,dims,modes", [(2, (8, 3, 2), 20), (1, (14, 9, 7), 40)])
def test_eigenpairs_match_oracle_fp64(order, dims, modes, form, monkeypatch):
    """Both forms of the shift-invert iteration (block: 8 Krylov vectors per panel solve; single: the reference's
    vector-at-a-time recurrence) against the oracle in FP64."""
    from mesheditor_b200 import mesh2modes, solver_config

    monkeypatch.setenv("ME_LANCZOS", form)

    points, tets = om.kuhn_block(*dims, size=(0.4, 0.15, 0.1))
    mat = om.MATERIALS["Ceramic"]
    ocfg = om.SolverConfig(num_modes=modes, num_fem_modes=modes + 15, max_mode_freq=1e9)
    ref = om.mesh2modes(points, tets, mat, points[:8].astype(np.float32), config=ocfg, order=order)
    r = mesh2modes(points, tets, mat, points[:8].astype(np.float32), config=solver_config(num_modes=modes, max_mode_freq=1e9, element_order=order), keep_basis=True)
    assert r.status == 0
    lam, ref_lam = r.eigenvalues, ref["eigenvalues"]
    elastic = ref_lam > 1e-3 * ref_lam[-1]  # the six rigid-body eigenvalues are ~0: compare those absolutely
    assert np.abs(lam[elastic] / ref_lam[elastic] - 1).max() <= 1e-6
    assert np.abs(lam[~elastic]).max() <= 1e-6 * ref_lam[-1]
    # full eigenvector basis: per cluster subspace angle against the oracle's M-orthonormal vectors
    Mfull = ref["M"].to_scipy_full()
    basis = r.basis.astype(np.float64)
    n_rigid = int((~elastic).sum())  # the rigid-body modes are one degenerate cluster at lambda ~ 0
    groups = [(0, n_rigid)] + [(n_rigid + lo, n_rigid + hi) for lo, hi in clusters(ref_lam[elastic], rel=1e-5)]
    for lo, hi in groups:
        if hi == len(ref_lam) or hi == lo:
            continue  # the last cluster may be cut by nev
        assert subspace_sine(basis[:, lo:hi], ref["eigenvectors"][:, lo:hi]) <= 1e-4
    gram = basis.T @ (Mfull @ basis)
    assert np.abs(gram - np.eye(gram.shape[0])).max() <= 1e-5  # mass-normalised (float32 basis)
    np.testing.assert_allclose(r.freqs, ref["modes"].freqs, rtol=1e-6)


def test_cancel_and_no_modes_return_empty_results():
    from mesheditor_b200 import mesh2modes, solver_config
    from mesheditor_b200._lib import ME_CANCELLED, ME_NO_MODES, MeJobMonitor

    points, tets = om.kuhn_block(4, 3, 2, size=(0.4, 0.3, 0.2))
    mon = MeJobMonitor(0.0, 1)
    r = mesh2modes(points, tets, "Steel", points[:4].astype(np.float32), monitor=mon)
    assert r.status == ME_CANCELLED and r.empty and len(r.eigenvalues) == 0
    # a 50 MHz audible floor: every elastic mode is below it => empty ModalModes (mesh2modes.cpp:548)
    r = mesh2modes(points, tets, "Steel", points[:4].astype(np.float32), config=solver_config(num_modes=10, min_mode_freq=5.0e7, max_mode_freq=1e9))
    assert r.status == ME_NO_MODES and r.empty


def test_excitations_sharing_a_point_merge():
    from mesheditor_b200 import mesh2modes

    points, tets = om.kuhn_block(4, 3, 2, size=(0.4, 0.3, 0.2))
    ex = np.array([points[5], points[5] + 1e-4, points[9], points[5]], np.float32)
    r = mesh2modes(points, tets, "Ceramic", ex)
    ref_pts, ref_pos, ref_map = om.sample_excitations(points, ex)
    np.testing.assert_array_equal(r.sample_point_of_excitation, ref_map)
    np.testing.assert_array_equal(r.positions, ref_pos)
    assert r.shapes.shape[0] == len(ref_pts) == 2
