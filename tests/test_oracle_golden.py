"""Pins the analysis oracle (oracle/modal.py) to the reference's golden modal models:
frequencies, decay rates, sample positions, mass properties and per-cluster shape subspaces of models solved by the
macOS reference binary over the same tet meshes (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from oracle import modal as om
from golden_util import compare_shapes, decay_rates, golden_names, load_golden


def solve_golden(g):
    material = om.Material(*g["material"].tolist())
    config = om.SolverConfig(min_mode_freq=float(g["min_freq"]), max_mode_freq=float(g["max_freq"]), num_modes=int(g["num_modes"]), num_fem_modes=int(g["num_modes"]) + 15)
    return om.mesh2modes(g["points"], g["tets"], material, g["surface"], config=config)


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_golden_model(name):
    g = load_golden(name)
    r = solve_golden(g)
    modes = r["modes"]
    assert len(modes.freqs) == len(g["golden_freqs"])
    # float32 frequencies: the survey's replay matched bit for bit on the bar; allow 2 ulp for LAPACK/ARPACK differences.
    np.testing.assert_allclose(modes.freqs, g["golden_freqs"], rtol=2.5e-7, atol=0)
    np.testing.assert_allclose(decay_rates(modes.t60s), g["golden_decay"], rtol=2e-6)
    np.testing.assert_array_equal(modes.positions, g["golden_positions"])
    assert abs(r["mass_props"]["mass"] - float(g["golden_mass"])) <= 1e-6 * float(g["golden_mass"])
    np.testing.assert_allclose(r["mass_props"]["inertia"], g["golden_inertia"], rtol=2e-5)
    sine, norm = compare_shapes(modes.shapes, g["golden_shapes"], g["golden_freqs"])
    assert sine < 2e-4 and norm < 1e-4, (sine, norm)


def test_pattern_counts_match_survey_probe():
    # 20x4x4 Kuhn bar, P2: 9,963 DOFs and 373,626 stored lower entries of K; P1: 1,575 DOFs / 55,053 full.
    pts, tets = om.kuhn_block(20, 4, 4, (1.0, 0.2, 0.2))
    mat = om.MATERIALS["Steel"]
    M, K, nodes, nc = om.assemble(pts, tets, mat, 2)
    assert (3 * nc, len(K.values)) == (9963, 373626)
    M1, K1, _, nc1 = om.assemble(pts, tets, mat, 1)
    assert 3 * nc1 == 1575 and 2 * len(K1.values) - 3 * nc1 == 55053


def test_bar_closed_forms():
    """tests/ModalSolverTest.cpp:228-245: 20x4x4 square bar, nu=0: longitudinal n*c/2L within 1 %."""
    L, w = 1.0, 0.1
    pts, tets = om.kuhn_block(20, 4, 4, (L, w, w))
    mat = om.Material(1000.0, 1e7, 0.0, 0.0, 0.0)
    r = om.mesh2modes(pts, tets, mat, pts.astype(np.float32))
    freqs = r["modes"].freqs.astype(np.float64)
    f_long = np.sqrt(mat.young / mat.density) / (2 * L)  # 50 Hz
    assert np.min(np.abs(freqs - f_long)) < 0.01 * f_long


def test_oracle_warm_path_reconverges_an_edited_material():
    """oracle.subspace_iterate (SubspaceIterate, mesh2modes.cpp:339-428) seeded by the basis of a cold solve re-converges
    a Poisson-ratio edit to the cold answer: the reference's own acceptance is |delta f1| < 0.05 Hz and equal mode
    counts (tests/ModalSolverBench.cpp:384)."""
    points, tets = om.kuhn_block(6, 3, 2, size=(0.4, 0.15, 0.1))
    mat = om.MATERIALS["Ceramic"]
    edited = om.Material(mat.density, mat.young, mat.poisson + 0.02, mat.alpha, mat.beta)
    cfg = om.SolverConfig(num_modes=12, num_fem_modes=27, max_mode_freq=1e9)
    ex = points[:4].astype(np.float32)
    initial = om.mesh2modes(points, tets, mat, ex, config=cfg)
    cold = om.mesh2modes(points, tets, edited, ex, config=cfg)
    warm = om.mesh2modes(points, tets, edited, ex, config=cfg, seed_basis=initial["eigenvectors"].astype(np.float32))
    assert len(warm["eigenvalues"]) == 27 and 2 <= warm["iterations"] <= 30
    assert len(warm["modes"].freqs) == len(cold["modes"].freqs)
    assert abs(float(warm["modes"].freqs[0]) - float(cold["modes"].freqs[0])) < 0.05
    elastic = cold["eigenvalues"] > 1e-3 * cold["eigenvalues"][-1]
    keep = np.flatnonzero(elastic)[:12]
    # pairs lock when their relative change drops under WarmTolerance = 1e-4: that, not 1e-6, is the reference's warm accuracy
    assert np.abs(warm["eigenvalues"][keep] / cold["eigenvalues"][keep] - 1).max() <= 1e-4
    V = warm["eigenvectors"]
    gram = V.T @ (cold["M"].to_scipy_full() @ V)
    assert np.abs(gram - np.eye(27)).max() <= 1e-8
    # a seed of the wrong height is ignored (cold path)
    again = om.mesh2modes(points, tets, edited, ex, config=cfg, seed_basis=np.zeros((5, 27), np.float32))
    assert again["iterations"] is None
