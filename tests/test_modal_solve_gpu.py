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


def _nu_edit(mat):
    """The reference's Poisson-ratio edit (tests/ModalSolverBench.cpp:388-389)."""
    return om.Material(mat.density, mat.young, min(mat.poisson + 0.02, 0.49), mat.alpha, mat.beta)


@pytest.mark.parametrize("order,dims,modes", [(2, (8, 3, 2), 20), (1, (14, 9, 7), 40)])
def test_warm_resolve_matches_cold_and_oracle(order, dims, modes):
    """SolveReuse::SeedBasis -> SubspaceIterate (mesh2modes.cpp:339-428, :459-472): the edit loop of
    tests/ModalSolverBench.cpp:346-411. Solve cold keeping the basis, edit the Poisson ratio, re-solve warm; the bench's
    acceptance is equal mode counts and |f1 warm - f1 cold| < 0.05 Hz (:384). Beyond that: eigenvalues against the cold
    FP64 oracle at the warm tolerance (pairs lock when their relative change drops under WarmTolerance = 1e-4, which is
    the reference's accuracy on this path, not 1e-6), against the oracle's own restatement of the iteration, and
    M-orthonormal Ritz vectors."""
    from mesheditor_b200 import mesh2modes, solver_config

    points, tets = om.kuhn_block(*dims, size=(0.4, 0.15, 0.1))
    mat = om.MATERIALS["Ceramic"]
    edited = _nu_edit(mat)
    ex = points[:8].astype(np.float32)
    cfg = solver_config(num_modes=modes, max_mode_freq=1e9, element_order=order)
    initial = mesh2modes(points, tets, mat, ex, config=cfg, keep_basis=True)
    assert initial.status == 0 and initial.basis.shape[1] == modes + 15
    cold = mesh2modes(points, tets, edited, ex, config=cfg)
    warm = mesh2modes(points, tets, edited, ex, config=cfg, seed_basis=initial.basis, keep_basis=True)
    assert warm.status == 0 and cold.status == 0
    assert len(warm.freqs) == len(cold.freqs) and abs(float(warm.freqs[0]) - float(cold.freqs[0])) < 0.05
    p = warm.profile
    assert p["restarts"] >= 2 and p["op_applications"] >= 2 * (modes + 15) - modes - 15  # block iterations x active widths
    assert p["op_applications"] <= p["restarts"] * (modes + 30)
    ocfg = om.SolverConfig(num_modes=modes, num_fem_modes=modes + 15, max_mode_freq=1e9)
    ref = om.mesh2modes(points, tets, edited, ex, config=ocfg, order=order)
    ref_warm = om.mesh2modes(points, tets, edited, ex, config=ocfg, order=order, seed_basis=initial.basis)
    ref_lam = ref["eigenvalues"]
    elastic = ref_lam > 1e-3 * ref_lam[-1]
    for lam in (warm.eigenvalues, ref_warm["eigenvalues"]):
        assert len(lam) == modes + 15
        # the pairs that survive into the model (lowest `modes` elastic ones) sit far inside the locked prefix
        keep = np.flatnonzero(elastic)[:modes]
        assert np.abs(lam[keep] / ref_lam[keep] - 1).max() <= 1e-4
        assert np.abs(lam[elastic] / ref_lam[elastic] - 1).max() <= 1e-3  # the trailing pairs converge slowest
    assert np.abs(warm.eigenvalues[keep] / ref_warm["eigenvalues"][keep] - 1).max() <= 1e-4  # device vs restated iteration
    np.testing.assert_allclose(warm.freqs, cold.freqs, rtol=1e-4)
    # same iteration, same seed: the device run and the oracle's restatement take the same number of block iterations (+-1)
    assert abs(int(p["restarts"]) - int(ref_warm["iterations"])) <= 2  # (the Gaussian filler columns differ: libstdc++ vs numpy)
    Mfull = ref["M"].to_scipy_full()
    basis = warm.basis.astype(np.float64)
    gram = basis.T @ (Mfull @ basis)
    assert np.abs(gram - np.eye(gram.shape[0])).max() <= 1e-5
    n_rigid = int((~elastic).sum())
    groups = [(0, n_rigid)] + [(n_rigid + lo, n_rigid + hi) for lo, hi in clusters(ref_lam[elastic], rel=1e-5)]
    for lo, hi in groups:
        if hi > modes or hi == lo:
            continue
        assert subspace_sine(basis[:, lo:hi], ref["eigenvectors"][:, lo:hi]) <= 2e-2  # vectors converge as the sqrt of the values


def test_mismatched_seed_falls_back_to_the_cold_path():
    """A basis solved over a different mesh cannot seed this solve (mesh2modes.cpp:459-464)."""
    from mesheditor_b200 import mesh2modes, solver_config

    points, tets = om.kuhn_block(5, 3, 2, size=(0.4, 0.3, 0.2))
    cfg = solver_config(num_modes=10, max_mode_freq=1e9)
    cold = mesh2modes(points, tets, "Steel", points[:4].astype(np.float32), config=cfg)
    wrong_rows = np.zeros((cold.profile["dofs"] + 3, 25), np.float32)
    too_few_cols = np.zeros((cold.profile["dofs"], 24), np.float32)
    for seed in (wrong_rows, too_few_cols):
        r = mesh2modes(points, tets, "Steel", points[:4].astype(np.float32), config=cfg, seed_basis=seed)
        assert r.status == 0 and r.profile["op_applications"] == cold.profile["op_applications"]
        np.testing.assert_array_equal(r.freqs, cold.freqs)


def test_solved_model_round_trips_through_the_interchange_formats():
    """A fresh solve -> `.modal` bytes -> parse gives the same model back; the MeshEditorModalSolve JSON relabels the mesh's
    triangles onto the sample points (tests/ModalSolveTool.cpp:84-96). Byte parity of the format itself: test_interchange_cpu.py."""
    import json

    from mesheditor_b200 import solver_config
    from mesheditor_b200.interchange import ModalModel

    points, tets = om.kuhn_block(6, 3, 2, (0.3, 0.15, 0.1))
    excite = np.vstack([points[:12], points[3:5]]).astype(np.float32)  # the last two repeat earlier points: they merge
    model = ModalModel.solve(points, tets, "Steel", excite, config=solver_config(num_modes=12, num_fem_modes=20, element_order=2), vertices=np.arange(12), solved_num_modes=12,
                             tet_positions=points[:5], tet_edge_indices=[0, 1, 1, 2], solved_vertices=np.arange(14), tet_inputs_hash=12345)
    r = model.result
    assert not r.empty and len(r.sample_point_of_excitation) == 14 and len(r.positions) == 12
    back = ModalModel.from_bytes(model.to_bytes())
    for a, b in ((back.result.freqs, r.freqs), (back.result.t60s, r.t60s), (back.result.shapes, r.shapes), (back.result.positions, r.positions), (back.result.eigenvalues, r.eigenvalues),
                 (back.result.summary_shapes, r.summary_shapes), (back.vertices, model.vertices), (back.solved_vertices, model.solved_vertices), (back.tet_positions, model.tet_positions)):
        np.testing.assert_array_equal(a, b)
    assert back.result.mass_props == r.mass_props and back.tet_inputs_hash == 12345 and back.solved_num_modes == 12
    assert back.to_bytes() == model.to_bytes()
    triangles = [0, 1, 2, 3, 12, 4, 5, 6, 13]  # excitation 12 is point 3 (degenerate with corner 3), 13 is point 4
    d = json.loads(model.solve_json(triangles))
    sp = r.sample_point_of_excitation
    want = []
    for t in range(0, len(triangles), 3):
        a, b, c = (int(sp[i]) for i in triangles[t:t + 3])
        if len({a, b, c}) == 3:
            want += [a, b, c]
    assert d["indices"] == want and len(want) == 6
    np.testing.assert_array_equal(np.asarray(d["frequencies"], np.float32), r.freqs)


def test_concurrent_solves_match_sequential():
    """Several me_modal_solve calls in flight on one device (host threads, every solve on its own stream: how bench.py runs
    BASELINE.json configs[3]) give what the same calls give one after the other."""
    from concurrent.futures import ThreadPoolExecutor

    from mesheditor_b200 import mesh2modes, solver_config

    mat = om.MATERIALS["Steel"]
    jobs = [om.kuhn_block(*dims, size=(0.3, 0.25, 0.2)) for dims in ((12, 10, 9), (7, 6, 5), (16, 12, 11), (9, 9, 8), (5, 4, 4), (14, 13, 10))]
    cfg = solver_config(num_modes=24, max_mode_freq=1e9, element_order=1)

    def solve(job):
        points, tets = job
        r = mesh2modes(points, tets, mat, points[:6].astype(np.float32), config=cfg)
        assert r.status == 0
        return r.eigenvalues.copy(), r.freqs.copy()

    sequential = [solve(j) for j in jobs]
    for _ in range(2):
        with ThreadPoolExecutor(max_workers=3) as pool:
            concurrent = list(pool.map(solve, jobs))
        for (lam, freqs), (ref_lam, ref_freqs) in zip(concurrent, sequential):
            elastic = ref_lam > 1e-3 * ref_lam[-1]
            assert np.abs(lam[elastic] / ref_lam[elastic] - 1).max() <= 1e-9  # (the factor's atomics reorder sums at rounding level)
            np.testing.assert_allclose(freqs, ref_freqs, rtol=1e-6)
