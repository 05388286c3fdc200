"""The reference's own solver tests (tests/ModalSolverTest.cpp:228-261) run through me_modal_solve: free-free bars whose
mode families have closed forms, with the reference's shape classifier (Classify, :83-116), tolerances (1 % longitudinal,
5 % torsional / thin bending, 10 % Euler-Bernoulli bending) and default SolverConfig. Plus size-independent properties at
BASELINE.json's full analysis size (configs[2], 998,250 tets), where the CPU oracle cannot follow."""
import math

import numpy as np
import pytest

from oracle import modal as om

pytestmark = pytest.mark.gpu

BENDING_BL = (4.73004074, 7.85320462, 10.9956078)  # ModalSolverTest.cpp:34


def classify(modes, mode, length, width, thickness, nx):
    """Classify (ModalSolverTest.cpp:83-116): kinetic-energy fractions; torsion = energy of the per-slice best-fit rotation."""
    u = modes.shapes[:, mode, :].astype(np.float64)
    p = modes.positions.astype(np.float64)
    ry, rz = p[:, 1] - width / 2, p[:, 2] - thickness / 2
    axial, lat_y, lat_z = (u[:, 0] ** 2).sum(), (u[:, 1] ** 2).sum(), (u[:, 2] ** 2).sum()
    total = axial + lat_y + lat_z
    if total <= 0:
        return "other"
    slices = np.rint(p[:, 0] * nx / length).astype(np.int64)
    rotation = 0.0
    for s in np.unique(slices):
        sel = slices == s
        circulation, r2 = (ry[sel] * u[sel, 2] - rz[sel] * u[sel, 1]).sum(), (ry[sel] ** 2 + rz[sel] ** 2).sum()
        if r2 > 0:
            rotation += circulation * circulation / r2
    if axial / total > 0.85:
        return "longitudinal"
    if rotation / total > 0.85:
        return "torsional"
    lateral = lat_y + lat_z
    if lateral / total > 0.6 and rotation / total < 0.5:
        if lat_y / lateral > 0.8:
            return "bending_y"
        if lat_z / lateral > 0.8:
            return "bending_z"
        return "bending"
    return "other"


def solve_bar(length, width, thickness, mat, nx, ny, nz):
    """SolveBar (ModalSolverTest.cpp:119-128): every mesh point is an excitation position, default SolverConfig."""
    from mesheditor_b200 import mesh2modes

    points, tets = om.kuhn_block(nx, ny, nz, (length, width, thickness))
    r = mesh2modes(points, tets, mat, points.astype(np.float32))
    assert r.status == 0 and len(r.freqs) > 0
    fem = {}
    for mode in range(len(r.freqs)):
        fem.setdefault(classify(r, mode, length, width, thickness, nx), []).append(float(r.freqs[mode]))
    return fem


def bending_theory(length, mat, thickness, per_root):
    base = math.sqrt(mat.young / mat.density) * (thickness / math.sqrt(12.0)) / (2 * math.pi * length * length)
    return [bl * bl * base for bl in BENDING_BL for _ in range(per_root)]


def check_family(fem, theory, tolerance, min_count=2):
    count = min(len(fem), len(theory))
    assert count >= min_count
    for f, t in zip(fem[:count], theory[:count]):
        assert abs(f / t - 1.0) < tolerance, (f, t)


def test_square_bar_modes_match_closed_forms():
    length, a = 0.3, 0.05
    mat = om.Material(1000.0, 1e7, 0.0, 0.0, 0.0)
    speed = math.sqrt(mat.young / mat.density)
    torsion_f1 = math.sqrt(mat.mu / mat.density * 0.140577 * 6) / (2 * length)
    fem = solve_bar(length, a, a, mat, 20, 4, 4)
    bending = sorted(fem.get("bending", []) + fem.get("bending_y", []) + fem.get("bending_z", []))[:2]
    check_family(fem.get("longitudinal", []), [n * speed / (2 * length) for n in (1, 2, 3)], 0.01)
    check_family(fem.get("torsional", []), [n * torsion_f1 for n in (1, 2, 3)], 0.05)
    check_family(bending, bending_theory(length, mat, a, 2), 0.10)


def test_thin_bar_bending_matches_closed_forms():
    length, width, thickness = 0.3, 0.05, 0.01
    mat = om.Material(1000.0, 1e9, 0.0, 0.0, 0.0)
    speed = math.sqrt(mat.young / mat.density)
    fem = solve_bar(length, width, thickness, mat, 30, 5, 1)
    check_family(fem.get("longitudinal", []), [n * speed / (2 * length) for n in (1, 2, 3)], 0.01)
    check_family(fem.get("bending_y", []), bending_theory(length, mat, width, 1)[:1], 0.10, 1)
    check_family(fem.get("bending_z", []), bending_theory(length, mat, thickness, 1), 0.05)


def solve_properties(points, tets, material, order, modes, dofs=None, cols=None, warm=True):
    """What can be checked beyond the CPU oracle's reach, through independent device operators (me_fem_spmv, me_factor_solve):
    (1) the six rigid-body modes are there (|lambda| tiny), the model starts above them, eigenvalues ascend;
    (2) Rayleigh quotients x^T K x / x^T M x of the returned (float32) basis reproduce the eigenvalues;
    (3) shift-invert residual ||A^-1 (K - lambda M) x||_2 / ||x||_2 with A = K - sigma M: the quantity the iteration
        converges (a plain ||K x - lambda M x|| would be dominated by lambda_max times the float32 rounding noise of the
        returned basis). A^-1 (K - lambda M) still amplifies the part of that noise lying in the six rigid-body modes by
        (lambda - sigma) / |sigma| ~ 1e6, so x is first M-orthogonalised against the ANALYTIC rigid-body modes
        (translations, e_a x p) in FP64; what remains is the float32 rounding itself, ~2.5e-8;
    (4) the basis is M-orthonormal;
    (5) a warm re-solve seeded by that basis reproduces the eigenvalues to the warm tolerance in a few block iterations."""
    from mesheditor_b200 import Factor, FemSystem, mesh2modes, solver_config
    from mesheditor_b200 import workloads as wl

    ex = wl.bench_excitations(points)
    cfg = solver_config(num_modes=modes, element_order=order, max_mode_freq=1e9)
    r = mesh2modes(points, tets, material, ex, config=cfg, keep_basis=True)
    assert r.status == 0 and len(r.freqs) == modes
    if dofs is not None:
        assert r.profile["dofs"] == dofs
    nev = modes + 15
    lam = r.eigenvalues
    assert len(lam) == nev and np.all(np.diff(lam) >= -1e-9 * lam[-1])
    assert np.abs(lam[:6]).max() < 1e-6 * lam[6] and lam[6] > 0  # rigid-body modes
    assert abs(float(r.freqs[0]) - math.sqrt(lam[6]) / (2 * math.pi)) < 1e-3 * float(r.freqs[0])  # material damping shifts f by ~1e-9
    fem = FemSystem(points, tets, material, order)
    n = fem.info["dofs"]
    # Node coordinates for the analytic rigid-body modes: the corners, and for the 10-node element the edge midpoints
    # (straight-sided tets; edge slots in the order 01, 02, 03, 12, 13, 23 of mesh2modes.cpp:200).
    node_xyz = np.asarray(points, np.float64)
    if order == 2:
        nodes = fem.element_nodes()
        node_xyz = np.zeros((n // 3, 3))
        node_xyz[:len(points)] = points
        for slot, (a, b) in enumerate([(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]):
            node_xyz[nodes[:, 4 + slot]] = 0.5 * (node_xyz[nodes[:, a]] + node_xyz[nodes[:, b]])
    sigma = -((2 * math.pi * 20.0) ** 2)
    factor = Factor(fem, sigma)
    cols = cols or [6, 7, 8, nev // 10 + 6, nev // 4, nev // 2, nev - 35, nev - 1]
    Mx = {}
    rigid = np.zeros((n, 6))
    for a in range(3):
        rigid[a::3, a] = 1.0
        rigid[:, 3 + a] = np.cross(np.eye(3)[a], node_xyz - node_xyz.mean(axis=0)).reshape(-1)
    m_rigid = np.stack([fem.spmv("M", rigid[:, i]) for i in range(6)], axis=1)
    gram = rigid.T @ m_rigid
    for j in cols:
        x = r.basis[:, j].astype(np.float64)
        kx, mx = fem.spmv("K", x), fem.spmv("M", x)
        Mx[j] = mx
        assert abs(float(x @ kx) / float(x @ mx) / lam[j] - 1) < 1e-5, j
        xd = x - rigid @ np.linalg.solve(gram, m_rigid.T @ x)
        assert np.linalg.norm(xd - x) < 1e-9 * np.linalg.norm(x), j  # the elastic modes hold no rigid-body motion
        z = factor.solve(fem.spmv("K", xd) - lam[j] * fem.spmv("M", xd))
        assert np.linalg.norm(z) / np.linalg.norm(xd) < 1e-6, (j, np.linalg.norm(z) / np.linalg.norm(xd))
    for i in cols:
        for j in cols:
            assert abs(float(r.basis[:, i].astype(np.float64) @ Mx[j]) - (1.0 if i == j else 0.0)) < 1e-5, (i, j)
    del factor
    if warm:
        again = mesh2modes(points, tets, material, ex, config=cfg, seed_basis=r.basis)
        assert again.status == 0 and len(again.freqs) == modes
        assert np.abs(again.eigenvalues[6:6 + modes] / lam[6:6 + modes] - 1).max() < 1e-4
        assert abs(float(again.freqs[0]) - float(r.freqs[0])) < 0.05
        assert again.profile["restarts"] <= 5  # an exact seed re-locks in a few block iterations
    return r


def test_full_size_solve_properties():
    """BASELINE.json configs[2] (55^3-cell block, 998,250 tets, P1, 200 modes) is beyond the CPU oracle: see solve_properties."""
    from mesheditor_b200 import workloads as wl

    points, tets = wl.kuhn_block(55, 55, 55, (0.3, 0.3, 0.3))
    solve_properties(points, tets, "Steel", 1, 200, dofs=526848, cols=[6, 7, 8, 20, 57, 111, 180, 214])


@pytest.mark.parametrize("order", [1, 2], ids=["P1", "P2"])
def test_torus_config_solve_properties(order):
    """BASELINE.json configs[1]: the synthetic ~200k-tet torus (210,912 Kuhn tets over a swept 13 x 13 x 208 grid, R = 0.10 m,
    r = 0.03 m), Ceramic, lowest 100 modes, in the order BASELINE names (P1) and in the reference's own element (P2: 10-node
    tets, 0.9 M DOFs, the only order with reference parity - SURVEY.md F1, section 8d "both orders for C2")."""
    from mesheditor_b200 import workloads as wl

    points, tets = wl.torus_mesh()
    assert len(tets) == 210_912
    r = solve_properties(points, tets, "Ceramic", order, 100, warm=(order == 1))
    # a body of revolution: the elastic spectrum comes in degenerate pairs (cos / sin of the azimuth) next to a few singlets
    lam = r.eigenvalues[6:]
    pairs = int(np.sum(np.abs(np.diff(lam)) < 1e-6 * lam[1:]))
    assert pairs >= 30


@pytest.mark.parametrize("order", [1, 2], ids=["P1", "P2"])
def test_coarse_torus_matches_oracle(order):
    """The same torus recipe coarsened until the CPU oracle follows (24 x 3 x 3 cells, 1,296 tets): eigenvalues 1e-6 relative,
    frequencies, and the cluster structure of a body of revolution."""
    from mesheditor_b200 import mesh2modes, solver_config
    from mesheditor_b200 import workloads as wl

    points, tets = wl.torus_mesh(24, 3)
    mat = om.MATERIALS["Ceramic"]
    ex = wl.bench_excitations(points)
    ocfg = om.SolverConfig(num_modes=40, num_fem_modes=55, max_mode_freq=1e9)
    ref = om.mesh2modes(points, tets, mat, ex, config=ocfg, order=order)
    r = mesh2modes(points, tets, mat, ex, config=solver_config(num_modes=40, max_mode_freq=1e9, element_order=order))
    assert r.status == 0
    lam, ref_lam = r.eigenvalues, ref["eigenvalues"]
    elastic = ref_lam > 1e-3 * ref_lam[-1]
    assert np.abs(lam[elastic] / ref_lam[elastic] - 1).max() <= 1e-6
    assert np.abs(lam[~elastic]).max() <= 1e-6 * ref_lam[-1]
    np.testing.assert_allclose(r.freqs, ref["modes"].freqs, rtol=1e-6)
