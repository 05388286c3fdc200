"""The device shift-invert operator (me_factor_*: supernodal Cholesky of K - sigma*M + triangular solves) against scipy
on the oracle's matrices. Reference operator: src/audio/CholeskyShiftInvert.cpp:26-62."""
import math

import numpy as np
import pytest

from golden_util import load_golden
from oracle import modal as om

pytestmark = pytest.mark.gpu
SIGMA = -((2 * math.pi * 20.0) ** 2)


def _backward_error(A, x, b):
    """Norm-wise backward error ||Ax - b|| / (||A||_1 ||x|| + ||b||): what a backward-stable solver keeps near machine epsilon
    whatever the conditioning (the plain residual/||b|| grows with cond(A), ~1e8 for steel bodies at this shift)."""
    import scipy.sparse.linalg as spla

    return np.linalg.norm(A @ x - b) / (spla.norm(A, 1) * np.linalg.norm(x) + np.linalg.norm(b))


def _system(points, tets, mat, order):
    M, K, _, _ = om.assemble(points, tets, mat, order)
    return (K.to_scipy_full() - SIGMA * M.to_scipy_full()).tocsc()


@pytest.mark.parametrize("order,dims", [(2, (5, 4, 3)), (1, (12, 10, 9)), (2, (9, 8, 8))])
def test_solve_matches_scipy(order, dims):
    import scipy.sparse.linalg as spla

    from mesheditor_b200 import Factor, FemSystem

    points, tets = om.kuhn_block(*dims, size=(0.5, 0.4, 0.3))
    mat = om.MATERIALS["Steel"]
    A = _system(points, tets, mat, order)
    fem = FemSystem(points, tets, mat, order)
    f = Factor(fem, SIGMA)
    rng = np.random.default_rng(3)
    b = rng.standard_normal(A.shape[0])
    x = f.solve(b)
    expect = spla.spsolve(A, b)
    assert _backward_error(A, x, b) <= 1e-14
    assert np.linalg.norm(A @ x - b) <= 10 * max(np.linalg.norm(A @ expect - b), 1e-12 * np.linalg.norm(b))
    assert np.linalg.norm(x - expect) <= 1e-6 * np.linalg.norm(expect)
    info = f.info
    assert info["dofs"] == A.shape[0] and info["supernodes"] > 0 and info["factor_nonzeros"] >= A.shape[0]


def test_panel_solve_and_unstructured_mesh():
    from mesheditor_b200 import Factor, FemSystem

    g = load_golden("bracket_steel")
    mat = om.MATERIALS["Steel"]
    A = _system(g["points"], g["tets"], mat, 2)
    fem = FemSystem(g["points"], g["tets"], mat, 2)
    f = Factor(fem, SIGMA)
    rng = np.random.default_rng(5)
    B = rng.standard_normal((A.shape[0], 3))
    X = f.solve(B)  # solve_panel: three right-hand sides, column-major
    for k in range(3):
        assert _backward_error(A, X[:, k], B[:, k]) <= 1e-14


@pytest.mark.parametrize("width,macro_panels", [(8, None), (11, None), (20, None), (8, 4), (11, 3), (8, 1)])
def test_wide_panel_solve_matches_single_solves(width, macro_panels, monkeypatch):
    """solve_panel through the 8-wide panel sweeps (one pass over the factor per 8 columns) against column-by-column solves; also
    with runs of chain panels solved through explicit inverses of their diagonal blocks (ME_MACRO_PANELS, symbolic.h)."""
    from mesheditor_b200 import Factor, FemSystem

    if macro_panels:  # (default: macro blocks of 8 panels in the forward sweep only)
        monkeypatch.setenv("ME_MACRO_PANELS", str(macro_panels))
        monkeypatch.setenv("ME_MACRO_BACKWARD", "1")

    points, tets = om.kuhn_block(9, 8, 7, size=(0.5, 0.4, 0.3))
    mat = om.MATERIALS["Ceramic"]
    A = _system(points, tets, mat, 2)
    fem = FemSystem(points, tets, mat, 2)
    f = Factor(fem, SIGMA)
    rng = np.random.default_rng(11)
    B = rng.standard_normal((A.shape[0], width))
    X = f.solve(B)
    for k in range(width):
        assert _backward_error(A, X[:, k], B[:, k]) <= 1e-14
        xs = f.solve(B[:, k].copy())
        assert np.linalg.norm(X[:, k] - xs) <= 1e-9 * np.linalg.norm(xs)


def test_indefinite_shift_fails_like_the_reference():
    from mesheditor_b200 import Factor, FemSystem, MeError
    from mesheditor_b200._lib import ME_FACTOR_FAILED

    points, tets = om.kuhn_block(4, 3, 3, size=(0.4, 0.3, 0.3))
    fem = FemSystem(points, tets, om.MATERIALS["Steel"], 2)
    with pytest.raises(MeError) as err:  # a positive shift inside the spectrum: K - sigma M is not positive definite
        Factor(fem, +1e9)
    assert err.value.status == ME_FACTOR_FAILED
