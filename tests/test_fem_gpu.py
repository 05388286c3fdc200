"""Parity of the device FEM assembly (me_fem_*) with the oracle restatement of the reference's BuildQuadMesh /
AssembleQuadratic (oracle/modal.py): node numbering and CSC sparsity pattern bit-exact, values to 1e-12 of the largest
entry (FP64; the reference sums triplets in element order and so does the device gather), colouring bit-exact,
SpMV against scipy."""
import numpy as np
import pytest

from golden_util import load_golden
from oracle import modal as om

pytestmark = pytest.mark.gpu


def _cases():
    g = load_golden("bracket_steel")
    yield "bracket", g["points"], g["tets"]
    g = load_golden("marble_glass")
    yield "marble", g["points"], g["tets"]
    yield "kuhn_6x5x4", *om.kuhn_block(6, 5, 4, (0.3, 0.25, 0.2))


CASES = list(_cases())


@pytest.mark.parametrize("order", [2, 1])
@pytest.mark.parametrize("name,points,tets", CASES, ids=[c[0] for c in CASES])
def test_pattern_and_values_match_oracle(name, points, tets, order):
    from mesheditor_b200 import FemSystem

    mat = om.MATERIALS["Steel"]
    M, K, nodes, node_count = om.assemble(points, om.filter_degenerate(points, tets), mat, order)
    fem = FemSystem(points, tets, mat, order)
    assert fem.info["node_count"] == node_count and fem.info["dofs"] == 3 * node_count
    np.testing.assert_array_equal(fem.element_nodes(), nodes)  # first-seen midside numbering, bit-exact
    for which, ref in (("K", K), ("M", M)):
        colptr, rowidx, values = fem.csc(which)
        np.testing.assert_array_equal(colptr.astype(np.int64), ref.colptr)   # CSC pattern, bit-exact
        np.testing.assert_array_equal(rowidx.astype(np.int64), ref.rowidx)
        scale = np.abs(ref.values).max()
        assert np.abs(values - ref.values).max() <= 1e-12 * scale
        big = np.abs(ref.values) > 1e-6 * scale
        assert np.abs(values[big] / ref.values[big] - 1).max() <= 1e-9


def test_degenerate_tets_are_dropped():
    from mesheditor_b200 import FemSystem

    points, tets = om.kuhn_block(3, 3, 3)
    flat = np.array([[0, 1, 2, 3]], np.uint32)  # four collinear-ish grid points along z: zero volume
    tets2 = np.concatenate([tets[:10], flat, tets[10:]])
    assert len(om.filter_degenerate(points, tets2)) == len(tets)
    fem = FemSystem(points, tets2, om.MATERIALS["Ceramic"], 2)
    assert fem.info["tets_kept"] == len(tets)
    M, K, nodes, nc = om.assemble(points, tets, om.MATERIALS["Ceramic"], 2)
    np.testing.assert_array_equal(fem.element_nodes(), nodes)
    np.testing.assert_array_equal(fem.csc("K")[1].astype(np.int64), K.rowidx)


@pytest.mark.parametrize("order", [2, 1])
def test_colouring_matches_sequential_first_fit(order):
    from mesheditor_b200 import FemSystem

    g = load_golden("slab_ceramic")
    nodes, nc = om.element_nodes(g["tets"], len(g["points"]), order)
    expect = om.greedy_colouring(nodes, nc)
    fem = FemSystem(g["points"], g["tets"], om.MATERIALS["Ceramic"], order)
    colours, n = fem.colour_elements()
    np.testing.assert_array_equal(colours, expect)
    assert n == expect.max() + 1
    # it is a proper colouring: no two elements of one colour share a node
    for c in range(n):
        used = nodes[colours == c].ravel()
        assert len(np.unique(used)) == len(used)


@pytest.mark.parametrize("order", [2, 1])
def test_spmv_matches_scipy(order):
    from mesheditor_b200 import FemSystem

    points, tets = om.kuhn_block(7, 6, 5, (0.7, 0.6, 0.5))
    mat = om.MATERIALS["Glass"]
    M, K, _, _ = om.assemble(points, tets, mat, order)
    fem = FemSystem(points, tets, mat, order)
    rng = np.random.default_rng(7)
    x = rng.standard_normal(M.n)
    for which, ref in (("K", K), ("M", M)):
        y, expect = fem.spmv(which, x), ref.to_scipy_full() @ x
        assert np.abs(y - expect).max() <= 1e-12 * np.abs(expect).max()


def test_bad_arguments_fail_loudly():
    from mesheditor_b200 import FemSystem, MeError

    points, tets = om.kuhn_block(2, 2, 2)
    with pytest.raises(MeError):
        FemSystem(points, tets, om.MATERIALS["Steel"], 3)
    bad = tets.copy()
    bad[0, 0] = 10_000
    with pytest.raises(MeError):
        FemSystem(points, bad, om.MATERIALS["Steel"], 2)
