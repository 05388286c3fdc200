"""CPU-only checks (no GPU): the C-ABI library loads and exports every symbol include/me_modal.h declares, the host-only
entry points (PostprocessModes / RescaleModes / symbolic analysis) agree with the oracle, compute entry points fail loudly
without a device, and the workload generators are well formed."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import modal as om

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built_lib):
    header = open(os.path.join(ROOT, "include", "me_modal.h")).read()
    names = sorted(set(re.findall(r"\b(me_[a-z0-9_]+)\s*\(", header)))
    assert len(names) > 50
    lib = C.CDLL(built_lib)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_compute_entry_points_fail_loudly_without_a_device(built_lib):
    from mesheditor_b200 import FemSystem, MeError, lib
    from mesheditor_b200._lib import ME_CUDA_ERROR

    if lib().me_device_count() > 0:
        pytest.skip("a CUDA device is present")
    points, tets = om.kuhn_block(2, 2, 2)
    with pytest.raises(MeError) as err:
        FemSystem(points, tets, "Steel", 2)
    assert err.value.status == ME_CUDA_ERROR  # no CPU fallback


@pytest.mark.parametrize("material", ["Steel", "Ceramic", "Wood"])
@pytest.mark.parametrize("fundamental", [None, 440.0])
def test_postprocess_modes_matches_oracle_bit_for_bit(built_lib, material, fundamental):
    from mesheditor_b200 import postprocess_modes, solver_config

    rng = np.random.default_rng(11)
    lam = np.sort(np.concatenate([rng.uniform(-1e-3, 1e-3, 6), rng.uniform(1e6, 4e10, 54)]))
    shapes = rng.standard_normal((5, 60, 3)).astype(np.float32)
    positions = rng.standard_normal((5, 3)).astype(np.float32)
    mat = om.MATERIALS[material]
    ours = postprocess_modes(lam, shapes, 0.75, mat, solver_config(num_modes=40, fundamental_freq=fundamental), positions)
    ref = om.postprocess_modes(lam, shapes, 0.75, mat, om.SolverConfig(num_modes=40, num_fem_modes=55, fundamental_freq=fundamental), positions)
    assert len(ref.freqs) > 10
    np.testing.assert_array_equal(ours.freqs, ref.freqs)
    np.testing.assert_array_equal(ours.t60s, ref.t60s)
    np.testing.assert_array_equal(ours.shapes, ref.shapes)
    assert ours.original_fundamental == np.float32(ref.original_fundamental)


def test_postprocess_without_audible_modes_is_empty(built_lib):
    from mesheditor_b200 import postprocess_modes, solver_config

    r = postprocess_modes(np.array([1.0, 2.0, 3.0]), np.zeros((1, 3, 3), np.float32), 1.0, "Steel", solver_config(), np.zeros((1, 3), np.float32))
    assert r.empty  # mesh2modes.cpp:548


@pytest.mark.parametrize("dims", [(6, 6, 6), (20, 4, 4), (13, 11, 9)])
def test_symbolic_analysis_is_structurally_valid(built_lib, dims):
    import scipy.sparse as sp

    from mesheditor_b200 import symbolic_analyse

    points, tets = om.kuhn_block(*dims)
    for order in (1, 2):
        nodes, nc = om.element_nodes(tets, len(points), order)
        npe = nodes.shape[1]
        r = np.repeat(nodes.astype(np.int64), npe, axis=1).ravel()
        c = np.tile(nodes.astype(np.int64), (1, npe)).ravel()
        g = sp.csr_matrix((np.ones(len(r)), (r, c)), shape=(nc, nc))
        g.sum_duplicates()
        xyz = np.zeros((nc, 3), np.float32)
        xyz[: len(points)] = points
        if order == 2:
            for e in range(6):
                a, b = om.EDGE_CORNERS[e]
                xyz[nodes[:, 4 + e]] = 0.5 * (points[nodes[:, a]] + points[nodes[:, b]])
        perm, info = symbolic_analyse(g.indptr, g.indices, xyz)
        assert sorted(perm.tolist()) == list(range(nc))
        assert info["violations"] == 0
        assert info["max_panel_columns"] <= 128 and info["supernodes"] >= 1 and info["levels"] >= 1
        # the dissection must beat the natural (banded) order's fill on a 3-D block: compare against the band bound
        assert info["factor_nonzeros"] >= 9 * g.nnz // 2


def test_workload_generators(built_lib):
    from mesheditor_b200 import workloads as wl

    for points, tets in (wl.kuhn_block(3, 4, 5, (0.3, 0.4, 0.5)), wl.torus_mesh(24, 4)):
        p = points[tets]
        det = np.einsum("ij,ij->i", p[:, 3] - p[:, 0], np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]))
        assert (det > 0).all()  # positively oriented (src/mesh/TetMesh.h:10-13)
    op, ot = om.kuhn_block(3, 4, 5, (0.3, 0.4, 0.5))
    np.testing.assert_array_equal(wl.kuhn_block(3, 4, 5, (0.3, 0.4, 0.5))[1], ot)
    pts, tets, surf = wl.config1_mesh()
    assert len(tets) == 8987 and len(surf) == 2562
    dims = wl.config4_dims()
    assert len(dims) == 64 and 9_000 < 6 * dims[0] ** 3 < 12_000 and 450_000 < 6 * dims[-1] ** 3 < 550_000
    owner = wl.lpt_assign([6.0 * d ** 3 for d in dims], 8)
    loads = np.bincount(owner, weights=[6.0 * d ** 3 for d in dims], minlength=8)
    assert loads.max() / loads.mean() < 1.15  # biggest-first greedy balances the batch


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/audio"), reason="the reference tree is not on this box")
def test_reference_side_bank_binding_compiles_against_the_reference_headers():
    """integration/modal_audio_b200.cpp is written against the reference's own ModalAudio.h / ModalModes.h and include/me_modal.h:
    it must at least be valid C++ there (it cannot run without a GPU, and mesh2modes_b200.cpp needs Eigen, which is not vendored)."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = "/root/reference"
    cmd = ["g++", "-std=c++23", "-fsyntax-only", "-Wall", "-Werror", f"-I{ref}/src", f"-I{root}/include", "-isystem", f"{ref}/lib/glm", "-isystem", f"{ref}/lib/entt/src", "-DGLM_ENABLE_EXPERIMENTAL",
           os.path.join(root, "integration", "modal_audio_b200.cpp")]
    done = subprocess.run(cmd, capture_output=True, text=True)
    assert done.returncode == 0, done.stderr


def test_header_is_plain_c():
    """The drop-in boundary is a C ABI: include/me_modal.h must compile as C11 and as C++17 on its own, warnings as errors."""
    import subprocess

    header = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "me_modal.h")
    for compiler, lang, std in (("gcc", "c", "-std=c11"), ("g++", "c++", "-std=c++17")):
        done = subprocess.run([compiler, std, "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-x", lang, header], capture_output=True, text=True)
        assert done.returncode == 0, done.stderr
