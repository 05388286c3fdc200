"""Model interchange through the C ABI: `.modal` bytes against the reference's own zpp::bits archive
(committed blobs tests/golden/interchange/*.modal written by oracle/_ref, and oracle/_ref live where it is built), and the
MeshEditorModalSolve JSON against a restatement of tests/ModalSolveTool.cpp:84-123. Host-only. Bar: byte-exact."""
import json
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_interchange_golden import SEEDS  # noqa: E402

from mesheditor_b200 import MeError  # noqa: E402
from mesheditor_b200.interchange import LN1000, ModalModel, bank_modes, khr_modal_model  # noqa: E402
from oracle import interchange as oi  # noqa: E402

pytestmark = pytest.mark.usefixtures("built_lib")  # builds libme_modal.so on demand (tests/conftest.py)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "interchange")


def check_fields(model: ModalModel, m):
    r = model.result
    np.testing.assert_array_equal(r.freqs, m["freqs"]), np.testing.assert_array_equal(r.t60s, m["t60s"])
    np.testing.assert_array_equal(r.shapes, m["shapes"]), np.testing.assert_array_equal(r.positions, m["positions"])
    np.testing.assert_array_equal(r.eigenvalues, m["eigenvalues"]), np.testing.assert_array_equal(r.summary_shapes, m["summary_shapes"])
    assert r.original_fundamental == float(m["original_fundamental"]) and r.mass_props["mass"] == m["mass"]
    np.testing.assert_array_equal(np.array(r.mass_props["center_of_mass"], np.float32), m["com"])
    np.testing.assert_array_equal(np.array(r.mass_props["inertia_diagonal"], np.float32), m["inertia"])
    np.testing.assert_array_equal(np.array(r.mass_props["inertia_orientation"], np.float32), m["quat_wxyz"])
    for ours, theirs in ((model.vertices, "vertices"), (model.indices, "indices"), (model.tet_edge_indices, "tet_edges"), (model.solved_vertices, "solved_vertices"), (model.tet_positions, "tet_positions")):
        np.testing.assert_array_equal(ours, m[theirs])
    np.testing.assert_array_equal(np.array(model.baked_scale, np.float32), m["baked_scale"])
    sm = model.solved_material
    assert (sm.density, sm.young_modulus, sm.poisson_ratio, sm.alpha, sm.beta) == tuple(m["material"])
    assert (model.solved_min_mode_freq, model.solved_max_mode_freq, model.solved_num_modes, model.tet_inputs_hash) == (float(m["min_freq"]), float(m["max_freq"]), m["num_modes"], m["tet_hash"])


@pytest.mark.parametrize("seed", sorted(SEEDS))
def test_golden_modal_files_parse_and_reserialise_byte_exact(seed):
    with open(os.path.join(GOLDEN, f"model_{seed}.modal"), "rb") as f:
        data = f.read()
    model = ModalModel.from_bytes(data)
    check_fields(model, oi.random_model(seed, **SEEDS[seed]))
    assert model.to_bytes() == data


@pytest.mark.skipif(not oi.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("seed", [101, 102, 103])
def test_against_the_reference_archive_live(seed):
    rng = np.random.default_rng(seed)
    m = oi.random_model(seed, n_modes=int(rng.integers(1, 40)), n_points=int(rng.integers(1, 20)), n_eigen=int(rng.integers(1, 60)))
    data = oi.serialize(m)
    model = ModalModel.from_bytes(data)
    check_fields(model, m)
    ours = model.to_bytes()
    assert ours == data
    assert oi.parses_to(m, ours) == 1  # the reference's loader reads our bytes back to the same ModalModelData


def test_bad_modal_data_is_rejected():
    with open(os.path.join(GOLDEN, "model_3.modal"), "rb") as f:
        data = f.read()
    for bad in (data[:-3], data + b"\0", data[:40]):
        with pytest.raises(MeError):
            ModalModel.from_bytes(bad)


def test_modal_solve_json_matches_the_tool():
    """tests/ModalSolveTool.cpp:84-123 restated: key order, decayRates in float, mode-major shapes, relabelled triangles."""
    with open(os.path.join(GOLDEN, "model_21.modal"), "rb") as f:
        model = ModalModel.from_bytes(f.read())
    r = model.result
    n_points = len(r.positions)
    # a parsed file carries no SamplePointOfExcitation: triangles cannot be relabelled, so an index is an error ...
    with pytest.raises(MeError):
        model.solve_json([0, 1, 2])
    text = model.solve_json()
    d = json.loads(text)
    assert list(d) == ["frequencies", "decayRates", "positions", "shapes", "indices", "mass", "centerOfMass", "inertiaDiagonal"]
    f32 = np.float32
    np.testing.assert_array_equal(np.asarray(d["frequencies"], f32), r.freqs)  # shortest round-trip text: exact float32 back
    want_rates = np.where(r.t60s > 0, f32(LN1000) / r.t60s, f32(0)).astype(f32)
    np.testing.assert_array_equal(np.asarray(d["decayRates"], f32), want_rates)
    np.testing.assert_array_equal(np.asarray(d["positions"], f32), r.positions)
    np.testing.assert_array_equal(np.asarray(d["shapes"], f32).reshape(len(r.freqs), n_points, 3), np.transpose(r.shapes, (1, 0, 2)))
    assert d["indices"] == [] and d["mass"] == r.mass_props["mass"]
    np.testing.assert_array_equal(np.asarray(d["centerOfMass"], f32), np.array(r.mass_props["center_of_mass"], f32))
    # ... and the KHR_audio_rigid_bodies view of it feeds the synthesis bank with the model's own T60s back (to rounding)
    bank = bank_modes(khr_modal_model(text))
    np.testing.assert_allclose(bank["t60s"], r.t60s, rtol=2e-7)
    np.testing.assert_array_equal(bank["shapes"], r.shapes)


def test_khr_audio_rigid_bodies_document_round_trip():
    """A modal model -> glTF document (extensions.KHR_audio_rigid_bodies) -> back: every array bit-identical."""
    from mesheditor_b200.interchange import read_gltf_modal_models, write_gltf_modal_models

    with open(os.path.join(GOLDEN, "model_21.modal"), "rb") as f:
        model = khr_modal_model(ModalModel.from_bytes(f.read()).solve_json())
    model.update(name="Bell", material=dict(name="Steel", density=8000.0, youngsModulus=2.0e11, poissonRatio=0.29, alpha=5.0, beta=3.0e-8), indices=np.array([0, 1, 2, 2, 3, 0], np.uint32))
    second = dict(model, name="Bell2", indices=np.zeros(0, np.uint32))
    doc = json.loads(json.dumps(write_gltf_modal_models([model, second])))  # through text, as a file would go
    assert doc["asset"]["version"] == "2.0" and doc["extensionsUsed"] == ["KHR_audio_rigid_bodies"]
    assert len(doc["extensions"]["KHR_audio_rigid_bodies"]["acousticMaterials"]) == 1
    assert "indices" not in doc["extensions"]["KHR_audio_rigid_bodies"]["modalModels"][1]
    back = read_gltf_modal_models(doc)
    assert [m["name"] for m in back] == ["Bell", "Bell2"]
    for key in ("frequencies", "decayRates", "positions", "shapes", "indices", "centerOfMass", "inertiaDiagonal"):
        np.testing.assert_array_equal(back[0][key], model[key])
    assert back[0]["mass"] == model["mass"] and back[0]["material"] == model["material"] and len(back[1]["indices"]) == 0


@pytest.mark.skipif(not os.path.exists("/root/reference/glTF_PhysicalAudio/samples/Pile.gltf"), reason="the reference's sample documents are only in the build container")
def test_reads_the_reference_golden_documents():
    """The reader decodes the reference's committed samples to exactly the arrays tests/golden/*.npz were made from."""
    import glob

    from golden_util import golden_names, load_golden
    from mesheditor_b200.interchange import read_gltf_modal_models

    models = []
    for path in glob.glob("/root/reference/glTF_PhysicalAudio/samples/**/*.gltf", recursive=True):
        models += read_gltf_modal_models(path)
    assert len(models) >= 7
    for name in golden_names():
        g = load_golden(name)
        match = [m for m in models if len(m["frequencies"]) == len(g["golden_freqs"]) and np.array_equal(m["frequencies"], g["golden_freqs"])]
        assert match, name
        m = match[0]
        np.testing.assert_array_equal(m["decayRates"], g["golden_decay"]), np.testing.assert_array_equal(m["positions"], g["golden_positions"])
        np.testing.assert_array_equal(m["shapes"], g["golden_shapes"])
        assert m["mass"] == float(g["golden_mass"])


def _physical_model(seed):
    """A stored model whose eigenvalues are audible (the random golden models' are, too) with a realistic material."""
    m = oi.random_model(seed, n_modes=9, n_points=6, n_eigen=24)
    m["eigenvalues"] = np.sort(np.random.default_rng(seed).uniform((2 * np.pi * 90.0) ** 2, (2 * np.pi * 7000.0) ** 2, 24))
    return m


@pytest.mark.skipif(not oi.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("seed,density,young", [(5, 2700.0, 7.2e10), (6, 1350.0, 9.0e10), (7, 8000.0, 2.0e11)])
def test_rescale_modes_matches_the_oracle_and_keeps_the_solved_summary(seed, density, young):
    """modal::RescaleModes (mesh2modes.cpp:590-603) through me_rescale_modes: modes bit-equal to the restatement, and the eigen
    summary stays the SOLVED one (the reference never touches ModalEigenSummary), so a second rescale does not compound."""
    from mesheditor_b200 import solver_config
    from oracle import modal as om

    m = _physical_model(seed)
    model = ModalModel.from_bytes(oi.serialize(m))
    solved = om.Material(*m["material"])
    edited = om.Material(density, young, solved.poisson, solved.alpha, solved.beta)
    cfg = solver_config(num_modes=12, max_mode_freq=16000.0)
    want = om.rescale_modes(m["eigenvalues"], m["summary_shapes"], solved, edited, om.SolverConfig(num_modes=12, num_fem_modes=27, max_mode_freq=16000.0), m["positions"])
    got = model.rescaled((density, young, solved.poisson, solved.alpha, solved.beta), cfg)
    np.testing.assert_array_equal(got.result.freqs, want.freqs), np.testing.assert_array_equal(got.result.t60s, want.t60s)
    np.testing.assert_array_equal(got.result.shapes, want.shapes)
    np.testing.assert_array_equal(got.result.eigenvalues, m["eigenvalues"])  # untouched
    assert got.result.mass_props["mass"] == m["mass"]
    again = got.rescaled((density, young, solved.poisson, solved.alpha, solved.beta), cfg)
    np.testing.assert_array_equal(again.result.freqs, want.freqs)
    # archived with the original solved material, the reference's loader reads the same summary back
    stored = ModalModel.from_bytes(got.to_bytes())
    np.testing.assert_array_equal(stored.result.eigenvalues, m["eigenvalues"])
    with pytest.raises(MeError):
        model.rescaled((density, young, solved.poisson + 0.01, solved.alpha, solved.beta), cfg)  # not exactly scalable -> nullopt


def test_inconsistent_modal_data_is_rejected():
    """Sizes the reference's nested vectors carry themselves but a flat table cannot: T60s vs Freqs, summary rows vs points."""
    with open(os.path.join(GOLDEN, "model_3.modal"), "rb") as f:
        data = bytearray(f.read())
    n = int(np.frombuffer(bytes(data[:4]), np.uint32)[0])
    t60_at = 4 + 4 * n
    assert int(np.frombuffer(bytes(data[t60_at:t60_at + 4]), np.uint32)[0]) == n
    short = bytearray(data)
    short[t60_at:t60_at + 4] = np.uint32(n - 1).tobytes()
    del short[t60_at + 4:t60_at + 8]  # one T60 fewer, everything after it intact
    with pytest.raises(MeError, match="T60s"):
        ModalModel.from_bytes(bytes(short))
