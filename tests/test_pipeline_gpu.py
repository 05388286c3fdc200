"""BASELINE.json configs[0] end to end through the public API: the tetrahedralised IcoSphere (8,987 tets, steel) -> modal solve
on the GPU (up to 30 modes under 16 kHz) -> MeshEditorModalSolve JSON -> KHR_audio_rigid_bodies document -> synthesis bank -> a mallet strike built by
the strike front-end from the solve's own mass properties -> 1 s of audio at 48 kHz, against the reference bank (oracle) driven by
the same modes and the same event. Bar: 1e-5 of peak, both forms of the resonator bank."""
import json
import os

import numpy as np
import pytest

from oracle import resonator as orc

pytestmark = pytest.mark.gpu


def test_icosphere_solve_to_struck_audio():
    from mesheditor_b200 import ModalBank, solver_config
    from mesheditor_b200 import contact as mc
    from mesheditor_b200 import workloads as wl
    from mesheditor_b200.interchange import ModalModel, bank_modes, khr_modal_model, read_gltf_modal_models, write_gltf_modal_models

    points, tets, surface = wl.config1_mesh()
    excite = wl.bench_excitations(points)
    model = ModalModel.solve(points, tets, "Steel", excite, config=solver_config(num_modes=30, num_fem_modes=45))
    r = model.result
    assert r.status == 0 and 5 <= len(r.freqs) <= 30 and np.all(np.diff(r.freqs) >= 0)
    assert r.freqs[0] > 1000.0 and r.freqs[-1] <= 16000.0  # a 0.1 m steel ball rings high: only the modes under max_mode_freq are kept
    sphere_mass = 7850.0 * 4.0 / 3.0 * np.pi * 0.1**3  # materials::acoustic::Steel; the inscribed polyhedron is a little lighter
    assert 0.95 * sphere_mass < r.mass_props["mass"] < sphere_mass

    # interchange: solve JSON -> glTF document -> back -> the bank's view of the model
    khr = khr_modal_model(model.solve_json())
    khr.update(name="IcoSphere", material=dict(name="Steel", density=7850.0, youngsModulus=2.0e11, poissonRatio=0.29, alpha=5.0, beta=3.0e-8))
    (read_back,) = read_gltf_modal_models(json.loads(json.dumps(write_gltf_modal_models([khr]))))
    np.testing.assert_array_equal(read_back["frequencies"], r.freqs)
    modes = bank_modes(read_back)
    np.testing.assert_allclose(modes["t60s"], r.t60s, rtol=2e-7)
    np.testing.assert_array_equal(modes["shapes"], r.shapes)

    # strike front-end: a default steel mallet at 1 m/s on sample point 3, contact time from the solve's mass properties
    mp = r.mass_props
    inv_inertia = mc.inverse_inertia_tensor(mp["mass"], mp["inertia_diagonal"], mp["inertia_orientation"])
    arms = r.positions - np.asarray(mp["center_of_mass"], np.float32)
    dyn = mc.ContactDynamics(mp["mass"], inv_inertia, arms)
    direction = -arms[3] / np.linalg.norm(arms[3])

    def render(path, oracle_cls=None):
        bank = oracle_cls(48000.0, 1) if oracle_cls else ModalBank(48000.0, 0)
        bank.add_modes(modes)
        bank.install()
        if not oracle_cls:
            bank.set_render_path(path)
            radius = bank.object_layout(0)["RadiantRadius"]
        else:
            radius = float(bank.object_column("RadiantRadius")[0])
        ev = mc.make_strike_event(0, 3, 1.0, 1.0, direction, dynamics=dyn, elastic=mc.STEEL, curvature=10.0, displaced_volume=mp["mass"] / 7850.0, radiant_radius=radius, sample_rate=48000.0)
        assert 2e-5 < 1.0 / (ev.pulse_step * 48000.0) < 5e-2 and ev.accel_amp > 0 and ev.click_b0 != 0
        if oracle_cls:
            bank.enqueue(orc.Event(ev.kind, ev.object, ev.ex_pos, ev.jx, ev.jy, ev.jz, ev.pulse_step, ev.pulse_gamma, ev.accel_amp, ev.click_b0, ev.click_a1, ev.click_a2))
            return bank.render_blocks(94)
        out = bank.render_offline([ev], [0], 94 * 512, 512)
        assert (bank.stats()["tensor_windows"] > 0) == (path == 2)
        return out

    ref = render(0, orc.RefScene if orc.have_ref() else orc.PortBank)
    peak = float(np.abs(ref).max())
    assert peak > 0
    for path in (1, 2):
        assert float(np.abs(render(path) - ref).max()) <= 1e-5 * peak, path


def test_generation_job_writes_the_reference_modal_file():
    """The modal generation job after tetrahedralization (AudioSystem.cpp:830-862) over configs[0]: ten excitation vertices of the
    IcoSphere's surface, two more that repeat earlier ones (their positions reach the same tet point and merge) -> solve ->
    ModalModes::Vertices / Indices, eigen summary, TetMeshData -> `.modal` bytes. The reference's own archive (oracle/_ref) over the
    same fields writes the same bytes, decodes ours to the same model, and the file parses back to it."""
    from mesheditor_b200 import solver_config
    from mesheditor_b200.interchange import ModalModel, relabel_sample_triangles, sample_surface_triangles
    from oracle import generation as og
    from oracle import interchange as oi

    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "meshes", "icosphere_c1.npz"))
    surface, triangles = z["surface"], z["triangles"]
    vertices = np.concatenate([(np.arange(10) * len(surface) // 10), [0, 512]]).astype(np.uint32)  # 512 = 2 * 2562 // 10 again
    scale = (1.0, 1.0, 1.0)
    cfg = solver_config(num_modes=30, num_fem_modes=45)
    model = ModalModel.generate(z["points"], z["tets"], "Steel", surface, triangles, vertices, scale, cfg, tet_inputs_hash=0x1234_5678_9ABC)
    r = model.result
    assert model.status == 0 and not r.empty and len(r.positions) == 10
    np.testing.assert_array_equal(r.sample_point_of_excitation, list(range(10)) + [0, 2])
    np.testing.assert_array_equal(model.vertices, vertices[:10]), np.testing.assert_array_equal(model.solved_vertices, vertices)
    sample = sample_surface_triangles(triangles, len(surface), vertices)
    np.testing.assert_array_equal(model.indices, relabel_sample_triangles(sample, r.sample_point_of_excitation))
    np.testing.assert_array_equal(np.sort(model.indices.reshape(-1, 3), 1), np.sort(og.relabel_sample_triangles(og.sample_surface_triangles(triangles, len(surface), vertices), r.sample_point_of_excitation).reshape(-1, 3), 1))
    assert len(model.indices) == 3 * 16 and model.indices.max() == 9  # a closed surface over the ten sample points
    want_positions, want_edges = og.build_tet_mesh_data(z["points"], z["tets"], scale)
    np.testing.assert_array_equal(model.tet_positions, want_positions), np.testing.assert_array_equal(model.tet_edge_indices, want_edges)
    assert (model.solved_num_modes, model.solved_min_mode_freq, model.solved_max_mode_freq) == (30, 20.0, 16000.0)

    data = model.to_bytes()
    back = ModalModel.from_bytes(data)
    assert back.to_bytes() == data and back.tet_inputs_hash == 0x1234_5678_9ABC
    np.testing.assert_array_equal(back.indices, model.indices), np.testing.assert_array_equal(back.tet_edge_indices, model.tet_edge_indices)
    if oi.have_ref():
        mp, sm = r.mass_props, model.solved_material
        fields = dict(freqs=r.freqs, t60s=r.t60s, shapes=r.shapes, positions=r.positions, vertices=model.vertices, indices=model.indices, original_fundamental=np.float32(r.original_fundamental),
                      baked_scale=np.array(scale, np.float32), mass=mp["mass"], com=np.array(mp["center_of_mass"], np.float32), inertia=np.array(mp["inertia_diagonal"], np.float32),
                      quat_wxyz=np.array(mp["inertia_orientation"], np.float32), tet_positions=model.tet_positions, tet_edges=model.tet_edge_indices, eigenvalues=r.eigenvalues,
                      summary_shapes=r.summary_shapes, material=np.array([sm.density, sm.young_modulus, sm.poisson_ratio, sm.alpha, sm.beta]), min_freq=np.float32(20), max_freq=np.float32(16000),
                      num_modes=30, tet_hash=0x1234_5678_9ABC, solved_vertices=model.solved_vertices)
        assert oi.serialize(fields) == data and oi.parses_to(fields, data) == 1

