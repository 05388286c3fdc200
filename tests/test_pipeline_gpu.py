"""BASELINE.json configs[0] end to end through the public API: the tetrahedralised IcoSphere (8,987 tets, steel) -> modal solve
on the GPU (up to 30 modes under 16 kHz) -> MeshEditorModalSolve JSON -> KHR_audio_rigid_bodies document -> synthesis bank -> a mallet strike built by
the strike front-end from the solve's own mass properties -> 1 s of audio at 48 kHz, against the reference bank (oracle) driven by
the same modes and the same event. Bar: 1e-5 of peak, both forms of the resonator bank."""
import json

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
