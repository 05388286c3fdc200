"""Generation-job glue (first slice) through the C ABI: the sample surface, ModalModes::Vertices and the display
TetMeshData the reference's modal generation job builds either side of mesh2modes (AudioSystem.cpp:838-862, Tets.cpp:268-293), and
the tuning front-end between a stored model and TuneModalObject (RetuneModalObject, AudioSystem.cpp:263-311),
against outputs of the UNMODIFIED reference functions - committed (tests/golden/generation/glue.npz) and live where
oracle/_ref is built - and against the restatement in oracle/generation.py. Host-only. Bar: bit-exact (index work; the float
positions are one rounded product each)."""
import hashlib
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_generation_golden import DIRECTION_SEEDS, DYNAMICS_SEEDS, MATERIAL_SEEDS, MONITOR_SEEDS, RECORDING_SEEDS, RETUNE_SEEDS, SEEDS, TET_SEEDS, icosphere, material_case  # noqa: E402

import mesheditor_b200 as me  # noqa: E402
from mesheditor_b200 import MeError  # noqa: E402
from mesheditor_b200 import interchange as mi  # noqa: E402
from oracle import generation as og  # noqa: E402

pytestmark = pytest.mark.usefixtures("built_lib")

GOLDEN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "generation", "glue.npz"))


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def triangle_sets(flat):
    return np.sort(np.asarray(flat).reshape(-1, 3), axis=1)


@pytest.mark.parametrize("seed", SEEDS)
def test_sample_surface_against_the_reference_outputs(seed):
    g = {k: GOLDEN[f"s{seed}_{k}"] for k in ("triangles", "vertex_count", "vertices", "sample_point_of", "sample_triangles", "compact", "relabelled")}
    case = og.case(seed)  # the seeded inputs are reproducible: what is stored is what the generator saw
    for k in ("triangles", "vertices", "sample_point_of"):
        np.testing.assert_array_equal(case[k], g[k])
    tri = mi.sample_surface_triangles(g["triangles"], int(g["vertex_count"]), g["vertices"])
    np.testing.assert_array_equal(tri, g["sample_triangles"])
    np.testing.assert_array_equal(mi.compact_excitation_vertices(g["vertices"], g["sample_point_of"]), g["compact"])
    np.testing.assert_array_equal(mi.relabel_sample_triangles(tri, g["sample_point_of"]), g["relabelled"])
    # the restatement: same point sets in the same order; a winding that the set was given with (the reference's survivor among
    # repeated sets is its unstable sort's choice, the restatement returns the first seen); compaction exact
    port = og.sample_surface_triangles(g["triangles"], int(g["vertex_count"]), g["vertices"])
    np.testing.assert_array_equal(triangle_sets(port), triangle_sets(tri))
    given = og.windings_of(og.collapsed_windings(g["triangles"], int(g["vertex_count"]), g["vertices"])) if len(g["vertices"]) >= 3 else {}
    for w in tri.reshape(-1, 3):
        assert tuple(int(x) for x in w) in given[tuple(sorted(int(x) for x in w))]
    np.testing.assert_array_equal(og.compact_excitation_vertices(g["vertices"], g["sample_point_of"]), g["compact"])
    np.testing.assert_array_equal(triangle_sets(og.relabel_sample_triangles(tri, g["sample_point_of"])), triangle_sets(g["relabelled"]))


@pytest.mark.parametrize("seed", TET_SEEDS)
def test_tet_mesh_data_against_the_reference_outputs(seed):
    g = {k: GOLDEN[f"t{seed}_{k}"] for k in ("points", "tets", "scale", "positions", "edges")}
    positions, edges = mi.build_tet_mesh_data(g["points"], g["tets"], g["scale"])
    np.testing.assert_array_equal(positions, g["positions"]), np.testing.assert_array_equal(edges, g["edges"])
    port_positions, port_edges = og.build_tet_mesh_data(g["points"], g["tets"], g["scale"])
    np.testing.assert_array_equal(port_positions, g["positions"]), np.testing.assert_array_equal(port_edges, g["edges"])
    pairs = edges.reshape(-1, 2).astype(np.int64)
    assert np.all(pairs[:, 0] < pairs[:, 1]) and np.all(np.diff(pairs[:, 0] * 2**32 + pairs[:, 1]) > 0)  # (low, high), strictly ascending


def test_icosphere_of_config1():
    """BASELINE.json configs[0]: 2,562 surface vertices, ten excitation vertices i*V/10 -> a closed sample surface of 2*10-4
    triangles; the tet mesh's 14,486 display edges."""
    z, vertices = icosphere()
    tri = mi.sample_surface_triangles(z["triangles"], len(z["surface"]), vertices)
    np.testing.assert_array_equal(tri, GOLDEN["ico_sample_triangles"])
    assert len(tri) == 3 * 16 and set(tri.tolist()) == set(range(10))
    # closed and consistently wound: every directed edge once, its reverse once
    edges = {(int(a), int(b)) for t in tri.reshape(-1, 3) for a, b in ((t[0], t[1]), (t[1], t[2]), (t[2], t[0]))}
    assert len(edges) == 48 and all((b, a) in edges for a, b in edges)
    positions, pairs = mi.build_tet_mesh_data(z["points"], z["tets"], GOLDEN["ico_scale"])
    assert [digest(positions), digest(pairs)] == GOLDEN["ico_tet_digests"].tolist() and len(pairs) == int(GOLDEN["ico_edge_count"]) == 2 * 14486
    # no merged excitation positions: vertices and triangles pass through the relabelling unchanged
    identity = np.arange(10, dtype=np.uint32)
    np.testing.assert_array_equal(mi.compact_excitation_vertices(vertices, identity), vertices)
    np.testing.assert_array_equal(mi.relabel_sample_triangles(tri, identity), tri)


@pytest.mark.skipif(not og.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_against_the_reference_functions_live():
    for seed in range(300, 420):
        c = og.case(seed)
        tri = mi.sample_surface_triangles(c["triangles"], c["vertex_count"], c["vertices"])
        np.testing.assert_array_equal(tri, og.ref_sample_surface_triangles(c["triangles"], c["vertex_count"], c["vertices"]))
        np.testing.assert_array_equal(mi.compact_excitation_vertices(c["vertices"], c["sample_point_of"]), og.ref_compact_excitation_vertices(c["vertices"], c["sample_point_of"]))
        np.testing.assert_array_equal(mi.relabel_sample_triangles(tri, c["sample_point_of"]), og.ref_relabel_sample_triangles(tri, c["sample_point_of"]))
    for seed in range(20, 40):
        t = og.tet_case(seed)
        ours, theirs = mi.build_tet_mesh_data(t["points"], t["tets"], t["scale"]), og.ref_build_tet_mesh_data(t["points"], t["tets"], t["scale"])
        np.testing.assert_array_equal(ours[0], theirs[0]), np.testing.assert_array_equal(ours[1], theirs[1])


@pytest.mark.parametrize("seed", RETUNE_SEEDS)
def test_retune_against_the_reference_outputs(seed):
    """RetuneModalObject's arithmetic (AudioSystem.cpp:271, 299-308): bit-exact against what the reference's own statements
    computed (committed) and compute (live), and against the float32 restatement."""
    c = og.retune_case(seed)
    want = GOLDEN[f"r{seed}_freqs"], GOLDEN[f"r{seed}_t60s"]
    got = me.retune_modes(c["freqs"], c["t60s"], me.retuning(c["scale"], c["fundamental"], c["t60_scale"], c["alpha"]))
    np.testing.assert_array_equal(got[0], want[0]), np.testing.assert_array_equal(got[1], want[1])
    port = og.retune_modes(**c)
    np.testing.assert_array_equal(port[0], want[0]), np.testing.assert_array_equal(port[1], want[1])
    assert np.all(got[1][c["t60s"] <= 0] == 0) and np.all(got[1][c["t60s"] > 0] > 0)  # the undamped sentinel survives, nothing else becomes it
    if og.have_ref():
        live = og.ref_retune_modes(**c)
        np.testing.assert_array_equal(got[0], live[0]), np.testing.assert_array_equal(got[1], live[1])


def test_retune_laws():
    f32 = np.float32
    freqs, t60s = np.array([200, 450, 900], f32), np.array([1.0, 0.5, 0.0], f32)
    same = me.retune_modes(freqs, t60s, me.retuning())
    np.testing.assert_array_equal(same[0], freqs)
    np.testing.assert_allclose(same[1], t60s, rtol=3e-7)  # ln1000 / (ln1000 / t60): two roundings
    up = me.retune_modes(freqs, t60s, me.retuning(fundamental=400.0))
    np.testing.assert_array_equal(up[0], freqs * f32(2))  # every mode follows the fundamental
    big = me.retune_modes(freqs, t60s, me.retuning(scale=2.0, alpha=0.0))
    np.testing.assert_array_equal(big[0], freqs / f32(2))  # twice the size: an octave down ...
    np.testing.assert_allclose(big[1][:2], t60s[:2] * 4, rtol=3e-7)  # ... and, without mass damping, four times the ring
    held = me.retune_modes(freqs, t60s, me.retuning(scale=2.0, alpha=2 * 6.9077554))  # all of mode 0's damping is alpha: size leaves it alone
    np.testing.assert_allclose(held[1][0], 1.0, rtol=1e-6)
    assert me.modal_out_gain(me.retuning(scale=2.0, modal_level=0.5, gain=3.0)) == 0.375
    assert me.uniform_scale_ratio((2, -2, 2)) == 2.0 and me.uniform_scale_ratio(None) == 1.0 and me.uniform_scale_ratio((1, 1, 1), (0, 0, 0)) == 1.0
    assert me.uniform_scale_ratio((1e-9, 0, 0)) == f32(0.001) and me.uniform_scale_ratio((1e9, 0, 0)) == 1000.0 and me.uniform_scale_ratio((3, 3, 3), (1, 2, 3)) == 1.5
    assert me.listener_gain(0.25) == 1.0 and me.listener_gain(4.0) == 0.25
    with pytest.raises(MeError):
        me.retune_modes(freqs, t60s, me.retuning(scale=0.0))


@pytest.mark.parametrize("seed", MONITOR_SEEDS)
def test_monitor_stage_against_the_reference_outputs(seed):
    """MonitorFrames (AudioSystem.cpp:1177-1189) in two calls with the envelope carried across: bit-exact against the reference's
    own loop (committed; live where oracle/_ref is built) and the float32 restatement."""
    c = og.monitor_case(seed)
    head, env = me.monitor_frames(c["frames"][: c["split"]], c["sample_rate"], 0.0)
    tail, env = me.monitor_frames(c["frames"][c["split"]:], c["sample_rate"], env)
    got = np.concatenate([head, tail])
    np.testing.assert_array_equal(got, GOLDEN[f"m{seed}_frames"])
    assert np.float32(env) == GOLDEN[f"m{seed}_envelope"]
    whole, env_whole = me.monitor_frames(c["frames"], c["sample_rate"])  # splitting the stream changes nothing
    np.testing.assert_array_equal(whole, got)
    assert env_whole == env and np.abs(got).max() <= 1.0  # never past the rail
    port, env_port = og.monitor_frames(c["frames"], c["sample_rate"])
    np.testing.assert_array_equal(port, got)
    assert np.float32(env_port) == np.float32(env)
    quiet = np.abs(c["frames"]) < 1e-3
    if og.have_ref():
        live, env_live = og.ref_monitor_frames(c["frames"], c["sample_rate"])
        np.testing.assert_array_equal(live, got)
        assert env_live == env
    assert quiet.sum() == 0 or np.all(np.abs(got[quiet]) <= 1e-3 / 20 + 1e-12)


def test_edit_loop_material_glue():
    """EffectiveModalMaterial and RescaledModes' pinned fundamental (AudioSystem.cpp:595-616), which feed me_rescale_modes: FP64,
    bit-exact against the reference's statements (committed and live) and the restatement."""
    from mesheditor_b200 import modal as mm

    for seed in MATERIAL_SEEDS:
        props, solved, solve_mass, body_mass = material_case(seed)
        got = mm.effective_modal_material(props, solved, solve_mass, body_mass)
        assert [got.density, got.young_modulus] == GOLDEN[f"e{seed}_material"].tolist()
        assert (got.density, got.young_modulus) == og.effective_modal_material(props, solved, solve_mass, body_mass)[:2]
        assert (got.poisson_ratio, got.alpha, got.beta) == props[2:]
        assert abs(got.young_modulus / got.density - props[1] / props[0]) <= 1e-15 * props[1] / props[0]  # E / rho kept: every frequency stays
        assert abs(got.density * solve_mass - solved[0] * body_mass) <= 1e-12 * solved[0] * body_mass  # the solve's mass at that density is the body's
        if og.have_ref():
            assert (got.density, got.young_modulus) == og.ref_effective_modal_material(props, solved, solve_mass, body_mass)[:2]
        for same in (mm.effective_modal_material(props, solved, solve_mass, 0.0), mm.effective_modal_material(props, solved, 0.0, body_mass), mm.effective_modal_material(props, (0.0,) + solved[1:], solve_mass, body_mass)):
            assert (same.density, same.young_modulus) == props[:2]  # the guard: no authoritative body, no solve mass, no solved density
    assert mm.pinned_fundamental([440.0, 900.0], 431.5) == 440.0  # matched to a recording at solve time: stays pinned
    assert mm.pinned_fundamental([431.5, 900.0], 431.5) is None and mm.pinned_fundamental([], 10.0) is None and mm.pinned_fundamental([440.0], 0.0) is None


@pytest.mark.parametrize("seed", RECORDING_SEEDS)
def test_fundamental_of_a_recorded_impact(seed):
    """ComputeFft + EstimateFundamentalFrequency (AudioSystem.cpp:492-560), the SolverConfig::FundamentalFreq a solve is matched to:
    the windowed segment's spectrum against numpy's transform (the reference uses FFTW; 1e-6 of the largest bin), the estimate
    against the reference's own function - committed answers, and live on both spectra - and the float32 restatement."""
    from mesheditor_b200 import modal as mm

    c = og.recording_case(seed)
    want = float(GOLDEN[f"f{seed}_hz"])
    want = None if want < 0 else want
    spectrum, n_real = mm.impact_spectrum(**c)
    port_spectrum, port_n = og.impact_spectrum(**c)
    assert n_real == port_n == c["sample_rate"] // 16 - 30 and len(spectrum) == n_real // 2 + 1
    assert np.abs(spectrum - port_spectrum).max() <= 1e-6 * np.abs(port_spectrum).max()
    assert mm.estimate_fundamental(**c) == want
    assert mm.estimate_fundamental_from_spectrum(port_spectrum, n_real, c["sample_rate"]) == want
    assert og.estimate_fundamental(port_spectrum, n_real, c["sample_rate"]) == want
    if seed % 5 == 4:
        assert want is None  # noise only: no prominent peak, the solve keeps its own fundamental
    else:
        assert want is not None and want >= 50
    if og.have_ref():
        assert og.ref_estimate_fundamental(port_spectrum, n_real, c["sample_rate"]) == want
        assert og.ref_estimate_fundamental(spectrum, n_real, c["sample_rate"]) == want


def test_strike_direction_and_colliding_curvature():
    """TiltAlongNormal / SphereEquivalentCurvature (AudioSystem.cpp:359-380), what TriggerModalStrike's callers hand it: bit-exact
    against the reference's own functions (committed and live)."""
    from mesheditor_b200 import contact as mc

    for seed in DIRECTION_SEEDS:
        n, joy = og.direction_case(seed)
        got = mc.tilt_along_normal(n, joy)
        np.testing.assert_array_equal(got, GOLDEN["d_directions"][seed])
        assert abs(float(np.linalg.norm(got.astype(np.float64))) - 1.0) < 2e-6  # a unit direction
        reach = min(float(np.hypot(*joy.astype(np.float64))), 1.0)
        assert abs(float(got.astype(np.float64) @ n.astype(np.float64)) - np.cos(reach * np.pi / 2)) < 2e-6  # tilted by radius * 90 degrees
        if og.have_ref():
            np.testing.assert_array_equal(got, og.ref_tilt_along_normal(n, joy))
    cases = ((7850.0, 2.0), (1000.0, 0.0), (2700.0, 1e3), (750.0, 1e-4))
    assert [mc.sphere_equivalent_curvature(rho, w) for rho, w in cases] == GOLDEN["d_curvatures"].tolist()
    radius = 0.05  # a 5 cm steel ball: curvature 1 / radius
    assert abs(mc.sphere_equivalent_curvature(7850.0, 1.0 / (7850.0 * 4.0 / 3.0 * np.pi * radius**3)) - 1.0 / radius) < 1e-12 / radius


@pytest.mark.parametrize("seed", DYNAMICS_SEEDS)
def test_contact_dynamics_against_the_reference_outputs(seed):
    """UpdateContactDynamics (ContactDynamics.cpp:19-46) past its registry lookups - the ContactDynamics a strike's contact time is
    estimated with, from a solve's mass properties and sample points: bit-exact against the reference's statements."""
    from mesheditor_b200 import contact as mc

    c = og.dynamics_case(seed)
    d = mc.contact_dynamics(dict(mass=c["mass"], center_of_mass=c["com"], inertia_diagonal=c["inertia_diagonal"], inertia_orientation=c["quat_wxyz"]), c["positions"], c["baked_scale"], c["mass_scale"])
    assert d.c.mass == float(GOLDEN[f"c{seed}_mass"]) == c["mass"] * c["mass_scale"]
    np.testing.assert_array_equal(np.array(list(d.c.inverse_inertia), np.float32), GOLDEN[f"c{seed}_inverse"])
    np.testing.assert_array_equal(d.arms, GOLDEN[f"c{seed}_arms"])
    if not c["baked_scale"].any():
        assert np.abs(d.arms).max() <= 1e-6 * np.abs(c["positions"] - c["com"]).max() * 1.01  # a zero baked scale is held at 1e-6
    if og.have_ref():
        mass, inverse, arms = og.ref_contact_dynamics(**c)
        assert d.c.mass == mass
        np.testing.assert_array_equal(np.array(list(d.c.inverse_inertia), np.float32), inverse), np.testing.assert_array_equal(d.arms, arms)


def test_desired_solve_vertices():
    """DesiredSolveVertices (AudioSystem.cpp:667-671): i * V / count, count clamped to 1..V; the 32-bit products wrap as the reference's."""
    cases = [(10, 2562), (1, 7), (0, 7), (50, 7), (7, 7), (3, 100000), (70000, 70000)]
    for requested, n in cases:
        got = mi.desired_solve_vertices(requested, n)
        count = min(max(requested, 1), n)
        want = ((np.arange(count, dtype=np.uint64) * n) % 2**32 // count).astype(np.uint32)
        np.testing.assert_array_equal(got, want)
        if og.have_ref():
            np.testing.assert_array_equal(got, og.ref_desired_solve_vertices(requested, n))
    np.testing.assert_array_equal(mi.desired_solve_vertices(10, 2562), icosphere()[1])  # the solver bench's rule on configs[0]
    assert len(set(mi.desired_solve_vertices(50, 7).tolist())) == 7  # capped: no vertex twice
    with pytest.raises(MeError):
        mi.desired_solve_vertices(3, 0)


def test_edge_cases():
    tri = np.array([0, 1, 2, 2, 1, 3], np.uint32)
    empty = np.zeros(0, np.uint32)
    assert len(mi.sample_surface_triangles(tri, 4, [0, 1])) == 0  # fewer than three excitation vertices
    assert len(mi.sample_surface_triangles(empty, 4, [0, 1, 2])) == 0  # no triangle
    np.testing.assert_array_equal(mi.sample_surface_triangles(tri, 4, [0, 1, 2, 3]), [0, 1, 2, 2, 1, 3])  # every vertex its own label
    np.testing.assert_array_equal(mi.sample_surface_triangles(tri, 4, [3, 2, 1]), [1, 2, 0])  # vertex 0 joins its first-listed neighbour 1: (2,1,1) collapses, (1,2,0) stays
    with pytest.raises(MeError):
        mi.sample_surface_triangles(tri, 3, [0, 1, 2])  # a triangle corner outside the vertices
    assert len(mi.relabel_sample_triangles(tri, empty)) == 0
    np.testing.assert_array_equal(mi.relabel_sample_triangles(tri, [0, 1, 1, 2]), [])  # (0,1,1) and (1,1,2) both collapse
    np.testing.assert_array_equal(mi.relabel_sample_triangles([0, 1, 2, 1, 2, 0, 3, 1, 0], [0, 1, 2, 2]), [0, 1, 2])  # three windings of one set: the first survives
    with pytest.raises(MeError):
        mi.relabel_sample_triangles(tri, [0, 1, 2])  # corner 3 has no sample point
    np.testing.assert_array_equal(mi.compact_excitation_vertices([7, 9, 4, 5], [0, 1, 0, 2]), [7, 9, 5])
    np.testing.assert_array_equal(mi.compact_excitation_vertices([7, 9, 4, 5], [0, 1]), [7, 9])  # the shorter of the two lists bounds the scan
    assert len(mi.compact_excitation_vertices(empty, empty)) == 0
    positions, edges = mi.build_tet_mesh_data(np.zeros((0, 3)), np.zeros((0, 4), np.uint32))
    assert positions.shape == (0, 3) and len(edges) == 0
    positions, edges = mi.build_tet_mesh_data([[0, 0, 0], [2, 0, 0], [0, 3, 0], [0, 0, 4]], [[3, 1, 2, 0]], (2.0, 3.0, 4.0))
    np.testing.assert_array_equal(positions, [[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]])
    np.testing.assert_array_equal(edges, [0, 1, 0, 2, 0, 3, 1, 2, 1, 3, 2, 3])
    with pytest.raises(MeError):
        mi.build_tet_mesh_data(np.zeros((3, 3)), [[0, 1, 2, 3]])
    out, env = me.monitor_frames([], 48000.0, 0.25)
    assert len(out) == 0 and env == 0.25
    out, env = me.monitor_frames([10.0, -40.0, 10.0], 48000.0)  # below the rail: plain scaling; above: held at it, and the release keeps the next sample down
    assert out[0] == 0.5 and out[1] == -1.0 and 0.25 < out[2] < 0.2501 and 1.99 < env < 2.0
    with pytest.raises(MeError):
        me.monitor_frames([1.0], 0.0)
    from mesheditor_b200 import modal as mm

    with pytest.raises(MeError):
        mm.impact_spectrum(np.zeros(2999, np.float32), 48000)  # shorter than sample_rate / 16 frames
    assert mm.estimate_fundamental(np.zeros(2999, np.float32), 48000) is None and mm.estimate_fundamental(np.zeros(3000, np.float32), 48000) is None
