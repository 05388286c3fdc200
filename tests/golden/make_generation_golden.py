"""Generates tests/golden/generation/glue.npz: outputs of the UNMODIFIED reference functions of the modal generation job
(SampleSurfaceTriangles / CompactExcitationVertices / RelabelSampleTriangles cut out of src/audio/AudioSystem.cpp, and
BuildTetMeshData from src/mesh/Tets.cpp, the statements of RetuneModalObject, MonitorFrames and EffectiveModalMaterial, EstimateFundamentalFrequency whole; oracle/_ref, `make -C oracle ref`) on the seeded cases of oracle/generation.py, the
inputs stored beside them, and on BASELINE.json configs[0]'s IcoSphere (its surface, its tet mesh, the solver bench's ten
excitation vertices), where the big arrays are kept as SHA-256 digests.
Run: python tests/golden/make_generation_golden.py"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import generation as og  # noqa: E402

SEEDS = list(range(24)) + [100, 101, 102, 205]
TET_SEEDS = list(range(6))
RETUNE_SEEDS = list(range(28))
MONITOR_SEEDS = list(range(10))
MATERIAL_SEEDS = list(range(12))
RECORDING_SEEDS = list(range(20))
DIRECTION_SEEDS = list(range(40))
DYNAMICS_SEEDS = list(range(16))


def material_case(seed):
    rng = np.random.default_rng(7000 + seed)
    props = (float(rng.uniform(500, 9000)), float(rng.uniform(1e9, 3e11)), 0.3, 5.0, 1e-7)
    solved = (float(rng.uniform(500, 9000)), 2e11, 0.3, 5.0, 1e-7)
    return props, solved, float(rng.uniform(0.01, 50)), float(np.float32(rng.uniform(0.01, 50)))


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def icosphere():
    z = np.load(os.path.join(ROOT, "tests", "meshes", "icosphere_c1.npz"))
    n = len(z["surface"])
    return z, (np.arange(10) * n // 10).astype(np.uint32)  # ModalSolverBench.cpp:220-221


if __name__ == "__main__":
    assert og.have_ref(), "build oracle/_ref first (make -C oracle ref; needs /root/reference)"
    out = {}
    for seed in SEEDS:
        c = og.case(seed)
        tri = og.ref_sample_surface_triangles(c["triangles"], c["vertex_count"], c["vertices"])
        for k, v in c.items():
            out[f"s{seed}_{k}"] = np.asarray(v)
        out[f"s{seed}_sample_triangles"] = tri
        out[f"s{seed}_compact"] = og.ref_compact_excitation_vertices(c["vertices"], c["sample_point_of"])
        out[f"s{seed}_relabelled"] = og.ref_relabel_sample_triangles(tri, c["sample_point_of"])
    for seed in TET_SEEDS:
        t = og.tet_case(seed)
        positions, edges = og.ref_build_tet_mesh_data(t["points"], t["tets"], t["scale"])
        for k, v in t.items():
            out[f"t{seed}_{k}"] = v
        out[f"t{seed}_positions"], out[f"t{seed}_edges"] = positions, edges
    for seed in RETUNE_SEEDS:
        c = og.retune_case(seed)
        out[f"r{seed}_freqs"], out[f"r{seed}_t60s"] = og.ref_retune_modes(**c)
    for seed in MONITOR_SEEDS:
        c = og.monitor_case(seed)
        head, env = og.ref_monitor_frames(c["frames"][: c["split"]], c["sample_rate"], 0.0)
        tail, env = og.ref_monitor_frames(c["frames"][c["split"]:], c["sample_rate"], env)
        out[f"m{seed}_frames"], out[f"m{seed}_envelope"] = np.concatenate([head, tail]), np.float32(env)
    for seed in MATERIAL_SEEDS:
        out[f"e{seed}_material"] = np.array(og.ref_effective_modal_material(*material_case(seed))[:2])
    for seed in RECORDING_SEEDS:  # the reference's estimate over numpy's spectrum of the windowed segment; -1 = nullopt
        c = og.recording_case(seed)
        spectrum, n_real = og.impact_spectrum(**c)
        hz = og.ref_estimate_fundamental(spectrum, n_real, c["sample_rate"])
        out[f"f{seed}_hz"] = np.float32(-1 if hz is None else hz)
    out["d_directions"] = np.stack([og.ref_tilt_along_normal(*og.direction_case(seed)) for seed in DIRECTION_SEEDS])
    out["d_curvatures"] = np.array([og.ref_sphere_equivalent_curvature(rho, w) for rho, w in ((7850.0, 2.0), (1000.0, 0.0), (2700.0, 1e3), (750.0, 1e-4))])
    for seed in DYNAMICS_SEEDS:
        mass, inverse, arms = og.ref_contact_dynamics(**og.dynamics_case(seed))
        out[f"c{seed}_mass"], out[f"c{seed}_inverse"], out[f"c{seed}_arms"] = np.float64(mass), inverse, arms
    z, vertices = icosphere()
    tri = og.ref_sample_surface_triangles(z["triangles"], len(z["surface"]), vertices)
    out["ico_sample_triangles"] = tri
    scale = np.array([2.0, 0.5, 1.25], np.float32)
    positions, edges = og.ref_build_tet_mesh_data(z["points"], z["tets"], scale)
    out["ico_scale"], out["ico_tet_digests"], out["ico_edge_count"] = scale, np.array([digest(positions), digest(edges)]), np.array(len(edges))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "generation", "glue.npz"), **out)
    print(len(out), "arrays;", "icosphere sample triangles", len(tri) // 3, "tet edges", len(edges) // 2)
