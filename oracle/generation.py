"""TEST INFRASTRUCTURE ONLY (oracle). The host-side steps either side of modal::mesh2modes and RenderModal that
src/audio/AudioSystem.cpp and src/mesh/Tets.cpp hold - the generation job's sample surface / excitation vertices / TetMeshData,
RetuneModalObject's arithmetic, MonitorFrames, EffectiveModalMaterial, EstimateFundamentalFrequency, TiltAlongNormal - two ways:
  * a plain-Python / numpy restatement of the reference's algorithm (small cases only), each function citing what it follows;
  * the UNMODIFIED reference code where oracle/_ref is built: libme_ref_glue.so (functions and statements cut out of
    src/audio/AudioSystem.cpp at build time - the file as a whole needs the editor's dependencies - oracle/ref_glue_driver.cpp)
    and libme_ref_tet.so (src/mesh/Tets.cpp: BuildTetMeshData, SimplifySurface; oracle/ref_tet_driver.cpp).
Parity pinned: the restatements are checked against the reference code live (tests/test_generation_cpu.py) and through
tests/golden/generation/glue.npz, which the reference code wrote (tests/golden/make_generation_golden.py)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GLUE_SO = os.path.join(HERE, "_ref", "libme_ref_glue.so")
TET_SO = os.path.join(HERE, "_ref", "libme_ref_tet.so")
UNLABELLED = 0xFFFFFFFF


def have_ref():
    if not (os.path.exists(GLUE_SO) and os.path.exists(TET_SO)):
        return False
    return hasattr(C.CDLL(TET_SO), "ref_build_tet_mesh_data")


# ---- restatement --------------------------------------------------------------------------------------------------------

def unique_sample_triangles(windings):
    """UniqueSampleTriangles (AudioSystem.cpp:675-695): drop triples that repeat a point; one winding per distinct point set,
    ordered by the sorted set. The reference keeps whichever duplicate its (unstable) sort leaves first; its comment states the
    intent - the winding first seen - which is what this returns. See windings_of() for what a test may demand of the rest."""
    first = {}
    for w in windings:
        w = tuple(int(x) for x in w)
        if w[0] == w[1] or w[1] == w[2] or w[0] == w[2]:
            continue
        first.setdefault(tuple(sorted(w)), w)
    return np.array([c for key in sorted(first) for c in first[key]], np.uint32)


def windings_of(windings):
    """Point set -> every winding it was given with (any of them is a legitimate survivor of the reference's unstable sort)."""
    out = {}
    for w in windings:
        w = tuple(int(x) for x in w)
        if len(set(w)) == 3:
            out.setdefault(tuple(sorted(w)), set()).add(w)
    return out


def surface_labels(triangle_indices, vertex_count, excitation_vertices):
    """The breadth-first labelling inside SampleSurfaceTriangles (AudioSystem.cpp:704-736)."""
    tri = np.asarray(triangle_indices, np.int64).reshape(-1)
    tri = tri[: len(tri) // 3 * 3].reshape(-1, 3)
    neighbours = [[] for _ in range(vertex_count)]
    for t in tri:
        for k in range(3):
            neighbours[t[k]] += [int(t[(k + 1) % 3]), int(t[(k + 2) % 3])]
    label = [UNLABELLED] * vertex_count
    queue = []
    for s, v in enumerate(excitation_vertices):
        if v < vertex_count and label[v] == UNLABELLED:
            label[v] = s
            queue.append(int(v))
    head = 0
    while head < len(queue):
        v = queue[head]
        head += 1
        for n in neighbours[v]:
            if label[n] == UNLABELLED:
                label[n] = label[v]
                queue.append(n)
    return label, tri


def collapsed_windings(triangle_indices, vertex_count, excitation_vertices):
    label, tri = surface_labels(triangle_indices, vertex_count, excitation_vertices)
    return [tuple(label[c] for c in t) for t in tri if all(label[c] != UNLABELLED for c in t)]


def sample_surface_triangles(triangle_indices, vertex_count, excitation_vertices):
    """SampleSurfaceTriangles (AudioSystem.cpp:701-746)."""
    if len(excitation_vertices) < 3 or len(triangle_indices) < 3:
        return np.zeros(0, np.uint32)
    return unique_sample_triangles(collapsed_windings(triangle_indices, vertex_count, excitation_vertices))


def compact_excitation_vertices(vertices, sample_point_of):
    """CompactExcitationVertices (AudioSystem.cpp:750-757)."""
    out = []
    for v, sp in zip(vertices, sample_point_of):
        if sp == len(out):
            out.append(int(v))
    return np.array(out, np.uint32)


def relabelled_windings(triangles, sample_point_of):
    tri = np.asarray(triangles, np.int64).reshape(-1)
    tri = tri[: len(tri) // 3 * 3].reshape(-1, 3)
    return [tuple(int(sample_point_of[c]) for c in t) for t in tri]


def relabel_sample_triangles(triangles, sample_point_of):
    """RelabelSampleTriangles (AudioSystem.cpp:761-769)."""
    if len(sample_point_of) == 0:
        return np.zeros(0, np.uint32)
    return unique_sample_triangles(relabelled_windings(triangles, sample_point_of))


def build_tet_mesh_data(points, tets, scale):
    """BuildTetMeshData (Tets.cpp:268-293): double points times the double reciprocal of the float scale, to float; the distinct
    edges (low << 32 | high) ascending."""
    inv = 1.0 / np.asarray(scale, np.float32).astype(np.float64)
    positions = (np.asarray(points, np.float64).reshape(-1, 3) * inv).astype(np.float32)
    t = np.asarray(tets, np.uint64).reshape(-1, 4)
    pairs = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]
    keys = np.concatenate([(np.minimum(t[:, a], t[:, b]) << np.uint64(32)) | np.maximum(t[:, a], t[:, b]) for a, b in pairs]) if len(t) else np.zeros(0, np.uint64)
    keys = np.unique(keys)
    edges = np.stack([keys >> np.uint64(32), keys & np.uint64(0xFFFFFFFF)], 1).astype(np.uint32).reshape(-1)
    return positions, edges


# ---- the reference's own functions (oracle/_ref) ---------------------------------------------------------------------------

def _u32(a):
    return np.ascontiguousarray(a, np.uint32).reshape(-1)


def _glue():
    L = C.CDLL(GLUE_SO)
    for name, args in {"ref_sample_surface_triangles": [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32], "ref_compact_excitation_vertices": [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32],
                       "ref_relabel_sample_triangles": [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]}.items():
        getattr(L, name).argtypes, getattr(L, name).restype = args, C.c_uint32
    L.ref_glue_copy.argtypes, L.ref_glue_copy.restype = [C.c_void_p], None
    return L


def _glue_result(L, n):
    out = np.zeros(n, np.uint32)
    if n:
        L.ref_glue_copy(out.ctypes.data)
    return out


def ref_sample_surface_triangles(triangle_indices, vertex_count, excitation_vertices):
    L, tri, ex = _glue(), _u32(triangle_indices), _u32(excitation_vertices)
    return _glue_result(L, L.ref_sample_surface_triangles(tri.ctypes.data, len(tri), vertex_count, ex.ctypes.data, len(ex)))


def ref_compact_excitation_vertices(vertices, sample_point_of):
    L, v, sp = _glue(), _u32(vertices), _u32(sample_point_of)
    return _glue_result(L, L.ref_compact_excitation_vertices(v.ctypes.data, len(v), sp.ctypes.data, len(sp)))


def ref_relabel_sample_triangles(triangles, sample_point_of):
    L, tri, sp = _glue(), _u32(triangles), _u32(sample_point_of)
    return _glue_result(L, L.ref_relabel_sample_triangles(tri.ctypes.data, len(tri), sp.ctypes.data, len(sp)))


def ref_build_tet_mesh_data(points, tets, scale):
    L = C.CDLL(TET_SO)
    L.ref_build_tet_mesh_data.argtypes, L.ref_build_tet_mesh_data.restype = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p], C.c_uint32
    L.ref_tet_data_copy.argtypes, L.ref_tet_data_copy.restype = [C.c_void_p, C.c_void_p], None
    pts, tt, sc = np.ascontiguousarray(points, np.float64).reshape(-1, 3), np.ascontiguousarray(tets, np.uint32).reshape(-1, 4), np.ascontiguousarray(scale, np.float32)
    n = L.ref_build_tet_mesh_data(pts.ctypes.data, len(pts), tt.ctypes.data, len(tt), sc.ctypes.data)
    positions, edges = np.zeros((max(len(pts), 1), 3), np.float32), np.zeros(max(n, 1), np.uint32)
    L.ref_tet_data_copy(positions.ctypes.data, edges.ctypes.data)
    return positions[: len(pts)], edges[:n]


def ref_simplify_surface(positions, triangle_indices, ratio):
    """SimplifySurface (Tets.cpp:249-262): fixture generation only."""
    L = C.CDLL(TET_SO)
    L.ref_simplify_surface.argtypes, L.ref_simplify_surface.restype = [C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32), C.c_float], C.c_uint32
    pos, tri = np.ascontiguousarray(positions, np.float32).reshape(-1, 3).copy(), _u32(triangle_indices).copy()
    n_tri = C.c_uint32(len(tri))
    n_pos = L.ref_simplify_surface(pos.ctypes.data, len(pos), tri.ctypes.data, C.byref(n_tri), ratio)
    return pos[:n_pos].copy(), tri[: n_tri.value].copy()


# ---- seeded cases (shared by the golden generator and the tests) ---------------------------------------------------------------

def case(seed):
    """A closed triangulated surface with a random excitation set, and a plausible SamplePointOfExcitation that merges some of them."""
    rng = np.random.default_rng(seed)
    kind = seed % 3
    if kind == 0:  # torus grid, consistently wound
        nu, nv = int(rng.integers(5, 14)), int(rng.integers(4, 11))
        idx = lambda i, j: (i % nu) * nv + (j % nv)  # noqa: E731
        tri = [c for i in range(nu) for j in range(nv) for c in (idx(i, j), idx(i + 1, j), idx(i + 1, j + 1), idx(i, j), idx(i + 1, j + 1), idx(i, j + 1))]
        n_vertices = nu * nv
    elif kind == 1:  # two disjoint shells (tetrahedron boundaries subdivided by a centre fan), one may hold no excitation vertex
        def shell(base):
            faces = [(0, 2, 1), (0, 1, 3), (1, 2, 3), (0, 3, 2)]
            out = []
            for f, (a, b, c) in enumerate(faces):
                m = 4 + f
                out += [a, b, m, b, c, m, c, a, m]
            return [base + x for x in out]
        tri, n_vertices = shell(0) + shell(8), 16 + int(rng.integers(0, 3))  # trailing vertices no triangle uses
    else:  # random triangle soup with repeats and degenerate triangles
        n_vertices = int(rng.integers(6, 40))
        tri = rng.integers(0, n_vertices, 3 * int(rng.integers(4, 120))).tolist()
    n_ex = int(rng.integers(2, max(4, n_vertices // 2)))
    if kind == 1 and seed % 2:
        vertices = rng.choice(8, size=min(n_ex, 6), replace=False)  # second shell left without excitation vertices
    else:
        vertices = rng.choice(n_vertices, size=min(n_ex, n_vertices), replace=False)
    if seed % 5 == 0:
        vertices = np.concatenate([vertices, vertices[:2], [n_vertices + 3]])  # repeats and one out of range
    # sample points numbered by first appearance, some excitation positions merged into an earlier one
    sp, count = [], 0
    for i in range(len(vertices)):
        if i and rng.random() < 0.25:
            sp.append(int(rng.integers(0, count)))
        else:
            sp.append(count)
            count += 1
    return dict(triangles=np.asarray(tri, np.uint32), vertex_count=n_vertices, vertices=np.asarray(vertices, np.uint32), sample_point_of=np.asarray(sp, np.uint32))


def tet_case(seed):
    """A Kuhn block (6 tets per cell) with jittered points and an anisotropic node scale."""
    rng = np.random.default_rng(1000 + seed)
    n = [int(x) for x in rng.integers(1, 5, 3)]
    grid = np.stack(np.meshgrid(*[np.arange(k + 1) for k in n], indexing="ij"), -1).reshape(-1, 3)
    points = grid * rng.uniform(0.01, 0.3, 3) + rng.normal(0, 1e-3, grid.shape)
    vid = lambda i, j, k: (i * (n[1] + 1) + j) * (n[2] + 1) + k  # noqa: E731
    tets = []
    for i in range(n[0]):
        for j in range(n[1]):
            for k in range(n[2]):
                c = [vid(i + a, j + b, k + d) for a in (0, 1) for b in (0, 1) for d in (0, 1)]
                for path in ((1, 3), (1, 5), (2, 3), (2, 6), (4, 5), (4, 6)):
                    tets.append([c[0], c[path[0]], c[path[1]], c[7]])
    scale = rng.uniform(0.3, 3.0, 3).astype(np.float32) if seed % 2 else np.ones(3, np.float32)
    return dict(points=np.ascontiguousarray(points, np.float64), tets=np.asarray(tets, np.uint32), scale=scale)


# ---- tuning front-end (RetuneModalObject's arithmetic, AudioSystem.cpp:263-311) ----------------------------------------------------

def retune_modes(freqs, t60s, scale=1.0, fundamental=0.0, t60_scale=1.0, alpha=None):
    """AudioSystem.cpp:271 and :299-308 in float32, one rounding per operation as the reference's float expressions have."""
    f32 = np.float32
    freqs, t60s, scale = np.asarray(freqs, f32), np.asarray(t60s, f32), f32(scale)
    ln1000 = f32(3) * f32(np.log(10.0))
    ratio = (f32(fundamental) / freqs[0] if fundamental > 0 and freqs[0] > 0 else f32(1)) / scale
    out_f, out_t = freqs * ratio, np.zeros(len(t60s), f32)
    half = f32(np.float64(alpha) / 2) if alpha is not None else None
    for k, t60 in enumerate(t60s):
        if t60 <= 0:
            continue
        d = ln1000 / t60
        if alpha is not None:
            d = half + (d - half) / (scale * scale)
        out_t[k] = f32(t60_scale) * ln1000 / max(d, f32(1e-9))
    return out_f, out_t


def ref_retune_modes(freqs, t60s, scale=1.0, fundamental=0.0, t60_scale=1.0, alpha=None):
    """The reference's own statements (cut out of RetuneModalObject at build time, oracle/ref_glue_driver.cpp)."""
    L = C.CDLL(GLUE_SO)
    L.ref_retune.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
    L.ref_retune.restype = None
    f, t = np.ascontiguousarray(freqs, np.float32), np.ascontiguousarray(t60s, np.float32)
    out_f, out_t = np.zeros(len(f), np.float32), np.zeros(len(f), np.float32)
    L.ref_retune(f.ctypes.data, t.ctypes.data, len(f), scale, fundamental, t60_scale, int(alpha is not None), 0.0 if alpha is None else alpha, out_f.ctypes.data, out_t.ctypes.data)
    return out_f, out_t


def retune_case(seed):
    rng = np.random.default_rng(5000 + seed)
    n = int(rng.integers(1, 40))
    freqs = np.sort(rng.uniform(30, 15000, n)).astype(np.float32)
    t60s = rng.uniform(1e-3, 8.0, n).astype(np.float32)
    t60s[rng.random(n) < 0.15] = 0.0  # undamped sentinel
    if seed % 7 == 0:
        freqs[0] = 0.0  # no usable fundamental to retarget
    return dict(freqs=freqs, t60s=t60s, scale=float(np.float32(rng.choice([1.0, 0.001, 1000.0, rng.uniform(0.05, 20.0)]))), fundamental=float(np.float32(rng.choice([0.0, -1.0, rng.uniform(40, 2000)]))),
                t60_scale=float(np.float32(rng.uniform(0.1, 3.0))), alpha=None if seed % 3 == 0 else float(rng.choice([0.0, 0.5, 5.0, 60.0])))


# ---- monitor stage (MonitorFrames, AudioSystem.cpp:1177-1189) ---------------------------------------------------------------------

def monitor_frames(frames, sample_rate=48000.0, envelope=0.0):
    """Float32 restatement: x = frame / 20 Pa; envelope = max(|x|, envelope * release); frame = x / envelope above the rail."""
    f32 = np.float32
    release = f32(np.exp(np.float64(f32(-1.0) / (f32(0.1) * f32(sample_rate)))))  # expf: the float argument's exponential, rounded once
    out, env = np.array(frames, f32).reshape(-1), f32(envelope)
    for i in range(len(out)):
        x = out[i] / f32(20.0)
        env = max(abs(x), f32(env * release))
        out[i] = x / env if env > 1 else x
    return out, float(env)


def ref_monitor_frames(frames, sample_rate=48000.0, envelope=0.0):
    """The reference's own loop (cut out of MonitorFrames at build time, oracle/ref_glue_driver.cpp)."""
    L = C.CDLL(GLUE_SO)
    L.ref_monitor_frames.argtypes, L.ref_monitor_frames.restype = [C.c_void_p, C.c_uint64, C.c_float, C.c_float], C.c_float
    out = np.array(frames, np.float32).reshape(-1)
    env = L.ref_monitor_frames(out.ctypes.data, len(out), sample_rate, envelope)
    return out, float(env)


def monitor_case(seed):
    """Decaying bursts in Pa, some far above the 20 Pa rail, some silent stretches; rendered in two calls to carry the envelope."""
    rng = np.random.default_rng(9000 + seed)
    n = int(rng.integers(200, 3000))
    t = np.arange(n)
    x = np.zeros(n)
    for _ in range(int(rng.integers(1, 6))):
        start, amp = int(rng.integers(0, n)), float(rng.choice([0.5, 15.0, 25.0, 400.0]))
        x[start:] += amp * np.exp(-(t[start:] - start) / rng.uniform(20, 800)) * np.sin(0.05 * rng.uniform(1, 40) * (t[start:] - start))
    return dict(frames=x.astype(np.float32), sample_rate=float(rng.choice([44100.0, 48000.0, 96000.0])), split=int(rng.integers(1, n)))


# ---- edit-loop material glue (EffectiveModalMaterial / RescaledModes, AudioSystem.cpp:595-616) ----------------------------------------

def effective_modal_material(props, solved, solve_mass, body_mass=None):
    """AudioSystem.cpp:595-601. props / solved: (density, young, poisson, alpha, beta); body_mass None: no authoritative dynamic body."""
    props = [float(x) for x in props]
    if body_mass is None or not body_mass > 0 or solve_mass <= 0 or solved[0] <= 0 or props[0] <= 0:
        return tuple(props)
    rho = float(solved[0]) * float(np.float32(body_mass)) / solve_mass
    props[1] *= rho / props[0]
    props[0] = rho
    return tuple(props)


def ref_effective_modal_material(props, solved, solve_mass, body_mass):
    """The reference's own statements past the guard (oracle/ref_glue_driver.cpp)."""
    L = C.CDLL(GLUE_SO)
    L.ref_effective_material.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double, C.c_double, C.c_int, C.c_float]
    L.ref_effective_material.restype = None
    d, e = C.c_double(props[0]), C.c_double(props[1])
    L.ref_effective_material(C.byref(d), C.byref(e), solved[0], solve_mass, 1, body_mass)
    return (d.value, e.value) + tuple(float(x) for x in props[2:])


# ---- impact spectrum analysis (AudioSystem.cpp:492-560) ------------------------------------------------------------------------------

def impact_spectrum(frames, sample_rate=48000):
    """ComputeFft (AudioSystem.cpp:553-558) with numpy's transform in place of FFTW: frames 30 .. sample_rate/16 under the float32
    Blackman-Harris window (CreateBlackmanHarris, :494-513), spectrum rounded to complex64. Returns (spectrum, n_real)."""
    f32 = np.float32
    n = sample_rate // 16 - 30
    i, coeff = np.arange(n, dtype=np.int64), [f32(0.35875), f32(-0.48829), f32(0.14128), f32(-0.01168)]
    w = np.zeros(n, f32)
    for j, c in enumerate(coeff):
        w = (w + c * np.cos(np.pi * ((2 * i * j).astype(f32) / f32(n)).astype(np.float64)).astype(f32)).astype(f32)
    x = (w * np.asarray(frames, f32)[30:30 + n]).astype(f32)
    return np.fft.rfft(x.astype(np.float64)).astype(np.complex64), n


def estimate_fundamental(spectrum, n_real, sample_rate=48000):
    """EstimateFundamentalFrequency (AudioSystem.cpp:522-550) in float32."""
    f32 = np.float32
    s = np.asarray(spectrum, np.complex64)
    power = (s.real * s.real + s.imag * s.imag).astype(f32)
    db = (f32(10) * np.log10(np.maximum(power, f32(1e-20)).astype(np.float64))).astype(f32)
    bins, reach = len(db), 15
    upper = np.sort(db[bins // 2:])
    threshold = f32(upper[len(upper) // 2] + f32(15))
    for k in range(max(50 * n_real // sample_rate, reach), bins - reach):
        if db[k] <= db[k - 1] or db[k] <= db[k + 1] or db[k] < threshold:
            continue
        total = f32(0)
        for v in db[k - reach:k + reach + 1]:
            total = f32(total + v)
        if f32(db[k] - f32(total / f32(2 * reach + 1))) >= 10:
            return float(k * sample_rate // n_real)
    return None


def ref_estimate_fundamental(spectrum, n_real, sample_rate=48000):
    """The reference's own function (cut whole out of AudioSystem.cpp at build time) over the caller's spectrum."""
    L = C.CDLL(GLUE_SO)
    L.ref_estimate_fundamental.argtypes, L.ref_estimate_fundamental.restype = [C.c_void_p, C.c_uint64, C.c_uint32, C.POINTER(C.c_float)], C.c_int
    s, hz = np.ascontiguousarray(spectrum, np.complex64), C.c_float()
    return hz.value if L.ref_estimate_fundamental(s.ctypes.data, n_real, sample_rate, C.byref(hz)) else None


def recording_case(seed):
    """A struck object's recording: decaying partials over a noise floor; some cases hold nothing but noise."""
    rng = np.random.default_rng(11000 + seed)
    sample_rate = int(rng.choice([44100, 48000, 96000]))
    n = sample_rate // 16 + int(rng.integers(0, 500))
    t = np.arange(n) / sample_rate
    x = rng.normal(0, 1e-4, n)
    if seed % 5 != 4:
        f0 = float(rng.uniform(80, 2500))
        for h, amp in zip([1.0, 2.76, 5.4, 8.93][: int(rng.integers(1, 5))], [1.0, 0.5, 0.3, 0.2]):
            x += amp * np.exp(-t * rng.uniform(3, 40)) * np.sin(2 * np.pi * f0 * h * t + rng.uniform(0, 6.28))
    return dict(frames=x.astype(np.float32), sample_rate=sample_rate)


# ---- strike direction (AudioSystem.cpp:359-380) ----------------------------------------------------------------------------------------

def ref_tilt_along_normal(normal, joystick):
    L = C.CDLL(GLUE_SO)
    L.ref_tilt_along_normal.argtypes, L.ref_tilt_along_normal.restype = [C.c_void_p, C.c_void_p, C.c_void_p], None
    n, j, out = np.ascontiguousarray(normal, np.float32), np.ascontiguousarray(joystick, np.float32), np.zeros(3, np.float32)
    L.ref_tilt_along_normal(n.ctypes.data, j.ctypes.data, out.ctypes.data)
    return out


def ref_sphere_equivalent_curvature(density, inv_mass):
    L = C.CDLL(GLUE_SO)
    L.ref_sphere_equivalent_curvature.argtypes, L.ref_sphere_equivalent_curvature.restype = [C.c_double, C.c_double], C.c_double
    return L.ref_sphere_equivalent_curvature(density, inv_mass)


def direction_case(seed):
    rng = np.random.default_rng(13000 + seed)
    n = rng.normal(size=3)
    if seed % 4 == 0:
        n = np.array([0.0, 0.0, -1.0 if seed % 8 else 1.0]) + rng.normal(0, 1e-3, 3)  # near the poles of the tangent-frame construction
    n = (n / np.linalg.norm(n)).astype(np.float32)
    joy = (rng.uniform(-1.2, 1.2, 2) if seed % 5 else np.zeros(2)).astype(np.float32)  # past the rim clamps; the centre keeps the normal
    return n, joy


# ---- contact dynamics (UpdateContactDynamics, src/audio/ContactDynamics.cpp:19-46) ----------------------------------------------------

AUDIO_SO = os.path.join(HERE, "_ref", "libme_ref_audio.so")


def ref_contact_dynamics(mass, com, inertia_diagonal, quat_wxyz, mass_scale, positions, baked_scale):
    """The reference's own statements past the registry lookups (oracle/ref_audio_driver.cpp) -> (mass, inverse inertia[9], arms)."""
    L = C.CDLL(AUDIO_SO)
    L.ref_contact_dynamics.argtypes = [C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(C.c_double), C.c_void_p, C.c_void_p]
    L.ref_contact_dynamics.restype = C.c_uint32
    f = lambda a: np.ascontiguousarray(a, np.float32)  # noqa: E731
    com, diag, quat, pos, baked = f(com), f(inertia_diagonal), f(quat_wxyz), f(positions).reshape(-1, 3), f(baked_scale)
    out_mass, inverse, arms = C.c_double(), np.zeros(9, np.float32), np.zeros_like(pos)
    L.ref_contact_dynamics(mass, com.ctypes.data, diag.ctypes.data, quat.ctypes.data, mass_scale, pos.ctypes.data, len(pos), baked.ctypes.data, C.byref(out_mass), inverse.ctypes.data, arms.ctypes.data)
    return out_mass.value, inverse, arms


def dynamics_case(seed):
    rng = np.random.default_rng(15000 + seed)
    q = rng.normal(size=4)
    return dict(mass=float(rng.uniform(0.01, 40)), com=rng.normal(0, 0.05, 3).astype(np.float32), inertia_diagonal=rng.uniform(1e-5, 2, 3).astype(np.float32),
                quat_wxyz=(q / np.linalg.norm(q)).astype(np.float32), mass_scale=1.0 if seed % 3 == 0 else float(rng.uniform(0.2, 4)),
                positions=rng.normal(0, 0.2, (int(rng.integers(1, 12)), 3)).astype(np.float32), baked_scale=(rng.uniform(-3, 3, 3) if seed % 4 else np.zeros(3)).astype(np.float32))


def ref_desired_solve_vertices(requested, num_vertices):
    """DesiredSolveVertices (AudioSystem.cpp:667-671) past its copied-vertices branch, the reference's own statements."""
    L = _glue()
    L.ref_desired_solve_vertices.argtypes, L.ref_desired_solve_vertices.restype = [C.c_uint32, C.c_uint32], C.c_uint32
    return _glue_result(L, L.ref_desired_solve_vertices(requested, num_vertices))
