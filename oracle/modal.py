"""CPU restatement of the reference's modal ANALYSIS path (modal::mesh2modes). TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module;
the product (mesheditor_b200, libme_modal.so) never does.

Every function cites the reference lines it follows (paths relative to /root/reference). The reference solver
itself cannot be built in this image (it needs Eigen and Apple Accelerate), so this restatement is
pinned by the reference's golden modal models instead (tests/golden/*.npz, produced by the macOS reference binary
and embedded in glTF_PhysicalAudio/samples; tests/test_oracle_golden.py): parity is PINNED, not unpinned.

Arithmetic: numpy float64 throughout, element tables from exact barycentric integrals. The eigensolve is scipy's
ARPACK shift-invert eigsh (any converged solver gives the eigenvalues to 1e-8).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

EDGE_CORNERS = np.array([[0, 1], [0, 2], [0, 3], [1, 2], [1, 3], [2, 3]], dtype=np.int64)  # mesh2modes.cpp:200


@dataclass
class Material:
    """AcousticMaterialProperties (src/audio/AcousticMaterialProperties.h:6-16)."""

    density: float
    young: float
    poisson: float
    alpha: float
    beta: float

    @property
    def lam(self):
        return (self.poisson * self.young) / ((1 + self.poisson) * (1 - 2 * self.poisson))

    @property
    def mu(self):
        return self.young / (2 * (1 + self.poisson))


# materials::acoustic (src/audio/AcousticMaterial.h:33-40)
MATERIALS = {
    "Ceramic": Material(2700, 7.2e10, 0.19, 6, 1e-7), "Glass": Material(2600, 6.2e10, 0.20, 1, 1e-7),
    "Wood": Material(750, 1.1e10, 0.25, 60, 2e-6), "Plastic": Material(1070, 1.4e9, 0.35, 30, 1e-6),
    "Iron": Material(8000, 2.1e11, 0.28, 5, 1e-7), "Polycarbonate": Material(1190, 2.4e9, 0.37, 0.5, 4e-7),
    "Steel": Material(7850, 2.0e11, 0.29, 5, 3e-8),
}


@dataclass
class SolverConfig:
    """modal::SolverConfig (src/audio/mesh2modes.h:17-26)."""

    min_mode_freq: float = 20.0
    max_mode_freq: float = 16000.0
    num_modes: int = 30
    num_fem_modes: int = 45
    tolerance: float = 1e-8
    warm_tolerance: float = 1e-4
    max_restarts: int = 100
    fundamental_freq: float | None = None


@dataclass
class Modes:
    """ModalModes (src/audio/ModalModes.h:7-20) plus the solve's raw eigenvalues."""

    freqs: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float32))
    t60s: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float32))
    shapes: np.ndarray = field(default_factory=lambda: np.zeros((0, 0, 3), np.float32))  # [point][mode][3]
    positions: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.float32))
    original_fundamental: float = 0.0
    lowest_mode: int = 0


# ------------------------------------------------------------------------------------------------ geometry

def filter_degenerate(points, tets):
    """FilterDegenerate (mesh2modes.cpp:42-60): keep tets with |det| > 1e-12 * lmax^3."""
    p = points[tets]  # [T,4,3]
    r0, r1, r2 = p[:, 1] - p[:, 0], p[:, 2] - p[:, 0], p[:, 3] - p[:, 0]
    det = np.abs(np.einsum("ij,ij->i", r0, np.cross(r1, r2)))
    lmax_sq = np.zeros(len(tets))
    for i in range(4):
        for j in range(i + 1, 4):
            d = p[:, i] - p[:, j]
            lmax_sq = np.maximum(lmax_sq, np.einsum("ij,ij->i", d, d))
    keep = det > 1e-12 * lmax_sq * np.sqrt(lmax_sq)
    return tets[keep]


def tet_determinant(p):
    """GetTetDeterminant (mesh2modes.cpp:64-66): dot(d - a, cross(b - a, c - a)) for p[T,4,3]."""
    a, b, c, d = p[:, 0], p[:, 1], p[:, 2], p[:, 3]
    return np.einsum("ij,ij->i", d - a, np.cross(b - a, c - a))


def mass_properties(points, tets, density, scale=(1.0, 1.0, 1.0), length_to_si=1.0):
    """ComputeMassProperties (mesh2modes.cpp:73-126): lumped quarter volumes -> mass, COM, principal inertia."""
    pos = points * (1.0 / np.asarray(scale, np.float64))
    quarter = np.abs(tet_determinant(pos[tets])) * float(np.float32(1.0) / np.float32(6.0)) * 0.25  # (1.f/6.f) is a float constant (:68)
    vol = np.zeros(len(points))
    np.add.at(vol, tets.ravel(), np.repeat(quarter, 4))
    total = vol.sum()
    if total <= 0:
        return dict(mass=0.0, com=np.zeros(3, np.float32), inertia=np.zeros(3, np.float32))
    com = (vol[:, None] * pos).sum(0) / total
    r = pos - com
    rr = np.einsum("ij,ij->i", r, r)
    inertia = np.einsum("i,ijk->jk", vol, rr[:, None, None] * np.eye(3) - r[:, :, None] * r[:, None, :])
    s = length_to_si
    inertia = inertia * density * s ** 5
    evals = np.linalg.eigvalsh(inertia)
    return dict(mass=density * total * s ** 3, com=com.astype(np.float32), inertia=evals.astype(np.float32))


def element_bases(points, tets):
    """ComputeElementBases (mesh2modes.cpp:137-165): volume and the 4 linear basis gradients of every tet.
    Phig[k] = grad(lambda_k): rows 1..3 of the inverse of [1 x y z], which is what the
    cofactor loop of the reference evaluates."""
    p = points[tets]
    det = tet_determinant(p)
    volume = np.abs(det / 6.0)
    a = np.concatenate([np.ones((len(tets), 4, 1)), p], axis=2)  # rows = corners, [1 x y z]
    inv = np.linalg.inv(a)  # inv[:, 1:4, k] = gradient of lambda_k
    phig = np.transpose(inv[:, 1:4, :], (0, 2, 1)).copy()  # [T, k, xyz]
    return volume, phig


# ------------------------------------------------------------------------------------------------ element tables

def _poly_mul(a, b):
    """Multiply (mesh2modes.cpp:177-186): polynomials in barycentric coordinates as lists of (coeff, exponents)."""
    return [(ca * cb, tuple(x + y for x, y in zip(ea, eb))) for ca, ea in a for cb, eb in b]


def _unit_integral(p):
    """UnitIntegral (mesh2modes.cpp:188-196): int l^e dV / V = 6 prod(e!) / (sum(e)+3)!"""
    f = math.factorial
    return sum(c * 6 * f(e[0]) * f(e[1]) * f(e[2]) * f(e[3]) / f(sum(e) + 3) for c, e in p)


def _unit(i):
    return tuple(1 if k == i else 0 for k in range(4))


def quad_basis():
    """GetQuadBasis (mesh2modes.cpp:209-237): Mass[10][10], Grad[10][4][10][4] of the 10-node tet."""
    n = [None] * 10
    dn = [[[] for _ in range(4)] for _ in range(10)]
    for i in range(4):
        n[i] = [(2.0, tuple(2 * x for x in _unit(i))), (-1.0, _unit(i))]
        dn[i][i] = [(4.0, _unit(i)), (-1.0, (0, 0, 0, 0))]
    for e, (i, j) in enumerate(EDGE_CORNERS):
        n[4 + e] = [(4.0, tuple(x + y for x, y in zip(_unit(i), _unit(j))))]
        dn[4 + e][i] = [(4.0, _unit(j))]
        dn[4 + e][j] = [(4.0, _unit(i))]
    mass = np.zeros((10, 10))
    grad = np.zeros((10, 4, 10, 4))
    for a in range(10):
        for c in range(10):
            mass[a, c] = _unit_integral(_poly_mul(n[a], n[c]))
            for k in range(4):
                for l in range(4):
                    if dn[a][k] and dn[c][l]:
                        grad[a, k, c, l] = _unit_integral(_poly_mul(dn[a][k], dn[c][l]))
    return mass, grad


def linear_basis():
    """The same tables for the 4-node tet (N_i = lambda_i): Mass = (1 + delta)/20, Grad = delta_ak delta_cl
    (SURVEY.md B.1). No reference counterpart exists (F1): P1 is this repo's extension for configs 2-3."""
    mass = (1.0 + np.eye(4)) / 20.0
    grad = np.zeros((4, 4, 4, 4))
    for a in range(4):
        for c in range(4):
            grad[a, a, c, c] = 1.0
    return mass, grad


# ------------------------------------------------------------------------------------------------ numbering, pattern

def build_quad_mesh(tets, n_points):
    """BuildQuadMesh (mesh2modes.cpp:246-264): midside ids = n_points + first-seen rank of the edge key, scanning
    elements in order and edge slots in EdgeCorners order. Returns (nodes[T,10] uint32, node_count)."""
    t = tets.astype(np.int64)
    a, b = t[:, EDGE_CORNERS[:, 0]], t[:, EDGE_CORNERS[:, 1]]  # [T,6]
    key = (np.minimum(a, b) << 32) | np.maximum(a, b)
    flat = key.ravel()
    uniq, first, inverse = np.unique(flat, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")  # unique keys in first-seen order
    rank = np.empty(len(uniq), np.int64)
    rank[order] = np.arange(len(uniq))
    mid = (n_points + rank[inverse]).reshape(key.shape)
    nodes = np.concatenate([t, mid], axis=1).astype(np.uint32)
    return nodes, n_points + len(uniq)


def element_nodes(tets, n_points, order):
    if order == 2:
        return build_quad_mesh(tets, n_points)
    return tets.astype(np.uint32), n_points


def greedy_colouring(nodes, node_count):
    """Deterministic first-fit element colouring in element order (the reference has none; this is the
    definition the CUDA path must reproduce bit-exactly): element e takes the smallest colour not used by an earlier
    element sharing one of its nodes."""
    used = [0] * node_count  # bitmask of colours seen at each node (Python ints are unbounded)
    colours = np.zeros(len(nodes), np.uint32)
    for e, el in enumerate(nodes.tolist()):
        mask = 0
        for n in el:
            mask |= used[n]
        c = (~mask & (mask + 1)).bit_length() - 1
        colours[e] = c
        bit = 1 << c
        for n in el:
            used[n] |= bit
    return colours


@dataclass
class Csc:
    """Lower-triangular CSC exactly as Eigen::SparseMatrix after setFromTriplets (column-major, rows ascending,
    duplicates summed, explicit zeros kept): mesh2modes.cpp:322-325."""

    n: int
    colptr: np.ndarray
    rowidx: np.ndarray
    values: np.ndarray

    def to_scipy_full(self):
        import scipy.sparse as sp

        lower = sp.csc_matrix((self.values, self.rowidx.astype(np.int64), self.colptr.astype(np.int64)), shape=(self.n, self.n))
        return (lower + sp.tril(lower, -1).T).tocsc()


def _coo_to_csc(n, rows, cols, vals):
    key = cols.astype(np.int64) * n + rows.astype(np.int64)
    order = np.argsort(key, kind="stable")  # insertion order kept inside a duplicate run, like Eigen's two-pass transposition
    key, vals = key[order], vals[order]
    boundary = np.concatenate([[True], key[1:] != key[:-1]])
    starts = np.nonzero(boundary)[0]
    summed = np.add.reduceat(vals, starts) if len(vals) else vals
    ukey = key[starts]
    ucol, urow = ukey // n, ukey % n
    colptr = np.zeros(n + 1, np.int64)
    np.add.at(colptr, ucol + 1, 1)
    return Csc(n, np.cumsum(colptr).astype(np.int64), urow.astype(np.int64), summed)


def assemble(points, tets, material, order=2):
    """AssembleQuadratic (mesh2modes.cpp:273-327) for order 2; the same formulas with the 4-node tables for order 1.
    Returns (M, K, nodes, node_count): lower-triangular CSC mass and stiffness over 3*node_count DOFs."""
    mass_t, grad_t = quad_basis() if order == 2 else linear_basis()
    nn = mass_t.shape[0]
    nodes, node_count = element_nodes(tets, len(points), order)
    volume, phig = element_bases(points, tets)
    lam, mu = material.lam, material.mu
    # g[e,a,c,p,q] = sum_kl Grad[a,k,c,l] Phig[e,k,p] Phig[e,l,q]  (:300-311)
    tmp = np.einsum("akcl,ekp->eaclp", grad_t, phig, optimize=True)
    g = np.einsum("eaclp,elq->eacpq", tmp, phig, optimize=True)
    trace = np.einsum("eacpp->eac", g)
    ke = volume[:, None, None, None, None] * (lam * g + mu * np.swapaxes(g, 3, 4) + mu * trace[..., None, None] * np.eye(3))  # (:313-318)
    me = material.density * volume[:, None, None] * mass_t[None]  # (:298)

    nd = nodes.astype(np.int64)
    row_node = np.broadcast_to(nd[:, :, None], (len(tets), nn, nn))
    col_node = np.broadcast_to(nd[:, None, :], (len(tets), nn, nn))
    keep = row_node >= col_node  # `if (row < col) continue` (:297)
    rn, cn = row_node[keep], col_node[keep]
    # mass: the three diagonal entries of each kept block (:299)
    m_rows = (3 * rn[:, None] + np.arange(3)).ravel()
    m_cols = (3 * cn[:, None] + np.arange(3)).ravel()
    m_vals = np.repeat(me[keep], 3)
    # stiffness: full 3x3 for off-diagonal node blocks, q <= p on diagonal node blocks (:314-317)
    kb = ke[keep]  # [nb,3,3]
    p_idx, q_idx = np.meshgrid(np.arange(3), np.arange(3), indexing="ij")
    k_rows = (3 * rn[:, None, None] + p_idx).ravel()
    k_cols = (3 * cn[:, None, None] + q_idx).ravel()
    k_vals = kb.ravel()
    lower = k_rows >= k_cols
    n = 3 * node_count
    M = _coo_to_csc(n, m_rows, m_cols, m_vals)
    K = _coo_to_csc(n, k_rows[lower], k_cols[lower], k_vals[lower])
    return M, K, nodes, node_count


# ------------------------------------------------------------------------------------------------ eigensolve, post-process

def solver_sizes(config, n):
    """nev / ncv / sigma of ComputeModes (mesh2modes.cpp:455-459)."""
    nev = min(config.num_fem_modes, n - 1)
    ncv = min(max(nev + 20, 20), n)
    sigma = -((2 * math.pi * float(np.float32(config.min_mode_freq))) ** 2)
    return nev, ncv, sigma


def eigensolve(M, K, config, tol=1e-12):
    """The cold path of ComputeModes (mesh2modes.cpp:485-491): lowest nev pairs of K x = lambda M x by shift-invert
    about sigma < 0; eigenvectors M-orthonormal, ascending. ARPACK instead of Spectra (same Krylov space)."""
    import scipy.sparse.linalg as spla

    nev, ncv, sigma = solver_sizes(config, M.n)
    Mf, Kf = M.to_scipy_full(), K.to_scipy_full()
    vals, vecs = spla.eigsh(Kf, k=nev, M=Mf, sigma=sigma, which="LM", tol=tol, ncv=ncv, v0=np.ones(M.n))
    idx = np.argsort(vals)
    return vals[idx], vecs[:, idx]



def subspace_iterate(M, K, nev, p, sigma, tol, max_iters, x0, rng_seed=20260710):
    """SubspaceIterate (mesh2modes.cpp:339-428): the warm path of ComputeModes. x0 (n x cols, float32) seeds the leading
    panel columns, Gaussian columns fill the rest (the reference draws them from std::mt19937_64{20260710} through
    libc++'s normal_distribution; numpy's generator stands in, converged results do not depend on the draw). Returns
    (eigenvalues ascending [nev] or empty, eigenvectors n x nev M-orthonormal, iterations, op_applications)."""
    import scipy.linalg as sla
    import scipy.sparse.linalg as spla

    Mf, Kf = M.to_scipy_full().tocsc(), K.to_scipy_full().tocsc()
    n = Mf.shape[0]
    lu = spla.splu((Kf - sigma * Mf).tocsc())
    X = np.zeros((n, p))
    seeded = min(x0.shape[1], p)
    X[:, :seeded] = np.asarray(x0, np.float32)[:, :seeded].astype(np.float64)
    X[:, seeded:] = np.random.default_rng(rng_seed).standard_normal((n, p - seeded))
    MX = Mf @ X
    XL, MXL, theta_locked = np.zeros((n, nev)), np.zeros((n, nev)), np.zeros(nev)
    prev = np.full(nev, np.finfo(np.float64).max)
    c = iterations = ops = 0
    for it in range(max_iters):
        w = p - c
        Xbar = lu.solve(MX)
        ops += w
        Kr = Xbar.T @ MX
        MXbar = Mf @ Xbar
        if c > 0:
            C = XL[:, :c].T @ MXbar
            Xbar = Xbar - XL[:, :c] @ C
            MXbar = MXbar - MXL[:, :c] @ C
            Kr = Kr - C.T @ (theta_locked[:c, None] * C)
        Mr = Xbar.T @ MXbar
        Kr, Mr = 0.5 * (Kr + Kr.T), 0.5 * (Mr + Mr.T)
        d = 1.0 / np.sqrt(np.diag(Mr))
        Kr, Mr = d[:, None] * Kr * d[None, :], d[:, None] * Mr * d[None, :]
        try:
            theta, y = sla.eigh(Kr, Mr)
        except np.linalg.LinAlgError:
            return np.zeros(0), np.zeros((n, 0)), iterations, ops
        q = d[:, None] * y
        newly = 0
        for i in range(min(w, nev - c)):
            lam = theta[i] + sigma
            rel = abs(lam - prev[c + i]) / max(abs(lam), abs(sigma))
            prev[c + i] = lam
            if newly == i and rel < tol:
                newly += 1
        if newly:
            XL[:, c:c + newly] = Xbar @ q[:, :newly]
            MXL[:, c:c + newly] = MXbar @ q[:, :newly]
            theta_locked[c:c + newly] = theta[:newly]
            c += newly
        iterations = it + 1
        if c >= nev:
            return prev.copy(), XL, iterations, ops
        MX = MXbar @ q[:, newly:]
    return np.zeros(0), np.zeros((n, 0)), iterations, ops


def postprocess_modes(eigenvalues, shapes, shape_scale, material, config, positions):
    """modal::PostprocessModes (mesh2modes.cpp:515-588). shapes: [point][eigenpair][3] float32."""
    f32 = np.float32
    lam = np.asarray(eigenvalues, np.float64)
    m = len(lam)
    min_f, max_f = float(f32(config.min_mode_freq)), f32(config.max_mode_freq)
    lambda_eps = (2 * math.pi * min_f) ** 2 * 1e-10
    omega = np.where(lam > lambda_eps, np.sqrt(np.maximum(lam, 0)), 0.0)

    def c_of(w):
        return material.alpha + material.beta * (w * w)

    def damped_hz(w, c):
        d = w * w - 0.25 * c * c
        return math.sqrt(d) / (2 * math.pi) if d > 0 else 0.0

    freqs, t60s = np.zeros(m, f32), np.zeros(m, f32)
    lowest, lowest_freq = m, f32(0)
    for k in range(m):
        if omega[k] <= 0:
            continue
        freqs[k] = f32(damped_hz(omega[k], c_of(omega[k])))
        if lowest == m and freqs[k] >= f32(config.min_mode_freq):
            lowest, lowest_freq = k, freqs[k]
    if lowest == m:
        return Modes()
    freq_scale = f32(config.fundamental_freq) / lowest_freq if config.fundamental_freq else f32(1)
    ln1000 = math.log(1000.0)
    for k in range(lowest, m):
        ws = omega[k] * float(freq_scale)
        c = c_of(ws)
        freqs[k] = f32(damped_hz(ws, c))
        t60s[k] = f32((2 * ln1000) / c) if c > 0 else f32(0)
    max_mode = max_f * max(f32(1), freq_scale)
    highest = m
    while highest > lowest and freqs[highest - 1] > max_mode:
        highest -= 1
    n_modes = min(config.num_modes, m, highest - lowest)
    out_shapes = (np.asarray(shapes, f32)[:, lowest:lowest + n_modes, :] * f32(shape_scale)).astype(f32)
    return Modes(freqs[lowest:lowest + n_modes].copy(), t60s[lowest:lowest + n_modes].copy(), out_shapes, np.asarray(positions, f32), float(lowest_freq), lowest)


def rescale_modes(eigenvalues, shapes, solved_material, material, config, positions):
    """modal::RescaleModes (mesh2modes.cpp:590-603). None when the edit is not exactly scalable."""
    if len(eigenvalues) == 0 or material.poisson != solved_material.poisson:
        return None
    rho_ratio = material.density / solved_material.density
    scale = (material.young / solved_material.young) / rho_ratio
    return postprocess_modes(np.asarray(eigenvalues) * scale, shapes, np.float32(1 / math.sqrt(rho_ratio)), material, config, positions)


def sample_excitations(points, excite_positions, baked_scale=(1.0, 1.0, 1.0)):
    """Nearest tet point per excitation position, first hit wins ties, positions reaching one point merge
    (mesh2modes.cpp:620-645). Returns (points[], local positions f32, sample_point_of_excitation[])."""
    inv = 1.0 / np.asarray(baked_scale, np.float64)
    pts, local, remap, seen = [], [], np.zeros(len(excite_positions), np.uint32), {}
    ex = np.asarray(excite_positions, np.float32).astype(np.float64)
    for i, p in enumerate(ex):
        d = ((points - p) ** 2).sum(1)
        nearest = int(np.argmin(d))
        if nearest not in seen:
            seen[nearest] = len(pts)
            pts.append(nearest)
            local.append((points[nearest] * inv).astype(np.float32))
        remap[i] = seen[nearest]
    return np.asarray(pts, np.uint32), np.asarray(local, np.float32).reshape(-1, 3), remap


def mesh2modes(points, tets, material, excite_positions, baked_scale=(1.0, 1.0, 1.0), config=None, order=2, tol=1e-12, seed_basis=None):
    """modal::mesh2modes (mesh2modes.cpp:605-658): cold path, or the warm path when seed_basis (n x >=nev float32) fits
    (:459-472). Returns a dict mirroring ModalResult; eigenvalues empty when the warm path fails to converge."""
    config = config or SolverConfig()
    points = np.asarray(points, np.float64)
    tets = filter_degenerate(points, np.asarray(tets, np.uint32))
    length_to_si = float(sum(float(np.float32(s)) for s in baked_scale)) / 3.0
    props = mass_properties(points, tets, material.density, baked_scale, length_to_si)
    M, K, nodes, node_count = assemble(points, tets, material, order)
    ex_points, positions, remap = sample_excitations(points, excite_positions, baked_scale)
    nev, _, sigma = solver_sizes(config, M.n)
    iterations = ops = None
    if seed_basis is not None and seed_basis.shape[0] == M.n and seed_basis.shape[1] >= nev:
        vals, vecs, iterations, ops = subspace_iterate(M, K, nev, min(nev + 15, M.n), sigma, config.warm_tolerance, config.max_restarts, seed_basis)
        if len(vals) == 0:
            return dict(modes=Modes(), eigenvalues=vals, eigenvectors=vecs, iterations=iterations, op_applications=ops)
    else:
        vals, vecs = eigensolve(M, K, config, tol)
    shapes = np.stack([vecs[3 * p:3 * p + 3, :].T for p in ex_points.tolist()]).astype(np.float32) if len(ex_points) else np.zeros((0, len(vals), 3), np.float32)
    modes = postprocess_modes(vals, shapes, 1.0, material, config, positions)
    return dict(modes=modes, mass_props=props, eigenvalues=vals, eigenvectors=vecs, shapes=shapes, sample_point_of_excitation=remap, dofs=3 * node_count, stiffness_nonzeros=len(K.values), M=M, K=K, nodes=nodes, iterations=iterations, op_applications=ops)


# ------------------------------------------------------------------------------------------------ synthetic meshes

def kuhn_block(nx, ny, nz, size=(1.0, 1.0, 1.0)):
    """MakeBarTets (tests/ModalSolverTest.cpp:38-69): an nx x ny x nz grid of cells, each split into six tets around
    its main diagonal (Kuhn subdivision), same vertex ids, tet order and corner order as the reference.
    Returns (points f64 [V,3], tets uint32 [T,4])."""
    vy, vz = ny + 1, nz + 1
    i, j, k = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
    points = np.stack([size[0] * i / nx, size[1] * j / ny, size[2] * k / nz], -1).reshape(-1, 3).astype(np.float64)
    ci, cj, ck = (a.ravel() for a in np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij"))

    def vid(a, b, c):
        return (a * vy + b) * vz + c

    c = np.stack([vid(ci, cj, ck), vid(ci + 1, cj, ck), vid(ci, cj + 1, ck), vid(ci + 1, cj + 1, ck),
                  vid(ci, cj, ck + 1), vid(ci + 1, cj, ck + 1), vid(ci, cj + 1, ck + 1), vid(ci + 1, cj + 1, ck + 1)], -1)
    corners = np.array([[0, 1, 3, 7], [0, 3, 2, 7], [0, 2, 6, 7], [0, 6, 4, 7], [0, 4, 5, 7], [0, 5, 1, 7]])
    tets = c[:, corners].reshape(-1, 4).astype(np.uint32)
    return points, tets
