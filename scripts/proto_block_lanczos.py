"""CPU prototype (numpy/scipy) of the block thick-restart shift-invert Lanczos used on the device, to calibrate block
size / basis size / restart rule against single-vector operator counts. Not part of the product or the tests."""
import math, sys, time
import numpy as np
import scipy.sparse.linalg as spla
import scipy.linalg as sla
sys.path.insert(0, ".")
from oracle import modal as om

def block_lanczos(solve, Mf, n, nev, b, mcap, tol, max_restarts=200, seed=1, keep_rule="spectra"):
    rng = np.random.default_rng(seed)
    eps = np.finfo(float).eps; eps23 = eps ** (2 / 3)
    V = np.zeros((n, mcap + b))
    T = np.zeros((mcap + b, mcap + b))
    ops = 0
    def op(X):
        nonlocal ops
        ops += X.shape[1]
        return solve(Mf @ X)
    def morth(W):
        R = np.eye(W.shape[1])
        for _ in range(2):
            G = W.T @ (Mf @ W)
            L = np.linalg.cholesky(G)
            W = sla.solve_triangular(L, W.T, lower=True).T
            R = L.T @ R
        return W, R
    W, _ = morth(op(rng.uniform(-0.5, 0.5, (n, b))))
    V[:, :b] = W
    cur = b   # columns in V; last b are the residual block (not yet expanded)
    restarts = 0
    while True:
        while cur + b <= mcap + b and cur <= mcap:
            j0 = cur - b
            W = op(V[:, j0:cur])
            H = np.zeros((cur, b))
            for _ in range(2):
                h = V[:, :cur].T @ (Mf @ W)
                W -= V[:, :cur] @ h
                H += h
            T[:cur, j0:cur] = H
            T[j0:cur, :cur] = H.T
            if cur + b > mcap + b: break
            Q, R = morth(W)
            T[cur:cur + b, j0:cur] = R
            T[j0:cur, cur:cur + b] = R.T
            V[:, cur:cur + b] = Q
            cur += b
            if cur > mcap: break
        m = cur - b
        Tm = 0.5 * (T[:m, :m] + T[:m, :m].T)
        theta, S = np.linalg.eigh(Tm)
        order = np.argsort(-np.abs(theta))
        theta, S = theta[order], S[:, order]
        R = T[m:m + b, m - b:m]
        est = np.linalg.norm(R @ S[m - b:m, :], axis=0)
        conv = est[:nev] < tol * np.maximum(eps23, np.abs(theta[:nev]))
        nconv = int(conv.sum())
        if nconv >= nev or restarts >= max_restarts:
            return theta[:nev], V[:, :m] @ S[:, :nev], ops, restarts, nconv
        k = nev + min(nconv, (m - nev) // 2)
        k = min(k, m - b)
        # keep so that the remaining capacity is a whole number of blocks
        k = mcap - ((mcap - k) // b) * b
        Vk = V[:, :m] @ S[:, :k]
        Vn = V[:, m:m + b].copy()
        V[:, :k] = Vk; V[:, k:k + b] = Vn
        T[:] = 0
        T[np.arange(k), np.arange(k)] = theta[:k]
        B = R @ S[m - b:m, :k]
        T[k:k + b, :k] = B; T[:k, k:k + b] = B.T
        cur = k + b
        restarts += 1

if __name__ == "__main__":
    cells = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    num_modes = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    order = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    pts, tets = om.kuhn_block(cells, cells, cells, (0.3, 0.3, 0.3))
    mat = om.MATERIALS["Steel"]
    M, K, _, _ = om.assemble(pts, tets, mat, order)
    cfg = om.SolverConfig(num_modes=num_modes, num_fem_modes=num_modes + 15)
    nev, ncv, sigma = om.solver_sizes(cfg, M.n)
    Mf, Kf = M.to_scipy_full().tocsc(), K.to_scipy_full().tocsc()
    lu = spla.splu((Kf - sigma * Mf).tocsc())
    n = M.n
    print("n", n, "nev", nev, "ncv", ncv)
    cnt = [0]
    def _mv(x):
        cnt[0] += 1
        return lu.solve(np.asarray(x).ravel())
    Op = lambda: spla.LinearOperator((n, n), matvec=_mv, dtype=float)
    t0 = time.time()
    vals, _ = spla.eigsh(Kf, k=nev, M=Mf, sigma=sigma, which="LM", tol=1e-8, ncv=ncv, OPinv=Op())
    print("ARPACK ops", cnt[0], "time %.1f" % (time.time() - t0))
    ref = np.sort(vals)
    for b, extra in [(8, 64)]:
        mcap = ((nev + extra + b - 1) // b) * b
        t0 = time.time()
        th, X, ops, rs, nconv = block_lanczos(lambda Y: lu.solve(Y), Mf, n, nev, b, mcap, 1e-8)
        lam = np.sort(1 / th + sigma)
        print(f"b={b} mcap={mcap} ops={ops} block_ops={ops // b} restarts={rs} nconv={nconv} maxrel={np.max(np.abs(lam - ref) / np.abs(ref)):.2e} time {time.time() - t0:.1f}", flush=True)
    bad = np.nonzero(np.abs(lam - ref) / np.abs(ref) > 1e-7)[0]
    print("mismatch idx", bad[:10], "block", lam[bad[:5]], "arpack", ref[bad[:5]])
    print("block tail", lam[-6:], "\narpack tail", ref[-6:])
