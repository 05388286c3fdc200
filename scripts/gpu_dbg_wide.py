import math, sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from golden_util import load_golden
from oracle import modal as om
from mesheditor_b200 import Factor, FemSystem
import scipy.sparse.linalg as spla
SIGMA = -((2 * math.pi * 20.0) ** 2)
def be(A, x, b):
    return np.linalg.norm(A @ x - b) / (spla.norm(A, 1) * np.linalg.norm(x) + np.linalg.norm(b))
def run(name, points, tets, mat, order):
    M, K, _, _ = om.assemble(points, tets, mat, order)
    A = (K.to_scipy_full() - SIGMA * M.to_scipy_full()).tocsc()
    fem = FemSystem(points, tets, mat, order)
    f = Factor(fem, SIGMA)
    rng = np.random.default_rng(5)
    for width in (1, 3, 8, 11):
        B = rng.standard_normal((A.shape[0], width))
        X = f.solve(B if width > 1 else B[:, 0].copy())
        X = X.reshape(A.shape[0], -1)
        errs = [be(A, X[:, k], B[:, k]) for k in range(width)]
        singles = [f.solve(B[:, k].copy()) for k in range(width)]
        d = [np.linalg.norm(X[:, k] - singles[k]) / np.linalg.norm(singles[k]) for k in range(width)]
        print(name, "width", width, "backward err max %.3e" % max(errs), "vs single max %.3e" % max(d), "single be %.3e" % max(be(A, singles[k], B[:, k]) for k in range(width)), flush=True)
g = load_golden("bracket_steel")
run("bracket", g["points"], g["tets"], om.MATERIALS["Steel"], 2)
p, t = om.kuhn_block(9, 8, 7, size=(0.5, 0.4, 0.3))
run("kuhn P2", p, t, om.MATERIALS["Ceramic"], 2)
