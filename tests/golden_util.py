"""Shared helpers for the golden modal models (tests/golden/*.npz; see tests/golden/make_golden.py)."""
import glob
import math
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LN1000 = np.float32(3 * math.log(10.0))


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: z[k] for k in z.files}


def decay_rates(t60s):
    """tests/ModalSolveTool.cpp:96-99: decay = ln1000 / T60 in float32."""
    t60s = np.asarray(t60s, np.float32)
    return np.where(t60s > 0, LN1000 / np.where(t60s > 0, t60s, 1), np.float32(0)).astype(np.float32)


def clusters(freqs, rel=2e-4):
    """Index ranges of (near-)degenerate frequency clusters: symmetric bodies have exact multiplicities and the
    eigenvectors inside one are an arbitrary rotation (SURVEY.md §7 hard parts), so shapes compare per cluster."""
    out, start = [], 0
    for k in range(1, len(freqs) + 1):
        if k == len(freqs) or abs(freqs[k] - freqs[k - 1]) > rel * abs(freqs[k]):
            out.append((start, k))
            start = k
    return out


def subspace_sine(a, b):
    """Largest principal-angle sine between the column spaces of a and b ([n, k] each)."""
    qa, _ = np.linalg.qr(a)
    qb, _ = np.linalg.qr(b)
    s = np.linalg.svd(qa.T @ qb, compute_uv=False)
    return math.sqrt(max(0.0, 1.0 - min(s) ** 2))


def compare_shapes(shapes_pm, golden_mp, freqs, tol_sine=1e-4, tol_norm=1e-4):
    """shapes_pm: ours [point][mode][3]; golden_mp: golden mode-major [mode][point][3]. Per frequency cluster the
    sampled shape vectors must span the same subspace and carry the same Frobenius norm."""
    ours = np.transpose(np.asarray(shapes_pm, np.float64), (1, 0, 2)).reshape(len(freqs), -1).T  # [3P, modes]
    gold = np.asarray(golden_mp, np.float64).reshape(len(freqs), -1).T
    worst_sine, worst_norm = 0.0, 0.0
    for lo, hi in clusters(freqs):
        worst_sine = max(worst_sine, subspace_sine(ours[:, lo:hi], gold[:, lo:hi]))
        na, nb = np.linalg.norm(ours[:, lo:hi]), np.linalg.norm(gold[:, lo:hi])
        worst_norm = max(worst_norm, abs(na - nb) / nb)
    return worst_sine, worst_norm
