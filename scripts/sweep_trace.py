"""Reads an ME_SWEEP_TRACE file (cholesky.cu SparseCholesky::Solve: the task timeline of one panel application) and prints where
the two sweeps' time goes: how many of the resident CTAs are working / waiting for inputs over time, the time per dependency level,
and the streaming rate of the tasks themselves (ready -> end per KB of matrix)."""
import gzip
import sys

import numpy as np

TASK = np.dtype([("Base", "<u8"), ("Kind", "<u4"), ("Super", "<u4"), ("K", "<u4"), ("Limit", "<u4"), ("Ld", "<u4"), ("Row0", "<u4"), ("VecOffset", "<u4"), ("RowsBase", "<u4"),
                 ("LinkBegin", "<u4"), ("LinkCount", "<u4"), ("Need", "<u4"), ("Count", "<u4"), ("DiagColumn", "<u4"), ("Pad", "<u4")])


def load(path):
    raw = (gzip.open(path) if path.endswith(".gz") else open(path, "rb")).read()
    nf, nb, ns, size = np.frombuffer(raw, "<u4", 4)
    assert size == TASK.itemsize == 64
    at = 16
    fwd = np.frombuffer(raw, TASK, nf, at); at += 64 * nf
    bwd = np.frombuffer(raw, TASK, nb, at); at += 64 * nb
    level = np.frombuffer(raw, "<u4", ns, at); at += 4 * ns
    first = np.frombuffer(raw, "<u4", ns, at); at += 4 * ns
    stamps = np.frombuffer(raw, "<u8", 3 * (nf + nb), at).reshape(-1, 3).astype(np.int64)
    return fwd, bwd, level, first, stamps[:nf], stamps[nf:]


def main(path, bins=24):
    fwd, bwd, level, first, sf, sb = load(path)
    for name, tasks, st in (("forward", fwd, sf), ("backward", bwd, sb)):
        st = (st - st[:, 0].min()) / 1e3  # us
        total = st[:, 2].max()
        wait, busy = st[:, 1] - st[:, 0], st[:, 2] - st[:, 1]
        lv = level[first[tasks["Super"]]]
        rows = np.where(tasks["Kind"] == 1, np.minimum(tasks["Count"] * 32, tasks["Limit"] - tasks["Row0"]), 32)
        cols = np.where(tasks["Kind"] == 2, tasks["Limit"], tasks["K"])
        kb = rows * cols * 8 / 1024
        print(f"{name}: {len(tasks)} tasks over {len(set(lv.tolist()))} levels, {total:.0f} us, {kb.sum() / 1e6:.2f} GB of matrix = {kb.sum() * 1024 / total / 1e6:.2f} TB/s")
        print(f"  CTA time: working {busy.sum() / total:.0f} CTAs on average, waiting for inputs {wait.sum() / total:.0f}; a task's own rate (ready -> end): median {np.median(busy / kb) * 32:.2f} us per 32 KB, "
              f"mean {busy.sum() / kb.sum() * 32:.2f}")
        edges = np.linspace(0, total, bins + 1)
        work = [np.clip(np.minimum(st[:, 2], b) - np.maximum(st[:, 1], a), 0, None).sum() / (b - a) for a, b in zip(edges[:-1], edges[1:])]
        idle = [np.clip(np.minimum(st[:, 1], b) - np.maximum(st[:, 0], a), 0, None).sum() / (b - a) for a, b in zip(edges[:-1], edges[1:])]
        mb = [kb[(st[:, 2] >= a) & (st[:, 2] < b)].sum() / 1024 for a, b in zip(edges[:-1], edges[1:])]
        print(f"  per {total / bins:.0f} us: working CTAs {np.round(work).astype(int).tolist()}")
        print(f"  {'':>10} waiting CTAs {np.round(idle).astype(int).tolist()}")
        print(f"  {'':>10} MB finished  {np.round(mb).astype(int).tolist()}")
        order = sorted(set(lv.tolist()), reverse=(name == "backward"))
        ends = np.array([st[lv == l, 2].max() for l in order])
        pitch = np.diff(np.concatenate([[0.0], ends]))
        groups = [order[i:i + max(1, len(order) // 12)] for i in range(0, len(order), max(1, len(order) // 12))]
        print("  levels (in sweep order)   tasks       MB    ends at   us per level   TB/s")
        at = 0
        for g in groups:
            m = np.isin(lv, g)
            span = pitch[at:at + len(g)].sum()
            print(f"  {g[0]:4d}..{g[-1]:4d} {m.sum():14d} {kb[m].sum() / 1024:8.1f} {ends[at + len(g) - 1]:10.0f} {span / len(g):14.1f} {kb[m].sum() * 1024 / max(span, 1e-9) / 1e6:6.2f}")
            at += len(g)


if __name__ == "__main__":
    main(*sys.argv[1:2])
