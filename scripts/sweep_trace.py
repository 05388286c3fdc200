"""Reads an ME_SWEEP_TRACE file (cholesky.cu SparseCholesky::Solve: one panel application's task timeline) and prints, per
dependency level of the panel sweeps, when its tasks ran: the picture of where a sweep's time goes (critical path on the narrow
upper levels against streaming on the wide lower ones)."""
import sys

import numpy as np

TASK = np.dtype([("Base", "<u8"), ("Kind", "<u4"), ("Super", "<u4"), ("K", "<u4"), ("Limit", "<u4"), ("Ld", "<u4"), ("Row0", "<u4"), ("VecOffset", "<u4"), ("RowsBase", "<u4"),
                 ("LinkBegin", "<u4"), ("LinkCount", "<u4"), ("Need", "<u4"), ("Count", "<u4"), ("DiagColumn", "<u4"), ("Pad", "<u4")])


def main(path):
    raw = open(path, "rb").read()
    nf, nb, ns, size = np.frombuffer(raw, "<u4", 4)
    assert size == TASK.itemsize == 64
    at = 16
    fwd = np.frombuffer(raw, TASK, nf, at); at += 64 * nf
    bwd = np.frombuffer(raw, TASK, nb, at); at += 64 * nb
    level = np.frombuffer(raw, "<u4", ns, at); at += 4 * ns
    first = np.frombuffer(raw, "<u4", ns, at); at += 4 * ns
    stamps = np.frombuffer(raw, "<u8", 3 * (nf + nb), at).reshape(-1, 3).astype(np.int64)
    for name, tasks, st in (("forward", fwd, stamps[:nf]), ("backward", bwd, stamps[nf:])):
        t0 = st[:, 0].min()
        st = (st - t0) * 1e-3  # us
        lv = level[first[tasks["Super"]]]
        rows = np.where(tasks["Kind"] == 1, np.minimum(tasks["Count"] * 32, tasks["Limit"] - tasks["Row0"]), 32)
        cols = np.where(tasks["Kind"] == 2, tasks["Limit"], tasks["K"])
        mbytes = rows * cols * 8 / 1e6
        print(f"{name}: {len(tasks)} tasks, {st[:, 2].max():.1f} us, {mbytes.sum() / 1e3:.2f} GB of matrix; task time busy (ready -> end) median {np.median(st[:, 2] - st[:, 1]):.2f} us, "
              f"waiting (taken -> ready) median {np.median(st[:, 1] - st[:, 0]):.2f} us, mean {np.mean(st[:, 1] - st[:, 0]):.2f} us")
        print(" level  supers  diag/panel tasks     MB   first start   last ready   last end    span    GB/s   busy med (diag, panel)")
        order = sorted(set(lv.tolist()), reverse=(name == "backward"))
        for l in order:
            m = lv == l
            d = m & (tasks["Kind"] != 1)
            p = m & (tasks["Kind"] == 1)
            span = st[m, 2].max() - st[m, 0].min()
            busy = lambda k: np.median(st[k, 2] - st[k, 1]) if k.any() else 0.0
            print(f"{l:6d} {len(set(tasks['Super'][m].tolist())):7d} {d.sum():8d} {p.sum():8d} {mbytes[m].sum():8.1f} {st[m, 0].min():11.1f} {st[m, 1].max():11.1f} {st[m, 2].max():11.1f} {span:8.1f} "
                  f"{mbytes[m].sum() / max(span, 1e-3) * 1e3 / 1e3:7.0f} {busy(d):8.2f} {busy(p):8.2f}")


if __name__ == "__main__":
    main(sys.argv[1])
