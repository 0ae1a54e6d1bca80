"""How much of the all-gathered panel does a rank actually read?  CPU-only analysis (numpy) of the 1-D row
partition: for a uniform and an R-MAT graph of the products shape, the share of each remote slot's
rows that appear as a column in the rank's row block -- the bytes a needed-rows-only ("halo") exchange would move,
relative to the full all-gather bench.py / dist.py perform today.

    python tools/halo_fraction.py [n=2449029] [avg_deg=25] [world=8]
"""
import sys

import numpy as np


def rmat(n, n_edges, rs, a=0.57, b=0.19, c=0.19):
    bits = max(1, int(n - 1).bit_length())
    src = np.zeros(n_edges, np.int64)
    dst = np.zeros(n_edges, np.int64)
    for _ in range(bits):
        r = rs.random(n_edges)
        src = src * 2 + (r >= a + b)
        dst = dst * 2 + (((r >= a) & (r < a + b)) | (r >= a + b + c))
    return src % n, dst % n


def analyse(name, src, dst, n, world):
    # max-symmetrise + self loops: the pattern of utils.py:360-368
    keys = np.unique(np.concatenate([src * n + dst, dst * n + src, np.arange(n, dtype=np.int64) * (n + 1)]))
    rows, cols = keys // n, keys % n
    nnz = keys.size
    rowptr = np.searchsorted(rows, np.arange(n + 1))
    report(name, rows, cols, n, nnz, rowptr, world,
           [0] + [int(np.searchsorted(rowptr, nnz * k // world)) for k in range(1, world)] + [n], "entries only")
    # cost = entries + (average degree) * rows: what dist.partition_rows_by_nnz(row_weight=avg degree) cuts
    w = nnz / n
    cost = rowptr + w * np.arange(n + 1)
    report(name, rows, cols, n, nnz, rowptr, world,
           [0] + [int(np.searchsorted(cost, cost[-1] * k / world)) for k in range(1, world)] + [n],
           "entries + avg degree x rows")


def report(name, rows, cols, n, nnz, rowptr, world, bounds, how):
    total_needed = total_full = 0
    worst = 0.0
    for p in range(world):
        lo, hi = rowptr[bounds[p]], rowptr[bounds[p + 1]]
        c = np.unique(cols[lo:hi])
        remote = c[(c < bounds[p]) | (c >= bounds[p + 1])]
        full = n - (bounds[p + 1] - bounds[p])
        total_needed += remote.size
        total_full += full
        worst = max(worst, remote.size / max(full, 1))
    per = [bounds[i + 1] - bounds[i] for i in range(world)]
    ent = [int(rowptr[bounds[i + 1]] - rowptr[bounds[i]]) for i in range(world)]
    print("%-8s n=%d nnz=%d world=%d, balanced by %s: rows per rank %s" % (name, n, nnz, world, how, per))
    print("         padded all-gather rows / real rows: %.2f; largest block's entries / mean: %.2f" % (
        world * max(per) / n, max(ent) / (nnz / world)))
    print("         remote panel rows a rank reads: %.1f %% of what the all-gather delivers (worst rank %.1f %%)" % (
        100.0 * total_needed / total_full, 100.0 * worst))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_449_029
    deg = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    world = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    rs = np.random.default_rng(0)
    n_raw = n * deg // 2
    analyse("uniform", rs.integers(0, n, n_raw), rs.integers(0, n, n_raw), n, world)
    s, d = rmat(n, n_raw, rs)
    analyse("R-MAT", s, d, n, world)


if __name__ == "__main__":
    main()
