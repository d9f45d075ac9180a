"""TEST INFRASTRUCTURE -- numpy model of the row-sharded loop's exchange protocol.

Restates, on the CPU and over any torch.distributed backend (gloo in the tests), exactly what the
CUDA path exchanges per iteration (csrc/sharded.cu): every rank scans only its own block-cyclic rows,
the ranks all-gather one candidate each and reduce them identically, then all-gather the three columns
(lo, hi, last) of their rows -- by symmetry the full rows -- and every rank applies the merge to its own
rows.  It uses the product's partition arithmetic (cassiopeia_b200.sharded.owner_of / local_of / ...), so
the tests pin that arithmetic and the sufficiency of the exchanged data against the C oracle.  Values are
float64 and the per-element arithmetic is the reference's, so UPGMA must reproduce the oracle join for
join; NJ uses maintained row sums like the GPU fast mode and is compared across world sizes.
"""
import numpy as np
import torch
import torch.distributed as dist

from cassiopeia_b200 import sharded


def _gather(arr: np.ndarray, world: int):
    t = torch.from_numpy(np.ascontiguousarray(arr))
    if world == 1:
        return [arr]
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [o.numpy() for o in out]


def run(D_full: np.ndarray, algo: str, rank: int, world: int):
    """D_full is only used to cut out this rank's rows; afterwards the rank sees its shard alone."""
    n = D_full.shape[0]
    rows = sharded.shard_rows(n, rank, world)
    cap = sharded.local_capacity(n, world)
    shard = np.zeros((cap, n), dtype=np.float64)
    shard[: len(rows)] = D_full[rows]
    ids = np.arange(n, dtype=np.int64)
    sizes = np.ones(n, dtype=np.int64)
    R = None
    if algo == "nj":
        mine = np.zeros(cap)
        mine[: len(rows)] = shard[: len(rows)].sum(axis=1)
        parts = _gather(mine, world)
        R = np.array([parts[sharded.owner_of(v, world)][sharded.local_of(v, world)] for v in range(n)])
    joins = []
    for t in range(n - 2):
        k = n - t
        nloc = sharded.local_rows(k, rank, world)
        # ---- scan of the local rows (lower triangle only), lexicographic (value, id_lo, id_hi)
        best = (np.inf, 2**62, 2**62, -1, -1)
        for lr in range(nloc):
            i = sharded.slot_of(lr, rank, world)
            if i == 0:
                continue
            d = shard[lr, :i]
            q = d - (1.0 / (k - 2)) * (R[i] + R[:i]) if algo == "nj" else d
            qmin = q.min()
            if qmin <= best[0]:
                for j in np.nonzero(q == qmin)[0]:
                    cand = (qmin, min(ids[i], ids[j]), max(ids[i], ids[j]), i, int(j))
                    if cand[:3] < best[:3]:
                        best = cand
        # ---- C1: one candidate per rank, reduced identically everywhere
        cands = _gather(np.array(best, dtype=np.float64), world)
        win = min(cands, key=lambda c: (c[0], c[1], c[2]))
        lo, hi = int(min(win[3], win[4])), int(max(win[3], win[4]))
        last = k - 1
        joins.append((int(win[1]), int(win[2])))
        # ---- C2: the three columns of the local rows == the full rows lo, hi, last
        chunk = sharded.local_capacity(k, world)
        cols = np.zeros((3, chunk))
        cols[:, :nloc] = shard[:nloc][:, [lo, hi, last]].T
        parts = _gather(cols, world)
        full = np.zeros((3, k))
        for v in range(k):
            full[:, v] = parts[sharded.owner_of(v, world)][:, sharded.local_of(v, world)]
        a, b, l = full
        dij = a[hi]
        if algo == "nj":
            nv = 0.5 * ((a + b) - dij)
        else:
            zi, zj = float(sizes[lo]), float(sizes[hi])
            nv = ((zi * a) + (zj * b)) / (zi + zj)
        # ---- merge on the local rows, swap compaction (slot k-1 moves into hi)
        others = np.array([v for v in range(k) if v not in (lo, hi)], dtype=np.int64)
        if algo == "nj":
            R[others] = R[others] - a[others] - b[others] + nv[others]
        for lr in range(nloc):
            v = sharded.slot_of(lr, rank, world)
            if v in (lo, hi):
                continue
            if hi != last and v != last:
                shard[lr, hi] = l[v]
            shard[lr, lo] = nv[v]
        if sharded.owner_of(lo, world) == rank:
            row = shard[sharded.local_of(lo, world)]
            row[others] = nv[others]
            row[lo] = 0.0
            if hi != last:
                row[hi] = nv[last]
        if hi != last and sharded.owner_of(hi, world) == rank:
            row = shard[sharded.local_of(hi, world)]
            keep = np.array([v for v in range(k) if v not in (lo, hi, last)], dtype=np.int64)
            row[keep] = l[keep]
            row[lo] = nv[last]
            row[hi] = 0.0
        new_size = sizes[lo] + sizes[hi]
        if algo == "nj":
            R_new = nv[others].sum()
            if hi != last:
                R[hi] = R[last]
            R[lo] = R_new
        if hi != last:
            ids[hi], sizes[hi] = ids[last], sizes[last]
        ids[lo], sizes[lo] = n + t, new_size
    return np.array(joins, dtype=np.int64)
