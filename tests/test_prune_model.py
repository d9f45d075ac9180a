"""Numpy model of the exact pruned scan's bookkeeping (csrc/agglom_pruned.cuh), run on the CPU:

* bounds E[i][s] <= min_j (d_ij - u_j) + C with the cumulative upward drift C of the u_j,
* the two columns a merge rewrites are folded in exactly by the next row test, the two rewritten rows are
  marked unknown,
* row bound beta, work list, refresh of every sub-segment that is read,
* gbest warm-started from a pool of pairs re-evaluated under the new row sums.

The model asserts, at every iteration of NJ and UPGMA loops on random (tie-heavy) maps, (1) the invariant
E - C <= min (d - u_j) for every live sub-segment, and (2) that the pair found by reading only what the
bounds allow is the dense scan's pair (same value, same lexicographic tie-break)."""
import numpy as np
import pytest

SW = 8  # columns per bound in the model


def run_model(n, nj, seed, ties):
    rng = np.random.default_rng(seed)
    A = rng.random((n, n))
    D = (A + A.T) / 2
    if ties:
        D = np.round(D * 4) / 4
    np.fill_diagonal(D, 0.0)
    ids = np.arange(n)
    k = n
    nsub = (n + SW - 1) // SW
    E = np.full((n, nsub), -np.inf)
    beta = np.full(n, -np.inf)
    R = D.sum(axis=1)
    C = 0.0
    pool = []            # slots (si > sj)
    prev = None          # (lo, hi, moved)
    read_total = dense_total = 0
    next_id = n
    while k > 2:
        tinv = 1.0 / (k - 2) if nj else 0.0
        u = tinv * R[:k] if nj else np.zeros(k)
        # ---- warm start: best pool pair under the current row sums
        gb = np.inf
        for (si, sj) in pool:
            gb = min(gb, D[si, sj] - u[si] - u[sj])
        # ---- row test with fold of the previous merge's two columns
        work = []
        for i in range(1, k):
            if prev is not None and i not in (prev[0], prev[1]):
                lo, hi, moved = prev
                if i > lo:
                    e = D[i, lo] - u[lo] + C
                    E[i, lo // SW] = min(E[i, lo // SW], e)
                    beta[i] = min(beta[i], e)
                if moved and i > hi:
                    e = D[i, hi] - u[hi] + C
                    E[i, hi // SW] = min(E[i, hi // SW], e)
                    beta[i] = min(beta[i], e)
            if not (beta[i] - C - u[i] > gb + 1e-9):
                work.append(i)
                beta[i] = np.inf
        # ---- invariant: every bound is valid
        for i in range(1, k):
            for s in range((i + SW - 1) // SW):
                c0, c1 = s * SW, min((s + 1) * SW, i)
                true = (D[i, c0:c1] - u[c0:c1]).min()
                assert E[i, s] - C <= true + 1e-12, (n, nj, seed, k, i, s)
        # ---- scan the work list in random order, lowering gb as better pairs are found
        best = None  # (key, (i, j))
        rng.shuffle(work)
        for i in work:
            for s in range((i + SW - 1) // SW):
                c0, c1 = s * SW, min((s + 1) * SW, i)
                if E[i, s] - C - u[i] > gb + 1e-9:
                    beta[i] = min(beta[i], E[i, s])
                    continue
                read_total += c1 - c0
                q = D[i, c0:c1] - u[i] - u[c0:c1]
                E[i, s] = (D[i, c0:c1] - u[c0:c1]).min() + C
                beta[i] = min(beta[i], E[i, s])
                for j in range(c0, c1):
                    key = (q[j - c0], min(ids[i], ids[j]), max(ids[i], ids[j]))
                    if best is None or key < best[0]:
                        best = (key, (i, j))
                gb = min(gb, q.min())
        dense_total += k * (k - 1) // 2
        # ---- the dense scan's answer
        dbest = None
        for i in range(1, k):
            for j in range(i):
                key = (D[i, j] - u[i] - u[j], min(ids[i], ids[j]), max(ids[i], ids[j]))
                if dbest is None or key < dbest[0]:
                    dbest = (key, (i, j))
        assert best is not None and best[0] == dbest[0], (n, nj, seed, k, best, dbest)
        i, j = dbest[1]
        lo, hi, last = min(i, j), max(i, j), k - 1
        # ---- merge (fast-mode arithmetic), drift of u for the columns that keep their node
        dij = D[lo, hi]
        newrow = np.zeros(k)
        for v in range(k):
            if v in (lo, hi):
                continue
            newrow[v] = 0.5 * ((D[lo, v] + D[hi, v]) - dij) if nj else (D[lo, v] + D[hi, v]) / 2
        Rn = R.copy()
        for v in range(k):
            if v not in (lo, hi):
                Rn[v] = R[v] - D[lo, v] - D[hi, v] + newrow[v]
        moved = hi != last
        Dn = D.copy()
        for v in range(k):
            if v in (lo, hi):
                continue
            Dn[lo, v] = Dn[v, lo] = newrow[v]
            if moved and v != last:
                Dn[hi, v] = Dn[v, hi] = D[last, v]
        if moved:
            Dn[hi, lo] = Dn[lo, hi] = newrow[last]
            Dn[hi, hi] = 0.0
            Rn[hi] = Rn[last]
            ids[hi] = ids[last]
        Dn[lo, lo] = 0.0
        ids[lo] = next_id
        next_id += 1
        D = Dn
        k -= 1
        Rn[lo] = D[lo, :k].sum()
        if nj and k > 2:
            un = Rn[:k] / (k - 2)
            keep = [v for v in range(k) if v not in (lo, hi)]
            C += max(0.0, max((un[v] - u[v] for v in keep), default=0.0))
        R = Rn
        E[lo, :] = -np.inf
        beta[lo] = -np.inf
        if moved:
            E[hi, :] = -np.inf
            beta[hi] = -np.inf
        prev = (lo, hi, moved)
        # ---- pool: keep pairs that still exist (rename the moved slot), add this scan's runner-up region
        npool = []
        for (si, sj) in pool + [best[1]]:
            if si in (lo, hi) or sj in (lo, hi):
                continue
            si = hi if si == last else si
            sj = hi if sj == last else sj
            if si >= k or sj >= k:
                continue
            npool.append((max(si, sj), min(si, sj)))
        # a few random pairs keep the pool alive, as the per-CTA suggestions do on the device
        for _ in range(3):
            a, b = rng.choice(k, size=2, replace=False)
            if a not in (lo, hi) and b not in (lo, hi):
                npool.append((max(a, b), min(a, b)))
        pool = npool[-8:]
    return read_total / max(dense_total, 1)


@pytest.mark.parametrize("nj", [True, False], ids=["nj", "upgma"])
@pytest.mark.parametrize("ties", [False, True], ids=["generic", "ties"])
@pytest.mark.parametrize("seed", range(3))
def test_bounds_stay_valid_and_selection_is_exact(nj, ties, seed):
    frac = run_model(36 + 5 * seed, nj, seed, ties)
    assert 0.0 < frac <= 1.0
