"""Pruned scan (default; lower bounds per row x 128-column sub-segment) == dense scan (CAS_PRUNE=0):
identical join lists, remaining pair and near-tie diagnostics; and both == the CPU oracle's joins where the
arithmetic is the reference's (UPGMA float64)."""
import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def _map(n, m, S, dtype, rate, miss, weighted, seed):
    from cassiopeia_b200 import engine
    from cassiopeia_b200.synthetic import simulate_character_matrix

    cm, priors, table = simulate_character_matrix(n, m, S, seed=seed, mutation_rate=rate, missing_fraction=miss)
    cm = np.vstack([cm, np.zeros((1, m), dtype=cm.dtype)])
    with np.errstate(invalid="ignore", divide="ignore"):
        w = np.nan_to_num(-np.log(table), nan=0.0, posinf=0.0)
    return engine.whd_map(cm, -1, w if weighted else None, S, dtype=dtype, method="exact"), n + 1


def _run(D, n, algo, mode, prune, monkeypatch, keyed=True):
    from cassiopeia_b200 import engine

    monkeypatch.setenv("CAS_PRUNE", "1" if prune else "0")
    monkeypatch.setenv("CAS_UPGMA_KEYED", "1" if keyed else "0")  # UPGMA: tie-aware bounds (default) / value-only
    monkeypatch.setenv("CAS_PRUNE_MIN_K", "0")  # keep the pruned scan down to the last iteration
    joins, last2 = engine.agglomerate(D.clone(), n, algo, mode)
    return joins, last2, dict(engine.last_stats)


@pytest.mark.parametrize("n,m,S,rate,miss", [(3, 5, 3, 0.5, 0.0), (130, 20, 5, 0.7, 0.1), (700, 40, 50, 0.1, 0.2),
                                             (2500, 30, 10, 0.4, 0.3), (4200, 60, 100, 0.1, 0.2)])
@pytest.mark.parametrize("algo,mode,dtype,weighted", [("nj", "fast", "f32", True), ("nj", "fast", "f64", True),
                                                      ("upgma", "fast", "f32", True),
                                                      ("upgma", "strict", "f64", False)])
def test_pruned_equals_dense(n, m, S, rate, miss, algo, mode, dtype, weighted, monkeypatch):
    D, nn = _map(n, m, S, dtype, rate, miss, weighted, seed=n)
    jd, ld, sd = _run(D, nn, algo, mode, False, monkeypatch)
    jp, lp, sp = _run(D, nn, algo, mode, True, monkeypatch)
    assert np.array_equal(jd, jp) and np.array_equal(ld, lp)
    if algo == "upgma":
        # tie-aware bounds skip tied pairs that lose on the id key: its tie counters see only the pairs that were read
        # (diagnostics, not part of the result); the value-only variant reproduces the dense scan's counters
        jv, lv, sv = _run(D, nn, algo, mode, True, monkeypatch, keyed=False)
        assert np.array_equal(jd, jv) and np.array_equal(ld, lv)
        # (exact ties only: which runner-ups inside the near-tie window are evaluated depends on the order in which the
        # scan meets them, for UPGMA the screening threshold carries no window)
        assert sd["exact_ties"] == sv["exact_ties"] and sp["exact_ties"] <= sd["exact_ties"]
    else:
        assert (sd["near_ties"], sd["exact_ties"]) == (sp["near_ties"], sp["exact_ties"])
    assert sd["scanned_elements"] == 0
    dense = sum(k * (k - 1) // 2 for k in range(3, nn + 1))
    assert 0 < sp["scanned_elements"] <= dense
    if algo == "upgma" and mode == "strict" and nn <= 1000:
        # float64 UPGMA has no reductions: the GPU loop is the reference's arithmetic
        oj, ol = po.upgma(D[:, :nn].cpu().numpy())
        assert np.array_equal(jp, oj)


@pytest.mark.parametrize("n,m,S,rate,miss,dup", [(300, 10, 3, 0.05, 0.0, 3), (1500, 20, 4, 0.1, 0.1, 4),
                                                 (5000, 40, 50, 0.1, 0.1, 2), (6000, 12, 2, 0.3, 0.0, 1)])
@pytest.mark.parametrize("dtype,mode", [("f64", "strict"), ("f32", "fast")])
def test_upgma_tie_heavy_inputs_keyed_bounds(n, m, S, rate, miss, dup, dtype, mode, monkeypatch):
    """Plain Hamming distances over few characters and duplicated cells: thousands of pairs tie with the minimum at
    most iterations (distance 0 inside duplicate groups, then a handful of distinct rationals).  The tie-aware
    pruned scan must select the dense scan's pairs -- the reference's first row-major arg-min -- and read far less
    than the value-only pruned scan, which has to re-read every tied pair every iteration."""
    from cassiopeia_b200 import engine
    from cassiopeia_b200.synthetic import simulate_character_matrix

    cm, _, _ = simulate_character_matrix(n, m, S, seed=n + dup, mutation_rate=rate, missing_fraction=miss)
    rng = np.random.default_rng(n)
    cm = np.concatenate([cm] + [cm[rng.integers(0, n, n // 2)] for _ in range(dup - 1)])  # duplicated cells, shuffled
    cm = cm[rng.permutation(cm.shape[0])]
    nn = cm.shape[0]
    D = engine.whd_map(cm, -1, None, S, dtype=dtype, method="exact")
    jd, ld, sd = _run(D, nn, "upgma", mode, False, monkeypatch)
    jk, lk, sk = _run(D, nn, "upgma", mode, True, monkeypatch)
    assert np.array_equal(jd, jk) and np.array_equal(ld, lk)
    assert sd["exact_ties"] > nn // 4  # the input really is tie-heavy
    jv, lv, sv = _run(D, nn, "upgma", mode, True, monkeypatch, keyed=False)
    assert np.array_equal(jd, jv)
    print(f"n={nn} {dtype}: elements read keyed {sk['scanned_elements']:.3e} value-only {sv['scanned_elements']:.3e}")
    if nn >= 1500:
        assert sk["scanned_elements"] < 0.5 * sv["scanned_elements"]
    if dtype == "f64" and nn <= 4000:
        oj, ol = po.upgma(D[:, :nn].cpu().numpy())
        assert np.array_equal(jk, oj) and np.array_equal(lk, ol)


@pytest.mark.parametrize("kind", ["small_ints", "gaussian", "all_equal", "two_values", "uniform"])
@pytest.mark.parametrize("n", [6, 200, 1300])
def test_upgma_keyed_bounds_on_arbitrary_symmetric_maps(n, kind, monkeypatch):
    """User-supplied maps: massive ties (integers, a single value, two values), negative entries, no ties at all.
    Tie-aware pruned UPGMA == dense == the oracle's first row-major arg-min."""
    import torch
    from cassiopeia_b200 import engine

    rng = np.random.default_rng(n * 13 + len(kind))
    if kind == "small_ints":
        A = rng.integers(0, 3, (n, n)).astype(np.float64)
    elif kind == "gaussian":
        A = rng.normal(size=(n, n))
    elif kind == "all_equal":
        A = np.full((n, n), 0.25)
    elif kind == "two_values":
        A = np.where(rng.random((n, n)) < 0.02, 0.0, 1.0 / 3.0)
    else:
        A = rng.random((n, n))
    D = np.triu(A, 1)
    D = D + D.T
    host = np.zeros((n, engine.padded_ld(n)), dtype=np.float64)
    host[:, :n] = D
    dev = torch.from_numpy(host).cuda()
    want, want_last = po.upgma(D)
    jk, lk, sk = _run(dev, n, "upgma", "strict", True, monkeypatch)
    assert np.array_equal(jk, want) and np.array_equal(lk, want_last), kind
    jd, ld, sd = _run(dev, n, "upgma", "strict", False, monkeypatch)
    assert np.array_equal(jd, want) and np.array_equal(ld, want_last)
    # float32 copy of the same map (values exactly representable or not: the loop's own arithmetic decides)
    dev32 = dev.float()
    j32k, l32k, _ = _run(dev32, n, "upgma", "fast", True, monkeypatch)
    j32d, l32d, _ = _run(dev32, n, "upgma", "fast", False, monkeypatch)
    assert np.array_equal(j32k, j32d) and np.array_equal(l32k, l32d)


def test_default_policy_switches_to_dense_at_small_sizes(monkeypatch):
    """Default: NJ prunes above 2048 live nodes and finishes with the dense scan; same joins either way."""
    from cassiopeia_b200 import engine

    D, nn = _map(9000, 30, 10, "f32", 0.4, 0.2, True, seed=4)
    monkeypatch.delenv("CAS_PRUNE", raising=False)
    monkeypatch.delenv("CAS_PRUNE_MIN_K", raising=False)
    j_default, l_default = engine.agglomerate(D.clone(), nn, "nj", "fast")
    assert engine.last_stats["scanned_elements"] > 0
    jd, ld, sd = _run(D, nn, "nj", "fast", False, monkeypatch)
    assert np.array_equal(j_default, jd) and np.array_equal(l_default, ld)


def test_pruned_scan_reads_less_at_size(monkeypatch):
    """At 12k cells the pruned NJ scan reads a small fraction of the dense scan's elements."""
    D, nn = _map(12000, 40, 50, "f32", 0.7, 0.1, True, seed=2)
    jp, lp, sp = _run(D, nn, "nj", "fast", True, monkeypatch)
    jd, ld, sd = _run(D, nn, "nj", "fast", False, monkeypatch)
    assert np.array_equal(jd, jp) and np.array_equal(ld, lp)
    dense = sum(k * (k - 1) // 2 for k in range(3, nn + 1))
    assert sp["scanned_elements"] < 0.25 * dense


def test_work_list_stays_short_as_the_drift_grows(monkeypatch):
    """The rows listed per iteration must not grow with the cumulative drift C: a sub-segment bound that decodes
    as 0 (an encoded NaN from the lanes past the triangle) is valid but puts every row on the list once
    C > gbest + u_i, i.e. for the whole second half of a solve (agglom_pruned.cuh, prune_bound)."""
    from cassiopeia_b200 import engine

    D, nn = _map(8000, 60, 100, "f32", 0.1, 0.2, True, seed=6)
    monkeypatch.setenv("CAS_PRUNE", "1")
    monkeypatch.setenv("CAS_PRUNE_MIN_K", "0")
    listed = {}
    for frac in (0.5, 0.9):
        engine.agglomerate(D.clone(), nn, "nj", "fast", max_iters=int(frac * nn))
        listed[frac] = engine.last_stats["rows_listed"]
        assert engine.last_stats["drift"] > 0.0
    late = (listed[0.9] - listed[0.5]) / (0.4 * nn)    # rows listed per iteration between 50 % and 90 %
    live = 0.3 * nn                                    # mean live rows in that window
    assert late < 0.1 * live, (listed, late)


@pytest.mark.parametrize("algo", ["nj", "upgma"])
def test_short_units_for_every_row_equal_wide_units(algo, monkeypatch):
    """CAS_PRUNE_WIDE=0 (every listed row in 8-sub-segment units) and the default split work list (rows with a
    known bound in 32-sub-segment units) select the same pairs."""
    D, nn = _map(5000, 40, 50, "f32", 0.3, 0.2, True, seed=9)
    # the switch belongs to the value-only bounds (NJ, and UPGMA with CAS_UPGMA_KEYED=0)
    monkeypatch.setenv("CAS_PRUNE_WIDE", "0")
    j0, l0, s0 = _run(D, nn, algo, "fast", True, monkeypatch, keyed=False)
    monkeypatch.setenv("CAS_PRUNE_WIDE", "1")
    j1, l1, s1 = _run(D, nn, algo, "fast", True, monkeypatch, keyed=False)
    assert np.array_equal(j0, j1) and np.array_equal(l0, l1)
    assert (s0["near_ties"], s0["exact_ties"]) == (s1["near_ties"], s1["exact_ties"])
    assert s0["rows_listed"] > 0 and s1["rows_listed"] > 0
