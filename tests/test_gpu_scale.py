"""Size-independent properties at benchmark-like sizes (the oracle is too slow here): exact == tensor
map, symmetry, zero diagonal, valid join lists, sharded == single GPU, strict UPGMA == fast-f64 UPGMA."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _workload(n, m, S, missing, seed=0):
    from cassiopeia_b200.synthetic import simulate_character_matrix

    cm, priors, table = simulate_character_matrix(n, m, S, seed=seed, mutation_rate=0.1, missing_fraction=missing)
    with np.errstate(invalid="ignore"):
        w = np.nan_to_num(-np.log(table))
    return cm, w


def _valid_join_list(joins, last2, n):
    used = np.concatenate([joins.ravel(), last2])
    return np.array_equal(np.sort(used), np.arange(2 * n - 2)) and bool((joins[:, 1] < n + np.arange(n - 2)).all())


def test_unweighted_map_20k_exact_equals_tensor_and_is_symmetric():
    import torch
    from cassiopeia_b200 import engine

    n, m, S = 20000, 60, 100
    cm, _ = _workload(n, m, S, 0.2)
    A = engine.whd_map(cm, -1, None, S, dtype="f32", method="exact")[:, :n]
    B = engine.whd_map(cm, -1, None, S, dtype="f32", method="tensor")[:, :n]
    assert torch.equal(A, B)
    assert torch.equal(A, A.T) and float(A.diagonal().abs().max()) == 0.0
    assert float(A.min()) >= 0.0 and float(A.max()) <= 2.0


def test_weighted_map_20k_tensor_within_tolerance():
    from cassiopeia_b200 import engine

    n, m, S = 20000, 40, 50
    cm, w = _workload(n, m, S, 0.1, seed=3)
    A = engine.whd_map(cm, -1, w, S, dtype="f32", method="exact")[:, :n]
    B = engine.whd_map(cm, -1, w, S, dtype="f32", method="tensor")[:, :n]
    rel = ((A.double() - B.double()).abs() / A.double().abs().clamp(min=1.0)).max().item()
    assert rel <= 1e-6


def test_fast_nj_20k_is_a_valid_tree_and_sharding_does_not_change_it():
    from cassiopeia_b200 import engine, sharded

    n, m, S = 6000, 40, 50
    cm, w = _workload(n, m, S, 0.2, seed=5)
    D = engine.whd_map(cm, -1, w, S, dtype="f32")
    joins, last2 = engine.agglomerate(D, n, "nj", "fast")
    assert _valid_join_list(joins, last2, n)
    j2, l2 = sharded.emulated_solve(cm, -1, w, S, "nj", 4, dtype="f32", p2p=True)
    assert np.array_equal(joins, j2) and np.array_equal(last2, l2)


def test_upgma_10k_float64_join_list_equals_the_oracle():
    """BASELINE config 2 exactly (UPGMASolver, 10 000 cells x 40 characters, plain Hamming, float64): the GPU join
    list against the CPU oracle's (oracle/oracle.c: oracle_upgma, all host threads; about half a minute), with the
    default (dense) and the pruned scan."""
    from cassiopeia_b200 import engine
    from oracle import pyoracle as po
    import os

    n, m, S = 10000, 40, 50
    cm, _ = _workload(n, m, S, 0.1, seed=7)
    D = engine.whd_map(cm, -1, None, S, dtype="f64", method="tensor")
    want, want_last = po.upgma(D[:, :n].cpu().numpy())
    got, got_last = engine.agglomerate(D.clone(), n, "upgma", "strict")
    assert np.array_equal(got, want) and np.array_equal(got_last, want_last)
    os.environ["CAS_PRUNE"] = "1"
    try:
        pruned, pruned_last = engine.agglomerate(D, n, "upgma", "strict")
    finally:
        del os.environ["CAS_PRUNE"]
    assert np.array_equal(pruned, want) and np.array_equal(pruned_last, want_last)


def test_upgma_10k_float64_strict_equals_fast_and_values_are_monotone():
    """BASELINE config 2 (UPGMA, 10k x 40, plain Hamming): the loop has no reductions, so the strict and
    the fast float64 paths must give the identical join list; UPGMA merge heights never decrease."""
    from cassiopeia_b200 import engine

    n, m, S = 10000, 40, 50
    cm, _ = _workload(n, m, S, 0.1, seed=7)
    D = engine.whd_map(cm, -1, None, S, dtype="f64", method="tensor")
    D0 = D[:, :n].clone()
    a, la = engine.agglomerate(D.clone(), n, "upgma", "strict")
    b, lb = engine.agglomerate(D, n, "upgma", "fast")
    assert np.array_equal(a, b) and _valid_join_list(a, la, n)
    # first join is the global minimum of the map with the lowest (i, j)
    iu = np.triu_indices(2000, 1)
    sub = D0[:2000, :2000].cpu().numpy()
    assert D0.cpu().numpy()[a[0, 0], a[0, 1]] <= sub[iu].min()
