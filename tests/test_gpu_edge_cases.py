"""Edge cases through the C ABI vs the CPU oracle: tiny n, ragged shapes, many states, all-missing rows,
all-duplicate rows, one character, more than 64 characters (multi-word presence masks)."""
import numpy as np
import pandas as pd
import pytest

from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def _case(n, m, S, missing_frac, seed, dup_frac=0.0, all_missing_rows=0):
    rng = np.random.default_rng(seed)
    cm = rng.integers(0, S + 1, size=(n, m))
    cm[rng.random((n, m)) < 0.4] = 0
    cm[rng.random((n, m)) < missing_frac] = -1
    for r in range(1, n):
        if rng.random() < dup_frac:
            cm[r] = cm[rng.integers(0, r)]
    for r in range(min(all_missing_rows, n)):
        cm[n - 1 - r] = -1
    w = np.concatenate([np.zeros((m, 1)), rng.uniform(0.05, 9.0, size=(m, S))], axis=1)
    return cm.astype(np.int32), w


SHAPES = [(2, 3, 2), (3, 1, 1), (4, 7, 3), (5, 65, 2), (33, 130, 5), (65, 9, 300), (129, 40, 50), (200, 2, 1)]


@pytest.mark.parametrize("n,m,S", SHAPES)
@pytest.mark.parametrize("weighted", [False, True])
def test_map_matches_oracle_on_ragged_shapes(n, m, S, weighted):
    from cassiopeia_b200 import engine

    for missing, dup, allm in ((0.0, 0.0, 0), (0.5, 0.3, 0), (1.0, 0.0, 0), (0.2, 0.9, 2)):
        cm, w = _case(n, m, S, missing, seed=n * 7 + m, dup_frac=dup, all_missing_rows=allm)
        ww = w if weighted else None
        want = po.whd_map(cm.astype(np.int64), -1, ww)
        got = engine.whd_map(cm, -1, ww, S, dtype="f64", method="exact")[:, :n].cpu().numpy()
        assert np.array_equal(got, want), (n, m, S, missing)
        if m <= 512:
            gt = engine.whd_map(cm, -1, ww, S, dtype="f64", method="tensor")[:, :n].cpu().numpy()
            if weighted:
                assert (np.abs(gt - want) / np.maximum(np.abs(want), 1.0)).max() <= 1e-6
            else:
                assert np.array_equal(gt, want), ("tensor", n, m, S, missing)


@pytest.mark.parametrize("n,m,S", SHAPES)
@pytest.mark.parametrize("algo", ["nj", "upgma"])
def test_strict_loops_match_oracle_on_edge_inputs(n, m, S, algo):
    from cassiopeia_b200 import engine

    for missing, dup, allm in ((0.0, 0.0, 0), (0.5, 0.5, 0), (1.0, 0.0, 0), (0.1, 0.95, 1)):
        cm, w = _case(n, m, S, missing, seed=n + 3 * m, dup_frac=dup, all_missing_rows=allm)
        D = po.whd_map(cm.astype(np.int64), -1, w)
        want, want_last = (po.nj if algo == "nj" else po.upgma)(D)
        Dg = engine.whd_map(cm, -1, w, S, dtype="f64", method="exact")
        got, got_last = engine.agglomerate(Dg, n, algo, "strict")
        assert np.array_equal(got, want.reshape(-1, 2)), (n, m, S, missing, algo)
        assert np.array_equal(np.sort(got_last), np.sort(want_last))


def test_solver_on_two_and_three_samples():
    import cassiopeia_b200 as cb

    for n in (2, 3):
        cm = pd.DataFrame(np.arange(n * 4).reshape(n, 4) % 3, index=[f"s{i}" for i in range(n)])
        t = cb.CassiopeiaTree(character_matrix=cm)
        cb.NeighborJoiningSolver(add_root=True).solve(t)
        assert sorted(t.leaves) == sorted(cm.index)
        t = cb.CassiopeiaTree(character_matrix=cm)
        cb.UPGMASolver().solve(t, collapse_mutationless_edges=True)
        assert sorted(t.leaves) == sorted(cm.index)


def test_existing_map_with_implicit_root_uses_the_gpu_row():
    """NJ add_root with a stored map: only the root row is added (NeighborJoiningSolver.py:293-307)."""
    import cassiopeia_b200 as cb

    cm = pd.DataFrame({"x": [1, 1, 2, 0], "y": [0, 2, 2, -1], "z": [3, 3, 0, 1]}, index=list("abcd"))
    base = cb.CassiopeiaTree(character_matrix=cm)
    base.compute_dissimilarity_map(cb.solver.dissimilarity.weighted_hamming_distance)
    stored = base.get_dissimilarity_map()
    tree = cb.CassiopeiaTree(character_matrix=cm, dissimilarity_map=stored)
    cb.NeighborJoiningSolver(add_root=True).solve(tree)
    m = tree.get_dissimilarity_map()
    f = cb.solver.dissimilarity.weighted_hamming_distance
    for leaf in "abcd":
        assert m.loc["root", leaf] == f([0, 0, 0], cm.loc[leaf].values, -1, None) == m.loc[leaf, "root"]
    assert sorted(tree.leaves) == list("abcd")
