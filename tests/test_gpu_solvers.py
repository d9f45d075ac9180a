"""Solver-level parity on the GPU: the reference's own unit-test cases re-pointed at the
B200 classes (test/solver_tests/neighborjoining_solver_test.py, upgma_test.py) and whole
solves compared with the reference's golden trees."""
import itertools

import networkx as nx
import numpy as np
import pandas as pd
import pytest

from conftest import Golden, golden_names
from oracle import treecmp

import cassiopeia_b200 as cb
from cassiopeia_b200.mixins import DistanceSolverError

pytestmark = pytest.mark.gpu

WHD = cb.solver.dissimilarity.weighted_hamming_distance


def triplet(tri, T):
    """Which pair of a leaf triplet shares the deepest ancestor (the reference tests' topology check)."""
    anc = [set(nx.ancestors(T, x)) for x in tri]
    ab, ac, bc = len(anc[0] & anc[1]), len(anc[0] & anc[2]), len(anc[1] & anc[2])
    if ab > bc and ab > ac:
        return "ab"
    if ac > bc and ac > ab:
        return "ac"
    if bc > ab and bc > ac:
        return "bc"
    return "-"


def same_triplets(leaves, expected_edges, observed):
    exp = nx.DiGraph()
    exp.add_edges_from(expected_edges)
    for tri in itertools.combinations(leaves, 3):
        assert triplet(tri, exp) == triplet(tri, observed), tri


CM5 = pd.DataFrame.from_dict({"a": [0, 1, 2], "b": [1, 1, 2], "c": [2, 2, 2], "d": [1, 1, 1], "e": [0, 0, 0]},
                             orient="index", columns=["x1", "x2", "x3"])
NJ_DELTA = pd.DataFrame([[0, 15, 21, 17, 12], [15, 0, 10, 6, 17], [21, 10, 0, 10, 23], [17, 6, 10, 0, 19],
                         [12, 17, 23, 19, 0]], index=list("abcde"), columns=list("abcde"))
UP_DELTA = pd.DataFrame([[0, 17, 21, 31, 23], [17, 0, 30, 34, 21], [21, 30, 0, 28, 39], [31, 34, 28, 0, 43],
                         [23, 21, 39, 43, 0]], index=list("abcde"), columns=list("abcde"))
PP_CM = pd.DataFrame.from_dict({"a": [1, 1, 0], "b": [1, 2, 0], "c": [1, 2, 1], "d": [2, 0, 0], "e": [2, 0, 2]},
                               orient="index", columns=["x1", "x2", "x3"])
PP_PRIORS = {0: {1: 0.5, 2: 0.5}, 1: {1: 0.2, 2: 0.8}, 2: {1: 0.3, 2: 0.7}}


# ------------------------------------------------------------------ single-step hooks
def test_compute_q_matches_reference_values():
    q = cb.NeighborJoiningSolver.compute_q(NJ_DELTA.values)
    expected = np.array([[0, -22.67, -22, -22, -33.33], [-22.67, 0, -27.33, -27.33, -22.67],
                         [-22, -27.33, 0, -28.67, -22], [-22, -27.33, -28.67, 0, -22],
                         [-33.33, -22.67, -22, -22, 0]])
    assert np.allclose(q, expected, atol=0.1)
    # exact restatement: d_ij - (1/(n-2)) * (rs_i + rs_j) in float64
    rs = NJ_DELTA.values.astype(float).sum(axis=1)
    assert q[4, 0] == 12.0 - (1 / 3) * (rs[4] + rs[0])


def test_find_cherry_nj_and_upgma():
    i, j = cb.NeighborJoiningSolver(add_root=True).find_cherry(NJ_DELTA.values)
    assert (NJ_DELTA.index[i], NJ_DELTA.index[j]) == ("a", "e")
    i, j = cb.UPGMASolver().find_cherry(UP_DELTA.values)
    assert (UP_DELTA.index[i], UP_DELTA.index[j]) == ("a", "b")


def test_update_dissimilarity_map_nj_exact():
    out = cb.NeighborJoiningSolver(add_root=True).update_dissimilarity_map(NJ_DELTA, ("a", "e"), "f")
    exp = pd.DataFrame([[0, 10, 16, 12], [10, 0, 10, 6], [16, 10, 0, 10], [12, 6, 10, 0]],
                       index=list("fbcd"), columns=list("fbcd"))
    assert list(out.index) == ["b", "c", "d", "f"]
    for r in exp.index:
        for c in exp.index:
            assert out.loc[r, c] == exp.loc[r, c]


def test_update_dissimilarity_map_upgma_two_steps_exact():
    s = cb.UPGMASolver()
    d1 = s.update_dissimilarity_map(UP_DELTA, ("a", "b"), "ab")
    exp1 = {"ab": {"c": 25.5, "d": 32.5, "e": 22.0}, "c": {"d": 28, "e": 39}, "d": {"e": 43}}
    for r, row in exp1.items():
        for c, v in row.items():
            assert d1.loc[r, c] == v and d1.loc[c, r] == v
    d2 = s.update_dissimilarity_map(d1, ("ab", "e"), "abe")
    assert d2.loc["abe", "c"] == 30 and d2.loc["abe", "d"] == 36 and d2.loc["c", "d"] == 28


# ------------------------------------------------------------------ reference solve cases
def test_nj_basic_solver_with_explicit_root_and_collapse():
    tree = cb.CassiopeiaTree(character_matrix=CM5, dissimilarity_map=NJ_DELTA, root_sample_name="b")
    solver = cb.NeighborJoiningSolver(add_root=True)
    solver.solve(tree)
    assert sorted(tree.leaves) == ["a", "c", "d", "e"] and len(tree.edges) == 6
    same_triplets(["a", "c", "d", "e"],
                  [("5", "a"), ("5", "e"), ("6", "5"), ("b", "6"), ("6", "7"), ("7", "d"), ("7", "c")],
                  tree.get_tree_topology())
    solver.solve(tree, collapse_mutationless_edges=True)  # second solve on the same tree
    exp_edges = [("6", "a"), ("6", "e"), ("b", "6"), ("6", "d"), ("6", "c")]
    obs = tree.get_tree_topology()
    same_triplets(["a", "c", "d", "e"], exp_edges, obs)
    exp = nx.Graph(exp_edges)
    und = obs.to_undirected()
    leaves = tree.leaves
    for x, y in itertools.combinations(leaves, 2):
        assert nx.shortest_path_length(und, x, y) == nx.shortest_path_length(exp, x, y)
    assert tree.get_dissimilarity_map().loc["a", "b"] == 15


def test_nj_solver_with_priors():
    tree = cb.CassiopeiaTree(character_matrix=PP_CM, priors=PP_PRIORS)
    solver = cb.NeighborJoiningSolver(dissimilarity_function=WHD, add_root=True)
    edges = [("root", "7"), ("7", "6"), ("6", "d"), ("6", "e"), ("7", "8"), ("8", "a"), ("8", "9"),
             ("9", "b"), ("9", "c")]
    solver.solve(tree)
    same_triplets(list("abcde"), edges, tree.get_tree_topology())
    solver.solve(tree, collapse_mutationless_edges=True)
    same_triplets(list("abcde"), edges, tree.get_tree_topology())
    assert tree.root_sample_name == "root" and "root" not in tree.character_matrix.index
    assert "root" in tree.get_dissimilarity_map().index


def test_nj_custom_function_needs_a_precomputed_map():
    def delta_fn(x, y, missing_state, priors):
        return sum(1 for a, b in zip(x, y) if a != b)

    solver = cb.NeighborJoiningSolver(dissimilarity_function=delta_fn, add_root=True)
    with pytest.raises(DistanceSolverError):
        solver.solve(cb.CassiopeiaTree(character_matrix=PP_CM))
    # with a stored map only the implicit root row is evaluated with the user's function
    names = list(PP_CM.index)
    D = pd.DataFrame([[delta_fn(PP_CM.loc[a].values, PP_CM.loc[b].values, -1, None) for b in names] for a in names],
                     index=names, columns=names, dtype=float)
    tree = cb.CassiopeiaTree(character_matrix=PP_CM, dissimilarity_map=D)
    solver.solve(tree)
    same_triplets(list("abcde"),
                  [("root", "9"), ("9", "8"), ("9", "7"), ("7", "6"), ("7", "a"), ("6", "b"), ("6", "c"),
                   ("8", "e"), ("8", "d")], tree.get_tree_topology())
    m = tree.get_dissimilarity_map()
    assert m.loc["root", "a"] == 2 and m.loc["e", "root"] == 2 and m.loc["root", "root"] == 0


def test_upgma_basic_and_priors_and_duplicates():
    tree = cb.CassiopeiaTree(character_matrix=CM5, dissimilarity_map=UP_DELTA)
    cb.UPGMASolver().solve(tree)
    assert len(tree.edges) == 8 and sorted(tree.leaves) == list("abcde")
    same_triplets(list("abcde"),
                  [("5", "a"), ("5", "b"), ("6", "5"), ("6", "e"), ("7", "c"), ("7", "d"), ("root", "6"),
                   ("root", "7")], tree.get_tree_topology())

    tree = cb.CassiopeiaTree(character_matrix=PP_CM, priors=PP_PRIORS)
    cb.UPGMASolver(dissimilarity_function=WHD).solve(tree)
    assert tree.get_dissimilarity_map().loc["a", "b"] == (-np.log(0.2) - np.log(0.8)) / 3
    solver = cb.UPGMASolver(dissimilarity_function=WHD)
    same_triplets(list("abcde"),
                  [("root", "a"), ("root", "7"), ("7", "8"), ("7", "9"), ("8", "d"), ("8", "e"), ("9", "b"),
                   ("9", "c")], tree.get_tree_topology())
    solver.solve(tree, collapse_mutationless_edges=True)
    same_triplets(list("abcde"),
                  [("root", "a"), ("root", "8"), ("root", "9"), ("8", "d"), ("8", "e"), ("9", "b"), ("9", "c")],
                  tree.get_tree_topology())

    tree = cb.CassiopeiaTree(character_matrix=PP_CM)
    cb.UPGMASolver(dissimilarity_function=WHD).solve(tree)
    assert tree.get_dissimilarity_map().loc["d", "e"] == 1 / 3

    dup = pd.DataFrame.from_dict({"a": [1, -1, 0], "b": [1, 2, 1], "c": [1, -1, 1], "d": [2, 0, -1],
                                  "e": [2, 0, 2], "f": [2, 0, 2]}, orient="index", columns=["x1", "x2", "x3"])
    tree = cb.CassiopeiaTree(character_matrix=dup)
    cb.UPGMASolver(dissimilarity_function=WHD).solve(tree)
    assert tree.get_dissimilarity_map().loc["b", "d"] == 1.5
    same_triplets(list("abcdef"),
                  [("root", "9"), ("root", "8"), ("9", "a"), ("9", "6"), ("6", "b"), ("6", "c"), ("8", "7"),
                   ("8", "f"), ("7", "d"), ("7", "e")], tree.get_tree_topology())


# ------------------------------------------------------------------ golden whole solves
def _make(g):
    names = [f"c{i}" for i in range(g.n)]
    df = pd.DataFrame(g.cm, index=names)
    return names, lambda: cb.CassiopeiaTree(character_matrix=df, priors=g.prior_dict(),
                                            missing_state_indicator=g.missing)


@pytest.mark.parametrize("name", golden_names())
def test_whole_solve_matches_reference_tree(name):
    g = Golden(name)
    names, make = _make(g)
    leaf_index = {nm: i for i, nm in enumerate(names)}
    for collapse in (False, True):
        tree = make()
        if g.algo == "nj":
            solver = cb.NeighborJoiningSolver(add_root=True, prior_transformation=g.prior_transformation)
        else:
            solver = cb.UPGMASolver(prior_transformation=g.prior_transformation)
        solver.solve(tree, collapse_mutationless_edges=collapse)
        joins, last2, map_names = solver.last_join_list
        assert np.array_equal(joins, g.joins)
        ext = names + ["root"]
        assert map_names == [ext[i] for i in g.order]
        hs, ss, _ = treecmp.clade_table(tree.get_tree_topology(), leaf_index)
        ghs, gss, _ = g.clades(collapse)
        assert np.array_equal(hs, ghs) and np.array_equal(ss, gss)
        stored = tree.get_dissimilarity_map()
        assert stored is not None and stored.to_numpy().dtype == np.float64
        if g.D is not None:
            assert np.array_equal(stored.to_numpy(), g.D)


@pytest.mark.parametrize("name", ["sim256_s1_nj_priors", "sim1024_s0_nj_priors", "sim1024_s0_upgma_plain"])
def test_fast_mode_topology_close_to_reference(name):
    """float32 fast mode: RF = 0 is only promised when no near-ties occur; report the distance and
    require it to stay small on the simulated cases."""
    g = Golden(name)
    names, make = _make(g)
    leaf_index = {nm: i for i, nm in enumerate(names)}
    strict_tree, fast_tree = make(), make()
    cls = cb.NeighborJoiningSolver if g.algo == "nj" else cb.UPGMASolver
    kw = {"add_root": True} if g.algo == "nj" else {}
    cls(**kw).solve(strict_tree)
    cls(fast=True, implementation="b200_fast", **kw).solve(fast_tree)
    rf = treecmp.rf_distance(strict_tree.get_tree_topology(), fast_tree.get_tree_topology(), leaf_index)
    n_splits = 2 * (len(names) - 3)
    print(f"{name}: RF(fast, strict) = {rf} of {n_splits}")
    assert rf <= 0.05 * n_splits


def test_layer_and_map_only_entry_points():
    g = Golden("sim64_s0_upgma_priors")
    names, make = _make(g)
    tree = make()
    tree.compute_dissimilarity_map(WHD, "negative_log")
    assert np.array_equal(tree.get_dissimilarity_map().to_numpy(), g.D)
    cond = cb.data.compute_dissimilarity_map(g.ordered_cm, g.ordered_cm.shape[0], WHD,
                                             cb.solver.solver_utilities.transform_priors(g.prior_dict()), -1)
    iu = np.triu_indices(g.D.shape[0], 1)
    assert np.array_equal(cond, g.D[iu])
    # a layer replaces the character matrix for UPGMA (DistanceSolver.py:386-395)
    tree2 = make()
    tree2.layers["alt"] = tree2.character_matrix.copy()
    cb.UPGMASolver().solve(tree2, layer="alt")
    assert np.array_equal(tree2.get_dissimilarity_map().to_numpy(), g.D)
