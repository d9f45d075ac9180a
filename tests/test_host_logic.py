"""CPU tests of the host-side mirror: prior transformation, row ordering, state compaction,
the tree tail (populate / root / collapse) replayed from the reference's golden join lists,
and constructor / error behaviour."""
import numpy as np
import pandas as pd
import pytest

from conftest import Golden, golden_names
from oracle import pyoracle as po
from oracle import treecmp

import cassiopeia_b200 as cb
from cassiopeia_b200 import host_prep
from cassiopeia_b200.mixins import DistanceSolverError, PriorTransformationError
from cassiopeia_b200.solver import solver_utilities


# ---- transform_priors: test/solver_tests/dissimilarity_functions_test.py:46-85 --------------
PRIORS = {0: {1: 0.5, 2: 0.5}, 1: {1: 0.2, 2: 0.8}, 2: {1: 0.9, 2: 0.1}}


def test_transform_priors_negative_log_inverse_sqrt():
    w = solver_utilities.transform_priors(PRIORS, "negative_log")
    assert w[1][1] == -np.log(0.2) and w[2][2] == -np.log(0.1)
    w = solver_utilities.transform_priors(PRIORS, "inverse")
    assert w[0][1] == 2.0 and w[1][2] == 1 / 0.8
    w = solver_utilities.transform_priors(PRIORS, "square_root_inverse")
    assert w[1][1] == np.sqrt(1 / 0.2)
    assert w == po.transform_priors(PRIORS, "square_root_inverse")


def test_transform_priors_errors():
    with pytest.raises(PriorTransformationError):
        solver_utilities.transform_priors({0: {1: 0.0}}, "negative_log")
    with pytest.raises(PriorTransformationError):
        solver_utilities.transform_priors({0: {1: 1.5}}, "negative_log")
    with pytest.raises(PriorTransformationError):
        solver_utilities.transform_priors(PRIORS, "cube_root")


def test_weighted_hamming_distance_python_definition():
    f = cb.solver.dissimilarity.weighted_hamming_distance
    s1, s2 = [0, 1, 0, -1, 1, 2], [1, 1, 0, 0, 1, 3]
    assert f(s1, s1) == 0
    assert f(s1, s2) == 3 / 5
    w = {0: {1: 2.0}, 5: {2: 3.0, 3: 4.0}}
    assert f(s1, s2, -1, w) == (2.0 + (3.0 + 4.0)) / 5
    assert f([-1, -1], [1, 2]) == 0


# ---- row order / compaction -------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_names())
def test_row_order_matches_reference(name):
    g = Golden(name)
    assert np.array_equal(host_prep.reference_row_order(g.rooted_cm), g.order)


def test_row_order_random_against_loop_restatement():
    rng = np.random.default_rng(3)
    for _ in range(20):
        cm = rng.integers(-1, 3, size=(rng.integers(1, 60), rng.integers(1, 5)))
        assert np.array_equal(host_prep.reference_row_order(cm), po.reference_row_order(cm))


def test_row_order_survives_hash_collisions(monkeypatch):
    """The row hash only proposes groups: with a degenerate hash (every row collides) the verification
    step must notice and the exact byte-string sort must take over; negative and huge states hash too."""
    rng = np.random.default_rng(5)
    cm = rng.integers(-3, 4, size=(300, 6)).astype(np.int64)
    cm[7] = cm[250]
    cm[100, :] = np.iinfo(np.int64).max
    cm[101, :] = np.iinfo(np.int64).min
    want = po.reference_row_order(cm)
    assert np.array_equal(host_prep.reference_row_order(cm), want)
    assert np.array_equal(host_prep._first_occurrence(cm), host_prep._first_occurrence_exact(cm))
    monkeypatch.setattr(host_prep, "_HASH_MULT", np.zeros_like(host_prep._HASH_MULT))
    assert np.array_equal(host_prep.reference_row_order(cm), want)
    # wider than the multiplier table: exact path
    wide = rng.integers(0, 2, size=(20, host_prep._HASH_MULT.size + 1)).astype(np.int64)
    wide[3] = wide[11]
    assert np.array_equal(host_prep.reference_row_order(wide), po.reference_row_order(wide))


def test_compact_states_missing_indicator_extremes():
    """The present-state range is found without gathering the present entries: the indicator may be the
    smallest or the largest value of the matrix, or collide with the code range."""
    base = np.array([[0, 1, 2], [2, 1, 0], [1, 1, 1]], dtype=np.int64)
    for miss in (-1, 7, 10**9, -10**9, 1):
        cm = base.copy()
        cm[0, 0] = miss
        cm[2, 2] = miss
        codes, mcode, table, S = host_prep.compact_states(cm, miss, None)
        present = cm != miss
        assert S == int(cm[present].max()) and np.array_equal(codes[present], cm[present])
        assert (codes[~present] == mcode).all() and not (0 <= mcode <= S)
    allmiss = np.full((2, 3), -1, dtype=np.int64)
    codes, mcode, table, S = host_prep.compact_states(allmiss, -1, None)
    assert (codes == mcode).all() and S == 1


def test_compact_states_direct_and_remapped():
    cm = np.array([[0, 5, -1], [3, 5, 2], [3, 0, 2]], dtype=np.int64)
    codes, miss, table, S = host_prep.compact_states(cm, -1, None)
    assert miss == -1 and table is None and S == 5 and np.array_equal(codes, cm)
    big = np.array([[0, 10**6, -7], [123456, 10**6, 9], [123456, 0, 9]], dtype=np.int64)
    weights = {0: {123456: 1.5}, 1: {10**6: 2.5}, 2: {9: 3.5}}
    codes, miss, table, S = host_prep.compact_states(big, -7, weights)
    assert codes[0, 2] == miss and set(np.unique(codes[codes != miss])) <= set(range(S + 1))
    # equality structure is preserved and weights follow their states
    assert codes[1, 0] == codes[2, 0] != 0 and codes[0, 0] == 0
    assert table[0, codes[1, 0]] == 1.5 and table[1, codes[0, 1]] == 2.5
    # the same distances as the oracle on the raw labels (remap to a dense table by hand)
    raw = {0: {1: 1.5}, 1: {1: 2.5}, 2: {1: 3.5}}
    D_codes = po.whd_map(codes.astype(np.int64), miss, table)
    D_raw = po.whd_map(np.where(big == -7, -7, (big != 0).astype(np.int64)), -7,
                       po.dense_weight_table(raw, 3, 1))
    assert np.array_equal(D_codes, D_raw)


def test_compact_states_missing_weight_raises_keyerror_like_the_reference():
    cm = np.array([[1, 0], [2, 0]], dtype=np.int64)
    with pytest.raises(KeyError):
        host_prep.compact_states(cm, -1, {0: {1: 0.5}, 1: {}})  # state 2 of character 0 has no weight
    # a character with a single distinct state is never looked up
    host_prep.compact_states(np.array([[1, 3], [2, 3]]), -1, {0: {1: 1.0, 2: 1.0}, 1: {}})


def test_ambiguous_states_are_rejected():
    df = pd.DataFrame({"a": [(1, 2), 0], "b": [1, 1]}, index=["x", "y"])
    with pytest.raises(DistanceSolverError):
        host_prep.as_int_matrix(df)


# ---- tree tail replayed from the reference's join lists -----------------------------------------
def _tree_and_names(g):
    names = [f"c{i}" for i in range(g.n)]
    df = pd.DataFrame(g.cm, index=names)
    tree = cb.CassiopeiaTree(character_matrix=df, priors=g.prior_dict(), missing_state_indicator=g.missing)
    ext = names + ["root"]
    map_names = [ext[i] for i in g.order]
    return tree, names, map_names


@pytest.mark.parametrize("collapse", [False, True])
@pytest.mark.parametrize("name", golden_names())
def test_tail_reproduces_reference_tree(name, collapse):
    g = Golden(name)
    tree, names, map_names = _tree_and_names(g)
    if g.algo == "nj":
        solver = cb.NeighborJoiningSolver(add_root=True)
        # what setup_root_finder leaves behind (NeighborJoiningSolver.py:277-309)
        tree.root_sample_name = "root"
    else:
        solver = cb.UPGMASolver()
        tree.root_sample_name = "root"
    n = len(map_names)
    used = set(g.joins.ravel().tolist())
    last2 = np.array(sorted(set(range(2 * n - 2)) - used))
    solver.finish_tree(tree, map_names, g.joins, last2, None, collapse)
    hs, ss, ds = treecmp.clade_table(tree.get_tree_topology(), {nm: i for i, nm in enumerate(names)})
    ghs, gss, gds = g.clades(collapse)
    assert np.array_equal(hs, ghs) and np.array_equal(ss, gss)
    # depth of every clade (unit branch lengths) via node times
    assert tree.get_tree_topology().number_of_nodes() == int(g.z[f"n_nodes_{'collapsed' if collapse else 'plain'}"])
    assert sorted(tree.leaves) == sorted(names)
    times = tree.get_times()
    assert times[tree.root] == 0 and all(t >= 1 for nm, t in times.items() if nm != tree.root)


def test_newick_is_iterative_and_wellformed():
    g = Golden("sim64_s0_upgma_plain")
    tree, names, map_names = _tree_and_names(g)
    tree.root_sample_name = "root"
    n = len(map_names)
    last2 = np.array(sorted(set(range(2 * n - 2)) - set(g.joins.ravel().tolist())))
    cb.UPGMASolver().finish_tree(tree, map_names, g.joins, last2)
    nwk = tree.get_newick()
    assert nwk.endswith(";") and nwk.count("(") == nwk.count(")") == len(tree.internal_nodes)


# ---- constructors and errors ---------------------------------------------------------------------
def test_constructor_keywords_and_invalid_implementations():
    s = cb.NeighborJoiningSolver(add_root=True, prior_transformation="inverse", threads=2)
    assert s.add_root and s.prior_transformation == "inverse" and s.threads == 2 and s.precision == "exact"
    assert cb.NeighborJoiningSolver(fast=True).precision == "fast"
    assert cb.NeighborJoiningSolver(fast=True, implementation="b200_exact").precision == "exact"
    assert cb.UPGMASolver(fast=True, implementation="b200_strict").precision == "strict"
    assert cb.UPGMASolver().add_root is True
    for bad in ("ccphylo_dnj", "nonsense"):
        with pytest.raises(DistanceSolverError):
            cb.NeighborJoiningSolver(fast=True, implementation=bad)
    with pytest.raises(DistanceSolverError):
        cb.UPGMASolver(fast=True, implementation="ccphylo_nj")


def test_solver_is_picklable():
    import pickle

    s = pickle.loads(pickle.dumps(cb.NeighborJoiningSolver(add_root=True)))
    assert s.add_root


CM5 = pd.DataFrame.from_dict({"a": [0, 1, 2], "b": [1, 1, 2], "c": [2, 2, 2], "d": [1, 1, 1], "e": [0, 0, 0]},
                             orient="index", columns=["x1", "x2", "x3"])
DELTA5 = pd.DataFrame([[0, 15, 21, 17, 12], [15, 0, 10, 6, 17], [21, 10, 0, 10, 23], [17, 6, 10, 0, 19],
                       [12, 17, 23, 19, 0]], index=list("abcde"), columns=list("abcde"))


def test_root_and_function_error_cases():
    # neighborjoining_solver_test.py:132-169 -- these raise before any GPU work
    nothing = cb.NeighborJoiningSolver(dissimilarity_function=None, add_root=False)
    no_root_tree = cb.CassiopeiaTree(character_matrix=CM5, dissimilarity_map=DELTA5)
    with pytest.raises(DistanceSolverError):
        nothing.solve(no_root_tree)
    with pytest.raises(DistanceSolverError):
        cb.NeighborJoiningSolver(dissimilarity_function=None, add_root=True).solve(no_root_tree)
    root_only = cb.CassiopeiaTree(character_matrix=CM5, root_sample_name="b")
    with pytest.raises(DistanceSolverError):
        nothing.solve(root_only)
    # arbitrary python callables have no CUDA kernel and there is no CPU fallback
    with pytest.raises(DistanceSolverError):
        cb.NeighborJoiningSolver(dissimilarity_function=lambda a, b, m, w: 0.0, add_root=False).solve(
            cb.CassiopeiaTree(character_matrix=CM5, root_sample_name="b"))


def test_tree_container_map_accessors():
    t = cb.CassiopeiaTree(character_matrix=CM5, dissimilarity_map=DELTA5)
    got = t.get_dissimilarity_map()
    got.iloc[0, 1] = -1  # a copy
    assert t.get_dissimilarity_map().iloc[0, 1] == 15
    t.set_dissimilarity("root", {**{k: 1.0 for k in "abcde"}, "root": 0})
    m = t.get_dissimilarity_map()
    assert list(m.index) == list("abcde") + ["root"] and m.loc["root", "c"] == 1.0 and m.loc["root", "root"] == 0
    with pytest.warns(Warning):
        t.set_dissimilarity_map(DELTA5.iloc[:4, :4])
